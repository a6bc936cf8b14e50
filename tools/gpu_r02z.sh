#!/bin/bash
TAG=${1:-r02z}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/gpu_tests.sh $TAG | tail -4
for lens in rf50mm rf35mm; do
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:psf_bank_run -s 2 -c 1 --csv --log-file $OUT/bank_traffic_$lens.csv python bench.py --lens $lens --steps 1 --warmup 2 --quick > /dev/null 2> $OUT/bank_traffic_$lens.err; echo "traffic $lens exit $?"; tail -3 $OUT/bank_traffic_$lens.csv | cut -c1-250
done
