#!/bin/bash
# A/B of library variants on the contract workload (kernel-only number of bench.py --quick)
TAG=${1:-ab}; OUT=gpurun_out/$TAG; mkdir -p $OUT
MODE=${2:-strict}
timeout 600 python bench.py --numerics $MODE --steps 4 --warmup 2 --quick 2> $OUT/base.err | tee $OUT/base.json
for v in $(ls sdirt_b200/_lib/variants | sed 's/.so//'); do
  SDIRT_ENGINE_LIB=$PWD/sdirt_b200/_lib/variants/$v.so timeout 600 python bench.py --numerics $MODE --steps 4 --warmup 2 --quick 2> $OUT/$v.err | tee $OUT/$v.json
done
