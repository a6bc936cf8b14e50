#!/bin/bash
# One GPU-box visit: parity tests, smoke, scratch timings, the contract bench, and the ncu evidence.
# Usage (from the repo root, under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E 'passed|failed' $OUT/pytest_gpu.log | tail -3
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -3 $OUT/smoke.log
echo "== quick bench" ; timeout 600 python tools/quick_bench.py rf50mm 1184 1048576 > $OUT/quick_rf50.log 2>&1; cat $OUT/quick_rf50.log
timeout 600 python tools/quick_bench.py rf35mm 1184 1048576 > $OUT/quick_rf35.log 2>&1; cat $OUT/quick_rf35.log
echo "== bench" ; timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== ncu launch list of the bench command"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches exit $?"
echo "== dram traffic of the bank kernel at the bench workload"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:psf_bank_run -s 1 -c 1 --csv --log-file $OUT/bank_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/bank_traffic.log 2>&1; echo "ncu traffic exit $?"
echo "== ncu full on the fused bank kernel"
QB_MODES=adaptive timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank -s 1 -c 1 -o $OUT/prof_bank -f python tools/quick_bench.py rf50mm 592 262144 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
QB_MODES=fast timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank -s 1 -c 1 -o $OUT/prof_bank_fast -f python tools/quick_bench.py rf50mm 592 262144 > $OUT/ncu_full_fast.log 2>&1; echo "ncu full (fast) exit $?"
echo "== render bench"; timeout 600 python tools/render_bench.py 1024 1536 2 21 > $OUT/render_bench.log 2>&1; cat $OUT/render_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_ -s 2 -c 1 -o $OUT/prof_render -f python tools/render_bench.py 512 768 1 21 > $OUT/ncu_render.log 2>&1; echo "ncu render exit $?"
ls -la $OUT
