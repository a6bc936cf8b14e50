#!/bin/bash
# One GPU-box visit (later rounds): parity tests, smoke, the contract bench (both arms), ncu launch list + full captures of the
# bank kernel and of the fused PSF-MLP kernel, PSFNet.render / fitting scratch benches.  Usage: bash tools/gpu_round2.sh tag
TAG=${1:-r02}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu" ; timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E 'passed|failed' $OUT/pytest_gpu.log | tail -3
echo "== smoke" ; timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?" | tee -a $OUT/smoke.log; tail -2 $OUT/smoke.log
echo "== bench" ; timeout 1200 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; cat $OUT/bench.json; tail -3 $OUT/bench.err
echo "== bench reference arm" ; timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > $OUT/bench_ref.json 2> $OUT/bench_ref.err; cat $OUT/bench_ref.json
echo "== render C4 / C5 shapes"; timeout 600 python tools/render_c4_bench.py 1024 1536 2 > $OUT/render_c4.log 2>&1; cat $OUT/render_c4.log
timeout 600 python tools/render_c4_bench.py 512 768 4 > $OUT/render_c5.log 2>&1; head -4 $OUT/render_c5.log
echo "== fused MLP kernel"; timeout 300 python tools/fused_debug.py 1024 1536 4 > $OUT/fused_big.log 2>&1; grep -v "layer [1-8]:" $OUT/fused_big.log
echo "== fit bench"; timeout 600 python tools/fit_bench.py > $OUT/fit_bench.log 2>&1; grep -E "ms /|get_test" $OUT/fit_bench.log
echo "== ncu launch list of the bench command"
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $OUT/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.log 2>&1; echo "ncu launches exit $?"
echo "== dram traffic of the bank kernel at the bench workload"
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:psf_bank_run -s 1 -c 1 --csv --log-file $OUT/bank_traffic.csv python bench.py --steps 1 --warmup 3 --no-cpu > $OUT/bank_traffic.log 2>&1; echo "ncu traffic exit $?"
echo "== ncu full: bank kernel, fused MLP kernel"
QB_MODES=adaptive timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank -s 1 -c 1 -o $OUT/prof_bank -f python tools/quick_bench.py rf50mm 592 262144 > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
NCTA=2 timeout 600 ncu --set full --import-source on --clock-control none -k regex:mlp_fused_pred -c 1 -f -o $OUT/prof_fused python tools/fused_debug.py 1024 1536 4 > $OUT/ncu_fused.log 2>&1; echo "ncu fused exit $?"
ls -la $OUT
