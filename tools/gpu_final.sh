#!/bin/bash
# Final check of a round: GPU tests, smoke, and that `bench.py` prints exactly one JSON line on stdout (1 GPU here; N > 1 with gpu_multi8.sh)
TAG=${1:-final}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -2 $OUT/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1; echo "smoke exit $?"; tail -1 $OUT/smoke.log
timeout 900 python bench.py --steps 3 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; wc -l $OUT/bench.json; python -c "
import json;d=json.loads(open('$OUT/bench.json').read());print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['cpu_baseline'],d['render_psfnet']['value'],d['render_psfnet']['roofline']['frac'])"
NCTA=2 timeout 300 python tools/fused_debug.py 1024 1536 4 > $OUT/fused_big.log 2>&1; grep -v "layer [1-8]:" $OUT/fused_big.log | tail -7
