#!/bin/bash
TAG=${1:-r02k}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -15
echo "== strict_check"; SC_MODES=strict,adaptive,fast timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_base_rf50.log 2>&1; tail -4 $OUT/strict_base_rf50.log
timeout 600 python tools/strict_check.py rf35mm 592 1048576 > $OUT/strict_base_rf35.log 2>&1; tail -2 $OUT/strict_base_rf35.log
for m in strict adaptive; do timeout 600 python bench.py --numerics $m --steps 6 --warmup 3 --quick 2> $OUT/q_$m.err | tee $OUT/q_$m.json; done
