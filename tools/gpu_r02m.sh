#!/bin/bash
TAG=${1:-r02m}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -15
echo "== strict_check"; SC_MODES=strict timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_base_rf50.log 2>&1; tail -2 $OUT/strict_base_rf50.log
timeout 600 python tools/strict_check.py rf35mm 592 1048576 > $OUT/strict_base_rf35.log 2>&1; tail -2 $OUT/strict_base_rf35.log
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so); SDIRT_ENGINE_LIB=$v timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_${n}_rf50.log 2>&1; echo "variant $n"; tail -2 $OUT/strict_${n}_rf50.log; done
for m in strict; do timeout 600 python bench.py --numerics $m --steps 6 --warmup 3 --quick 2> $OUT/q_$m.err | tee $OUT/q_$m.json; done
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so); SDIRT_ENGINE_LIB=$v timeout 600 python bench.py --numerics strict --steps 6 --warmup 3 --quick 2> $OUT/q_strict_$n.err | tee $OUT/q_strict_$n.json; done
bash tools/gpu_ncu_strict.sh $TAG
