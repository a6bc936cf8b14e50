#!/bin/bash
# Round-2 visit A: parity of the specialised strict kernel, its build variants, and a full ncu capture of it.
TAG=${1:-r02a}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -15
echo "== strict_check base"; timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_base_rf50.log 2>&1; tail -9 $OUT/strict_base_rf50.log
timeout 600 python tools/strict_check.py rf35mm 592 1048576 > $OUT/strict_base_rf35.log 2>&1; tail -3 $OUT/strict_base_rf35.log
for v in noseed ctas4 ctas2 shortdiv; do
  echo "== variant $v"; SDIRT_ENGINE_LIB=$PWD/sdirt_b200/_lib/variants/$v.so timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_${v}_rf50.log 2>&1; tail -3 $OUT/strict_${v}_rf50.log
done
echo "== quick bench (all modes, sorted samples for strict)"; QB_STRICT_SORTED=1 timeout 600 python tools/quick_bench.py rf50mm 1184 1048576 > $OUT/quick_rf50.log 2>&1; cat $OUT/quick_rf50.log
echo "== generic strict kernel for comparison"; SDIRT_DEBUG_GENERIC_STRICT=1 QB_MODES=strict timeout 600 python tools/quick_bench.py rf50mm 1184 1048576 > $OUT/quick_rf50_generic.log 2>&1; cat $OUT/quick_rf50_generic.log
echo "== ncu full on the strict kernel"
SC_MODES=strict timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank_run -s 1 -c 1 -o $OUT/prof_bank_strict -f python tools/strict_check.py rf50mm 592 262144 > $OUT/ncu_full_strict.log 2>&1; echo "ncu full exit $?"
ls -la $OUT
