#!/bin/bash
TAG=${1:-r02h}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -15
echo "== strict_check base"; SC_MODES=strict,adaptive,fast timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_base_rf50.log 2>&1; tail -4 $OUT/strict_base_rf50.log
timeout 600 python tools/strict_check.py rf35mm 592 1048576 > $OUT/strict_base_rf35.log 2>&1; tail -2 $OUT/strict_base_rf35.log
for v in $(ls sdirt_b200/_lib/variants | sed 's/.so//'); do
  echo "== variant $v"; SC_MODES=strict,adaptive,fast SDIRT_ENGINE_LIB=$PWD/sdirt_b200/_lib/variants/$v.so timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_${v}_rf50.log 2>&1; tail -4 $OUT/strict_${v}_rf50.log
done
echo "== bench strict"; timeout 900 python bench.py --numerics strict --steps 4 --warmup 3 --no-cpu > $OUT/bench_strict.json 2> $OUT/bench_strict.err; echo "exit $?"; python -c "
import json;d=json.loads(open('$OUT/bench_strict.json').read());print('strict bench value',d['value'],'e2e',d['e2e']['value'],'ms/step',d['ms_per_step'],d['numerics_modes_rays_per_s'])"
ls $OUT
