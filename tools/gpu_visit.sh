#!/bin/bash
# Short GPU-box visit: all GPU tests + the scratch benches of the PSFNet path.  Usage: bash tools/gpu_visit.sh tag
TAG=${1:-visit}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?"; tail -6 $OUT/pytest_gpu.log
timeout 600 python tools/fit_bench.py > $OUT/fit_bench.log 2>&1; grep -E "ms /|get_test" $OUT/fit_bench.log
timeout 600 python tools/render_c4_bench.py 512 768 4 > $OUT/render_c5.log 2>&1; cat $OUT/render_c5.log
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; python -c "
import json;d=json.load(open('$OUT/bench.json'));print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['render_psfnet'])"; tail -3 $OUT/bench.err
