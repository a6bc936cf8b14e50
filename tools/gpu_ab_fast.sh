#!/bin/bash
TAG=${1:-r02w}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests/test_engine_gpu.py -m gpu -q -x -k "two_ray or packed_strict or psf_bank or pixel_assignment or depth_sweep or generic or smoke or ragged or empty or vignet" > $OUT/pytest_fast.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_fast.log
for m in adaptive fast strict; do timeout 600 python bench.py --numerics $m --steps 6 --warmup 3 --quick 2> $OUT/q_$m.err | cut -c1-120; done
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so); for m in adaptive fast strict; do SDIRT_ENGINE_LIB=$v timeout 600 python bench.py --numerics $m --steps 6 --warmup 3 --quick 2> $OUT/q_${m}_$n.err | cut -c1-120; done; done
