#!/bin/bash
TAG=${1:-r02B}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1500 python -m pytest tests -m gpu -q -k "render or tone or smoke or focal" > $OUT/pytest_render.log 2>&1; echo "pytest exit $?"; tail -3 $OUT/pytest_render.log
timeout 600 python tools/render_bench.py > $OUT/render_bench.log 2>&1; cat $OUT/render_bench.log | tail -12
