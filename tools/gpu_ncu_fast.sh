#!/bin/bash
TAG=${1:-r02E}; OUT=gpurun_out/$TAG; mkdir -p $OUT
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank_run -s 1 -c 1 -o $OUT/prof_bank_fast -f python bench.py --numerics fast --steps 1 --warmup 1 --quick > $OUT/ncu_full_fast.log 2>&1; echo "ncu full exit $?"
