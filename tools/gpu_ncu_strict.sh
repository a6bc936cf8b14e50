#!/bin/bash
TAG=${1:-r02g}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SC_MODES=strict timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank_run -s 1 -c 1 -o $OUT/prof_bank_strict -f python tools/strict_check.py rf50mm 592 262144 > $OUT/ncu_full_strict.log 2>&1; echo "ncu full exit $?"
