#!/bin/bash
# --set full captures of the strict bank kernel: the product build and every variant under sdirt_b200/_lib/variants/
TAG=${1:-r02n}; OUT=gpurun_out/$TAG; mkdir -p $OUT
SC_MODES=strict timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank_run -s 1 -c 1 -o $OUT/prof_bank_strict -f python tools/strict_check.py rf50mm 592 262144 > $OUT/ncu_full_strict.log 2>&1; echo "ncu full exit $?"
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so)
SDIRT_ENGINE_LIB=$v SC_MODES=strict timeout 1200 ncu --set full --clock-control none --import-source on -k regex:psf_bank_run -s 1 -c 1 -o $OUT/prof_bank_strict_$n -f python tools/strict_check.py rf50mm 592 262144 > $OUT/ncu_full_strict_$n.log 2>&1; echo "ncu $n exit $?"; done
