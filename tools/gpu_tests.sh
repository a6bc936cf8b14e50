#!/bin/bash
TAG=${1:-r02u}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log
grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -25
grep -E "psf_map per-tile|get_training_data L1|get_test_data L1|rms \(avg|in-focus corner|adaptive vs strict|rf35mm (replay|strict|hybrid|adaptive|fast) 2M" $OUT/pytest_gpu.log | cut -c1-400
