#!/bin/bash
# Scaling check on one box: the contract bench at N = 2 and N = 4 (own arm, no CPU leg).  Usage: bash tools/gpu_scale.sh tag
TAG=${1:-scale}; OUT=gpurun_out/$TAG; mkdir -p $OUT
for N in 2 4; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N bench.py --gpus $N --steps 4 --warmup 3 --no-cpu > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
  echo "bench N=$N exit $? lines $(wc -l < $OUT/bench_n$N.json)"; python -c "
import json;d=json.loads(open('$OUT/bench_n$N.json').read());print(d['n_gpus'],d['value'],d['e2e']['value'],d['ms_per_step'],d['clocks'])"
done
