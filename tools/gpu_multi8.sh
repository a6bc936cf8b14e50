#!/bin/bash
# 8-GPU check of the contract bench (own arm only) + the bank gather.  Usage: bash tools/gpu_multi8.sh N tag
N=${1:-8}; TAG=${2:-multi8}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N exit $?"; cat $OUT/bench_n$N.json | cut -c1-1500; tail -3 $OUT/bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/gather_check.py > $OUT/gather_check.log 2>&1
echo "gather check exit $?"; tail -3 $OUT/gather_check.log
