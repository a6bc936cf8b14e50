#!/bin/bash
# Multi-GPU visit: the contract bench under torchrun on N GPUs of one box, both arms.  Usage: bash tools/gpu_multi.sh N [tag]
N=${1:-2}
TAG=${2:-multi}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
nvidia-smi topo -m > $OUT/topo.txt 2>&1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N exit $?"; cat $OUT/bench_n$N.json; tail -5 $OUT/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "reference arm N=$N exit $?"; cat $OUT/bench_ref_n$N.json
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 tools/gather_check.py > $OUT/gather_check.log 2>&1
echo "gather check exit $?"; tail -3 $OUT/gather_check.log
