#!/bin/bash
# Multi-GPU visit (round 2): the contract bench under torchrun on N GPUs of one box, both arms, and a summary of the new legs.
N=${1:-2}; TAG=${2:-multi}; OUT=gpurun_out/$TAG; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.max.sm --format=csv > $OUT/smi.txt 2>&1
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,GRAPH timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 4 --warmup 3 > $OUT/bench_n$N.json 2> $OUT/bench_n$N.err
echo "bench N=$N exit $?"; grep -E "via |NVLS|Channel 00|P2P|SHM|NET/" $OUT/bench_n$N.err | head -12 | cut -c1-200
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench_n$N.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "clocks", d["clocks"])
    print("conformant", d["conformant"] and (d["conformant"]["value"], d["conformant"]["e2e"]["value"]))
    print("rf35mm", d["rf35mm"] and (d["rf35mm"]["value"], d["rf35mm"]["strict_rays_per_s"]))
    print("strong", {k: v for k, v in d["strong"].items() if k != "metric"})
    print("render_sharded", d["render_sharded"]["value"], d["render_sharded"]["ms"]); print("datagen", d["datagen"]["value"], d["datagen"].get("with_dfdp_forward"))
except Exception as e: print("parse failed", e)
PY
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $OUT/bench_ref_n$N.json 2> $OUT/bench_ref_n$N.err
echo "reference arm N=$N exit $?"; cut -c1-200 $OUT/bench_ref_n$N.json
