#!/bin/bash
# Round-2 evidence visit (strip-walking render kernel): GPU tests, contract bench (both arms), launch list of the bench command, --set full capture of the render kernel
TAG=${1:-r02P}; OUT=gpurun_out/$TAG; mkdir -p $OUT
bash tools/gpu_tests.sh $TAG
echo "== bench"; timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
timeout 900 python bench.py --impl reference > $OUT/bench_reference_arm.json 2> $OUT/bench_reference_arm.err; echo "reference arm exit $?"; cut -c1-260 $OUT/bench_reference_arm.json
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
    print("roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"], "eager", {k: v for k, v in (d["eager_gpu_baseline"] or {}).items() if "per_s" in k or "error" in k})
    print("conformant", d["conformant"] and (d["conformant"]["value"], d["conformant"]["e2e"]["value"], d["conformant"]["roofline"]["frac"]))
    print("rf35mm", d["rf35mm"] and (d["rf35mm"]["value"], d["rf35mm"]["strict_rays_per_s"]))
    print("strong", d["strong"]["value"], "render_sharded", d["render_sharded"]["value"], "datagen", d["datagen"]["value"], d["datagen"]["with_dfdp_forward"])
    print("modes", d["numerics_modes_rays_per_s"])
    print("render", d["render"]["value"], d["render"]["roofline"]["frac"], "psfnet", d["render_psfnet"]["value"], d["render_psfnet"]["roofline"]["frac"])
except Exception as e: print("parse failed", e)
PY
echo "== launch list"; timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/bench_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $OUT/bench_under_ncu.json 2> $OUT/bench_under_ncu.err; echo "ncu list exit $?"
python tools/launch_summary.py $OUT/bench_launches.csv > $OUT/bench_launches_summary.txt 2>&1; head -30 $OUT/bench_launches_summary.txt
echo "== ncu full, render kernel (2 x 3 x 1024 x 1536, ks = 21, tone curves on)"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:render_lanes -s 3 -c 1 -o $OUT/prof_render -f python tools/render_bench.py > $OUT/ncu_full_render.log 2>&1; echo "ncu full exit $?"
timeout 120 python tools/render_bench.py 2>&1 | grep float16 > $OUT/render_bench.txt; cat $OUT/render_bench.txt
