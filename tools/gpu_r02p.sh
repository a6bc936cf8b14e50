#!/bin/bash
TAG=${1:-r02p}; OUT=gpurun_out/$TAG; mkdir -p $OUT
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests -m gpu -q -rP > $OUT/pytest_gpu.log 2>&1; echo "pytest exit $?" | tee -a $OUT/pytest_gpu.log; grep -E "passed|failed|^FAILED|^ERROR" $OUT/pytest_gpu.log | tail -15
echo "== strict_check"; SC_MODES=strict timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_base_rf50.log 2>&1; tail -2 $OUT/strict_base_rf50.log
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so); SDIRT_ENGINE_LIB=$v timeout 600 python tools/strict_check.py rf50mm 592 1048576 > $OUT/strict_${n}_rf50.log 2>&1; echo "variant $n"; tail -2 $OUT/strict_${n}_rf50.log; done
timeout 600 python bench.py --numerics strict --steps 6 --warmup 3 --quick 2> $OUT/q_strict.err | tee $OUT/q_strict.json
for v in $(ls sdirt_b200/_lib/variants/*.so 2>/dev/null); do n=$(basename $v .so); SDIRT_ENGINE_LIB=$v timeout 600 python bench.py --numerics strict --steps 6 --warmup 3 --quick 2> $OUT/q_strict_$n.err | tee $OUT/q_strict_$n.json; done
echo "== bench"; timeout 1500 python bench.py > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"; tail -c 1500 $OUT/bench.err | grep -E "Elapsed|Error|error" ; python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    for k in ("value", "ms_per_step"): print(k, d[k])
    print("e2e", d["e2e"]["value"], "launches", d["gpu_launches"], "clocks", d["clocks"])
    print("roofline", d["roofline"]["frac"], "cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], "eager", d["eager_gpu_baseline"])
    print("conformant", d["conformant"] and (d["conformant"]["value"], d["conformant"]["e2e"]))
    print("rf35mm", d["rf35mm"] and (d["rf35mm"]["value"], d["rf35mm"]["strict_rays_per_s"], [(s["focus_mm"], round(s["d_sensor"], 3), s["rays_per_s"]) for s in d["rf35mm"]["focus_sweep"]]))
    print("strong", d["strong"]); print("render_sharded", d["render_sharded"]); print("datagen", d["datagen"]); print("modes", d["numerics_modes_rays_per_s"])
    print("render", d["render"]["value"], d["render"]["roofline"]["frac"], "psfnet", d["render_psfnet"]["value"], d["render_psfnet"]["roofline"]["frac"])
except Exception as e: print("parse failed", e)
PY
