mkdir -p gpurun_out/r01L
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r01L/pytest_gpu.log 2>&1; tail -2 gpurun_out/r01L/pytest_gpu.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/r01L/bench.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r01L/bench.json').read());print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['numerics_modes_rays_per_s'],d['ms_per_step'])"
