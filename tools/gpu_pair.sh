mkdir -p gpurun_out/r01E
timeout 600 python -m pytest tests -m gpu -q -x -k "strict or hybrid or adaptive or psf" > gpurun_out/r01E/pytest_gpu.log 2>&1; tail -2 gpurun_out/r01E/pytest_gpu.log
QB_MODES=hybrid,adaptive,fast timeout 300 python tools/quick_bench.py rf50mm 1184 1048576 > gpurun_out/r01E/quick_rf50.log 2>&1; cat gpurun_out/r01E/quick_rf50.log
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r01E/bench.json 2>/dev/null; python -c "
import json;d=json.loads(open('gpurun_out/r01E/bench.json').read());print(d['value'],d['e2e']['value'],d['roofline']['frac'],d['numerics_modes_rays_per_s'],d['ms_per_step'])"
