mkdir -p gpurun_out/r01E
QB_MODES=hybrid,adaptive,fast timeout 300 python tools/quick_bench.py rf50mm 1184 1048576 > gpurun_out/r01E/quick_rf50.log 2>&1; cat gpurun_out/r01E/quick_rf50.log
QB_MODES=hybrid,adaptive,fast timeout 300 python tools/quick_bench.py rf35mm 1184 1048576 > gpurun_out/r01E/quick_rf35.log 2>&1; cat gpurun_out/r01E/quick_rf35.log
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r01E/pytest_gpu.log 2>&1; tail -15 gpurun_out/r01E/pytest_gpu.log
