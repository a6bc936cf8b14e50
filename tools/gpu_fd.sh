mkdir -p gpurun_out/r01z
timeout 300 python -m pytest tests -m gpu -q -x -k "fused" > gpurun_out/r01z/pytest_fused.log 2>&1; tail -3 gpurun_out/r01z/pytest_fused.log
timeout 120 python tools/fused_debug.py 1024 1536 4 > gpurun_out/r01z/fused_big.log 2>&1; echo "exit $?"; head -22 gpurun_out/r01z/fused_big.log | cut -c1-330
