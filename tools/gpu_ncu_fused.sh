mkdir -p gpurun_out/r01z
NCTA=2 timeout 300 ncu --set full --import-source on --clock-control none -k regex:mlp_fused_pred -c 1 -f -o gpurun_out/r01z/prof_fused3 python tools/fused_debug.py 1024 1536 4 > gpurun_out/r01z/ncu_fused3.log 2>&1; echo "exit $?"; tail -3 gpurun_out/r01z/ncu_fused3.log
