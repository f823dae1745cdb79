mkdir -p gpurun_out
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"^k_" -o gpurun_out/prof_r2g_all -f python scripts/prof_all.py 55296 1 > gpurun_out/ncu_r2g.log 2>&1
tail -n 2 gpurun_out/plain_all.log gpurun_out/ncu_r2g.log
