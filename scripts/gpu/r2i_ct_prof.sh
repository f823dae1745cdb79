mkdir -p gpurun_out
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_convtran_c" -s 2 -c 2 -o gpurun_out/prof_r2h_ct -f python scripts/prof_all.py 55296 2 > gpurun_out/ncu_r2h.log 2>&1
tail -2 gpurun_out/plain_all.log gpurun_out/ncu_r2h.log
