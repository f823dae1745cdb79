mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:"k_conv_evap|k_momtran_t|k_plume_w" -s 3 -c 3 -o gpurun_out/prof_r2l -f python scripts/prof_all.py 55296 2 > gpurun_out/ncu_r2l.log 2>&1
tail -n 2 gpurun_out/ncu_r2l.log
