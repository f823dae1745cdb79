mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_conservation|k_add_seten|k_conv_evap|k_momtran_t|k_cldprp_pass1" -o gpurun_out/prof_r3d_tail -f python scripts/prof_all.py 55296 1 > gpurun_out/ncu_r3d.log 2>&1
tail -1 gpurun_out/ncu_r3d.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r3d_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3d_f09_final.json 2> gpurun_out/bench_r3d_f09_final.err; tail -c 300 gpurun_out/bench_r3d_f09_final.err
python bench.py --steps 10 --warmup 3 --convtran 41 > gpurun_out/bench_r3d_config4_convtran41.json 2> gpurun_out/bench_r3d_config4.err
python bench.py --steps 10 --warmup 3 --ncols 13824 --no-extra-configs > gpurun_out/bench_r3d_config2_f19.json 2> gpurun_out/bench_r3d_config2.err
python bench.py --steps 10 --warmup 3 --ncols 131072 --pver 58 --parcel-pbl --pconv 0.4 --no-extra-configs > gpurun_out/bench_r3d_config5_shard_L58.json 2> gpurun_out/bench_r3d_config5.err
for f in f09_final config4_convtran41 config2_f19 config5_shard_L58; do python -c "
import json
d=json.load(open('gpurun_out/bench_r3d_$f.json')); print('$f', round(d['ms_per_step'],3), round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d.get('cpu_baseline',{}).get('value'))"; done
