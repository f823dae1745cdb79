set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/prof_all.py 55296 2 > gpurun_out/plain_all.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"^k_" -o gpurun_out/prof_r3b_all -f python scripts/prof_all.py 55296 1 > gpurun_out/ncu_r3b.log 2>&1
tail -2 gpurun_out/ncu_r3b.log
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/plain_bench.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r3b_bench_steps2.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra-configs > gpurun_out/ncu_launches.log 2>&1
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3b_f09_final.json 2> gpurun_out/bench_r3b_f09_final.err; tail -c 300 gpurun_out/bench_r3b_f09_final.err
python bench.py --steps 10 --warmup 3 --convtran 41 > gpurun_out/bench_r3b_config4_convtran41.json 2> gpurun_out/bench_r3b_config4.err
python bench.py --steps 10 --warmup 3 --ncols 13824 --no-extra-configs > gpurun_out/bench_r3b_config2_f19.json 2> gpurun_out/bench_r3b_config2.err
python bench.py --steps 10 --warmup 3 --ncols 131072 --pver 58 --parcel-pbl --pconv 0.4 --no-extra-configs > gpurun_out/bench_r3b_config5_shard_L58.json 2> gpurun_out/bench_r3b_config5.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r3b_reference.json 2>/dev/null
python scripts/microbench.py > gpurun_out/microbench_r3b_final.txt 2>&1
for f in gpurun_out/bench_r3b_f09_final.json gpurun_out/bench_r3b_config4_convtran41.json gpurun_out/bench_r3b_config2_f19.json gpurun_out/bench_r3b_config5_shard_L58.json; do python -c "
import json,sys
d=json.load(open('$f')); print('$f', round(d['ms_per_step'],3), round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d.get('cpu_baseline',{}).get('value'))"; done
