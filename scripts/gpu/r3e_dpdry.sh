mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 10 --warmup 3 --convtran 41 > gpurun_out/bench_r3e_config4_convtran41.json 2> gpurun_out/bench_r3e_config4.err
python -c "
import json
d=json.load(open('gpurun_out/bench_r3e_config4_convtran41.json')); print('config4', round(d['ms_per_step'],3), round(d['value']/1e6,2), round(d['e2e']['value']/1e6,2), d['roofline']['kernel_ms'])"
