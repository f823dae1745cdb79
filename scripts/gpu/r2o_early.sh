mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2o.json 2> gpurun_out/bench_r2o.err; tail -c 300 gpurun_out/bench_r2o.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2o.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms'], {k:(v['ms_per_step'],v['kernel_ms']['buoyan_dilute_pass1'],v['kernel_ms']['buoyan_dilute_pass2']) for k,v in d.get('configs').items()})"
ZM_CAPE_EARLY_EXIT=0 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2o_off.json 2> gpurun_out/bench_r2o_off.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2o_off.json'));print('OFF',d['ms_per_step'],d['roofline']['kernel_ms']['buoyan_dilute_pass2'])"
