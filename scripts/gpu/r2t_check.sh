mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2t.json 2> gpurun_out/bench_r2t.err; tail -c 300 gpurun_out/bench_r2t.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2t.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms'], {k:(v['ms_per_step']) for k,v in d.get('configs').items()})"
