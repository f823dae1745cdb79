mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r3i.json 2> gpurun_out/bench_r3i.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r3i.json'));print(d['ms_per_step'],d['roofline']['kernel_ms'],{k:(v['ms_per_step']) for k,v in d.get('configs').items()})"
