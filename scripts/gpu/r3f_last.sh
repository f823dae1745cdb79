mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -2
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r3f.json 2> gpurun_out/bench_r3f.err; tail -c 200 gpurun_out/bench_r3f.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r3f.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['executed_flops_stale'],{k:(v['ms_per_step']) for k,v in d.get('configs').items()})"
