mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in 1 2; do
python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_r3c_$m.json 2> gpurun_out/bench_r3c_$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r3c_$m.json'));print('RUN $m',d['ms_per_step'],d['e2e']['value'],{k:(v['ms_per_step']) for k,v in d.get('configs').items()}, d['conservation'])"
done
