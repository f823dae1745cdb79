mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in 1 0 1 0; do
ZM_ORDER_SIDE=$m python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2u_$m.json 2> gpurun_out/bench_r2u_$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2u_$m.json'));print('ORDER_SIDE $m',d['ms_per_step'],d['e2e']['value'],{k:(v['ms_per_step']) for k,v in d.get('configs').items()})"
done
