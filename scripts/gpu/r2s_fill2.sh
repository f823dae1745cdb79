mkdir -p gpurun_out
ZM_FILL_SIDE=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k error_conventions 2>&1 | grep -E "^E|Error|rc=" | head -20
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in 0 2 1; do
ZM_FILL_SIDE=$m python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2s_$m.json 2> gpurun_out/bench_r2s_$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2s_$m.json'));print('MODE $m',d['ms_per_step'],d['e2e']['value'],{k:(v['ms_per_step']) for k,v in d.get('configs').items()})"
done
