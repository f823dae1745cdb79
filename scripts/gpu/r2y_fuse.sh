mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
for m in 1 0 1 0; do
ZM_TEND_FUSE_EVAP=$m python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_r2y_$m.json 2> gpurun_out/bench_r2y_$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2y_$m.json'));print('FUSE_EVAP $m',d['ms_per_step'],d['e2e']['value'],d['roofline']['kernel_ms'])"
done
