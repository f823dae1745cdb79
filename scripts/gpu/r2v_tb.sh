mkdir -p gpurun_out
for m in 128 64 32; do
ZM_CAPE_TB=$m python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extra-configs > gpurun_out/bench_r2v_$m.json 2> gpurun_out/bench_r2v_$m.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2v_$m.json'));print('TB $m',d['ms_per_step'],d['roofline']['kernel_ms']['buoyan_dilute_pass1'])"
done
