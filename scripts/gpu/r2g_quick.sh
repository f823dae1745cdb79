mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
python scripts/prof_convr.py 55296 32 2 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:"k_buoyan_dilute" -s 2 -c 3 -o gpurun_out/prof_r2e_cape -f python scripts/prof_convr.py 55296 32 2 > gpurun_out/ncu_r2e.log 2>&1
tail -1 gpurun_out/ncu_r2e.log
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_f09_final.json 2> gpurun_out/bench_r2_f09_final.err; tail -c 300 gpurun_out/bench_r2_f09_final.err
python -c "
import json;d=json.load(open('gpurun_out/bench_r2_f09_final.json'));print(d['ms_per_step'],d['value'],d['e2e']['value'],d['roofline']['frac'],d['roofline']['kernel_ms'])"
