set -x
mkdir -p gpurun_out
python bench.py --steps 10 --warmup 3 > gpurun_out/bench_r2c_default.json 2> gpurun_out/bench_r2c_default.err; tail -c 400 gpurun_out/bench_r2c_default.err
python bench.py --steps 6 --warmup 3 --convtran 41 --no-cpu-baseline > gpurun_out/bench_r2c_convtran.json 2> gpurun_out/bench_r2c_convtran.err; tail -c 400 gpurun_out/bench_r2c_convtran.err
python - <<'PY'
import json
for f in ("gpurun_out/bench_r2c_default.json","gpurun_out/bench_r2c_convtran.json"):
    try:
        d=json.load(open(f)); print(f, d["ms_per_step"], d["e2e"]["value"], d["roofline"]["kernel_ms"]); print({k:(v["ms_per_step"],v["kernel_ms"]) for k,v in d.get("configs",{}).items()})
        print(d.get("cpu_baseline"))
    except Exception as e: print(f, "ERR", e)
PY
