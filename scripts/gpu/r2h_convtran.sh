mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "convtran or evap_momtran or conv_tend or reference_source_text or randomised or golden or device_mirror or other_level" 2>&1 | tail -5
python bench.py --convtran 41 --steps 10 --warmup 3 > gpurun_out/ct_new.json 2> gpurun_out/ct_new.err; tail -c 300 gpurun_out/ct_new.err
python - <<'PY'
import json
for n in ("new",):
    d=json.load(open(f"gpurun_out/ct_{n}.json")); print(n, d["ms_per_step"], d.get("kernel_ms") or d["roofline"].get("kernel_ms"), d.get("timers"))
PY
ncu --set full --clock-control none --import-source on -k regex:"k_convtran_c" -s 2 -c 2 -o gpurun_out/prof_r2h_ct -f python scripts/prof_all.py 55296 2 > gpurun_out/ncu_r2h.log 2>&1
tail -n 2 gpurun_out/ncu_r2h.log
