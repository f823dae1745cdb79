"""Measures the FP64 (non-tensor) FMA peak of the current B200 twice -- the library's own microbenchmark
(zm_fp64_peak_flops) and an independent standalone probe (scripts/probes/fp64_peak_probe.cu) -- samples the clocks
during the runs and writes profiles/fp64_peak.json (the denominator bench.py's FP64 roofline uses)."""
import json, os, subprocess, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import zm_conv as Z
samples, stop = [], threading.Event()
def sampler():
    while not stop.is_set():
        o = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active",
                            "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True).stdout.strip()
        samples.append(o); stop.wait(0.1)
t = threading.Thread(target=sampler); t.start()
lib = [Z.fp64_peak_flops(40000) / 1e12 for _ in range(5)]
exe = "/tmp/fp64_peak_probe"
subprocess.run(["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", exe,
                os.path.join(ROOT, "scripts", "probes", "fp64_peak_probe.cu")], check=True)
probe = [json.loads(subprocess.run([exe], capture_output=True, text=True).stdout) for _ in range(3)]
stop.set(); t.join()
name = subprocess.run(["nvidia-smi", "--query-gpu=name", "--format=csv,noheader", "-i", "0"], capture_output=True, text=True).stdout.strip()
clk = [float(s.split(",")[0]) for s in samples if s]
d = {"gpu": name, "fp64_tflops": max(max(lib), max(p["fp64_tflops_probe"] for p in probe)),
     "library_microbenchmark_tflops": lib, "independent_probe_tflops": [p["fp64_tflops_probe"] for p in probe],
     "nominal": "148 SMs x 64 FP64 FMA lanes x 2 flop x 1.965 GHz = 37.2 TFLOP/s",
     "clocks": {"sm_mhz_median": sorted(clk)[len(clk) // 2] if clk else None, "sm_mhz_max": max(clk) if clk else None,
                "samples": samples[:40]},
     "when": time.strftime("%Y-%m-%dT%H:%M:%SZ", time.gmtime()),
     "how": "scripts/freeze_fp64_peak.py: best of the library's 8-chain FMA microbenchmark and a standalone 12-chain probe"}
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(d, open(os.path.join(ROOT, "gpurun_out", "fp64_peak.json"), "w"), indent=1)
print(json.dumps({k: v for k, v in d.items() if k != "clocks"}, indent=1), d["clocks"]["sm_mhz_median"])
