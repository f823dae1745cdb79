"""Writes profiles/flops_per_column.json: FP64 flops per column of the first CAPE pass (k_buoyan_dilute<1>, the
dominant kernel) on the default bench workload (f09 shard, 55,296 columns L32, p_conv 0.35, seed 20261018).

  algorithmic : the CPU oracle's instrumented operation count of its first buoyan_dilute call on these soundings
                (oracle/zm_oracle.cpp FL(n) counters; SURVEY.md 8d: + - * / compare = 1 each; log, log10, 10**x, exp,
                x**y counted as calls and expanded at the FP64 cost of the CUDA libm routine, measured from the SASS
                of log / log10 / pow(10,x) / exp / pow for sm_100a: 2*DFMA + DADD + DMUL of the routine)
  executed    : 2*DFMA + DADD + DMUL thread-level instruction counts of the kernel from an ncu capture
                (scripts/ncu_summary.py JSON), with the hash of the kernel sources at capture time so that bench.py can
                tell when the kernel has changed since.

usage: python scripts/count_flops.py [profiles/<capture>.json]      (no argument: keep the executed part on file)"""
import hashlib, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from oracle_lib import Oracle
from cam_nor_physics_b200 import soundings as S

NCOLS, L, PCONV = 55296, 32, 0.35
# cuobjdump -sass of a one-line kernel per routine, nvcc 12.9 -gencode arch=compute_100a,code=sm_100a -O3:
# (DFMA, DADD, DMUL) = log (16, 9, 5), log10 (17, 9, 6), pow(10,x) (38, 35, 8), exp (14, 2, 1), pow (38, 36, 9)
LIBM_FLOPS = {"log": 46, "log10": 49, "pow10": 119, "exp": 31, "pow": 121}
OUT = os.path.join(ROOT, "profiles", "flops_per_column.json")


def kernel_hash():
    h = hashlib.sha256()
    for f in ("zm_kernels.cuh", "zm_device.cuh", "zm_math.h", "zm_svp_table.h"):
        h.update(open(os.path.join(ROOT, "cam_nor_physics_b200", "csrc", f), "rb").read())
    return h.hexdigest()[:16]


def main():
    o = Oracle("libm")
    o.convi(o.default_params(16, L, S.limcnv_for(L)))
    ch = S.make_chunks(NCOLS, L, 16, p_conv=PCONV)
    o.counters_reset()
    r = o.convr_batch(ch, nthreads=1)
    fl = o.flops()
    names = ["basic", "log", "log10", "pow10", "exp", "pow", "state_function_evaluations", "brent_iterations"]
    ph = {p: dict(zip(names, [int(v) for v in fl[p]])) for p in (1, 2)}
    def expanded(d):
        return d["basic"] + sum(d[k] * LIBM_FLOPS[k] for k in LIBM_FLOPS)
    try:
        old = json.load(open(OUT))
    except Exception:
        old = {}
    d = {"ncols": NCOLS, "pver": L, "pconv": PCONV, "seed": 20261018,
         "algorithmic_flops_per_column": expanded(ph[1]) / NCOLS,
         "algorithmic_basic_flops_per_column": ph[1]["basic"] / NCOLS,
         "algorithmic_calls_per_column": {k: ph[1][k] / NCOLS for k in names[1:]},
         "algorithmic_pass2_flops_per_column_all_columns": expanded(ph[2]) / NCOLS,
         "libm_flops_per_call": LIBM_FLOPS,
         "algorithmic_source": "scripts/count_flops.py: oracle (glibc-libm flavour) operation count of the first "
                               "buoyan_dilute call, transcendental calls expanded at the CUDA libm FP64 cost",
         "convective_columns": int(r["lengath"].sum())}
    for k in ("executed_flops_per_column", "executed_source", "dram_bytes_per_launch", "executed_kernel_hash"):
        if k in old:
            d[k] = old[k]
    if len(sys.argv) > 1:
        cap = json.load(open(sys.argv[1]))
        rows = cap if isinstance(cap, list) else cap.get("kernels", [])
        for row in rows:
            name = row.get("Kernel Name", {}).get("value", "")
            if "k_buoyan_dilute<1" in name or "k_buoyan_dilute<(int)1" in name:
                d["executed_flops_per_column"] = row["fp64_flops_per_launch"] / NCOLS
                scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}     # ncu picks the unit per value
                d["dram_bytes_per_launch"] = sum(row[m]["value"] * scale[row[m]["unit"]]
                                                 for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
                d["executed_source"] = f"{os.path.relpath(sys.argv[1], ROOT)} (ncu --set full): 2*dfma+dadd+dmul thread-level executed ops of {name.strip()}"
                d["executed_kernel_hash"] = kernel_hash()
    d["kernel_hash_now"] = kernel_hash()
    json.dump(d, open(OUT, "w"), indent=1)
    print(json.dumps(d, indent=1))


if __name__ == "__main__":
    main()
