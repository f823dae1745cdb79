#!/usr/bin/env python
"""bench.py -- ZM deep-convection throughput (columns/s) on 1..N B200s.

Metric (BASELINE.json): "ZM convection columns/sec (f09 L32, r8)".  One step = one pass of the hot
path -- zm_conv_tend = zm_convr + physics_update + zm_conv_evap + momtran (BASELINE config 3;
reference physics/zm_conv_intr.F90:662-836) -- over one rank's synthetic f09 shard (55,296 columns,
pver=32, pcols=16, r8).  Columns shard with no hot-path communication: every rank owns its own
55,296-column shard (weak scaling); the only collective is one NCCL all-reduce of six doubles per
step for the global water/energy budget check.

  value  : whole-job columns/s with physics_state resident in HBM (CUDA events, max over ranks)
  e2e    : the same step through the host-pointer C ABI (zm_conv_tend_batch) from pinned host
           buffers, H2D + kernels + D2H inside the timed region
  roofline: dominant kernel k_buoyan_dilute<1> (FP64-pipe bound), timed live with CUDA events
  cpu_baseline / --impl reference: the CPU oracle port of the reference (OpenMP over chunks,
           glibc libm) timed on this box's host cores.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "ZM convection columns/sec (f09 L32, r8)"
UNIT = "columns/s"
NCOLS_F09 = 55296

# executed FP64 flops per column of one k_buoyan_dilute<1> launch on this workload (FMA = 2), from
# the ncu capture committed under profiles/ (dfma*2 + dadd + dmul thread-level counts / columns).
FLOPS_FILE = os.path.join(ROOT, "profiles", "flops_per_column.json")


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0}, "fallback"


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    def __init__(self, index: int):
        self.index = index
        self.samples = []
        self.reasons = set()
        self._stop = threading.Event()
        self._t = None

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                      "-i", str(self.index)], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append((float(parts[0]), float(parts[1])))
                    for n, v in zip(names, parts[2:6]):
                        if v.lower().startswith("active"):
                            self.reasons.add(n)
            except Exception:
                pass
            self._stop.wait(0.1)

    def start(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def stop(self):
        self._stop.set()
        if self._t:
            self._t.join(timeout=10)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": statistics.median(s[0] for s in self.samples),
                "sm_max_mhz": max(s[1] for s in self.samples), "reasons": sorted(self.reasons),
                "samples": len(self.samples)}


def cpu_reference_rate(ncols, pver, pconv, reps, nthreads=0, parcel_pbl=False):
    """Times the CPU port of the reference (oracle, glibc libm flavour, OpenMP over chunks)."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from oracle_lib import Oracle
    from cam_nor_physics_b200 import soundings as S
    o = Oracle("libm")
    op = o.default_params(16, pver, S.limcnv_for(pver))
    op.lparcel_pbl = int(parcel_pbl)
    o.convi(op)
    ch = S.make_chunks(ncols, pver, 16, p_conv=pconv)
    cores = nthreads or (os.cpu_count() or 1)
    o.conv_tend_batch(ch, nthreads=cores)          # warm-up
    best = 1e30
    for _ in range(reps):
        t0 = time.perf_counter()
        r = o.conv_tend_batch(ch, nthreads=cores)
        best = min(best, time.perf_counter() - t0)
        if r["rc"]:
            raise RuntimeError("oracle Brent failure")
    return ncols / best, cores, best, o.backend()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ncols", type=int, default=NCOLS_F09, help="columns per GPU")
    ap.add_argument("--pver", type=int, default=32)
    ap.add_argument("--pconv", type=float, default=0.35, help="convective fraction of the synthetic grid")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--parcel-pbl", action="store_true", help="zmconv_parcel_pbl=.true. (CAM6 L58 default)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries the JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints its
    # version banner there) are sent to stderr for the whole run, the line goes to the saved descriptor
    sys.stdout.flush()
    json_out = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.ncols == NCOLS_F09 and args.pver == 32:
        grid = f"f09 FV grid shard: {args.ncols} columns x L32 per GPU (BASELINE config 3)"
    else:
        grid = f"{args.ncols} columns x L{args.pver} per GPU (non-default size; BASELINE config 5 uses 131072 x L58 per GPU)"
    workload = grid + ", pcols=16, zm_conv_tend = zm_convr+physics_update+zm_conv_evap+momtran"
    step_gb = args.ncols * ((26 * args.pver + 15) * 8 + 12 + (14 * args.pver + 6) * 8 + 19 * args.pver * 8 + 12) / 1e9
    config = {"workload": workload, "columns_per_gpu": args.ncols, "pver": args.pver, "pcols": 16,
              "convective_fraction_target": args.pconv, "parcel_pbl": bool(args.parcel_pbl), "seed": 20261018, "parallelism": f"columns x{world}",
              "l2": f"per-step algorithmic inputs+outputs ({step_gb:.2f} GB) exceed the 126 MB L2; no explicit flush"}

    # ---------------- reference arm: CPU port of the reference on host cores ---------------------
    if args.impl == "reference":
        if rank != 0:
            return
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        from oracle_lib import Oracle
        from cam_nor_physics_b200 import soundings as S
        o = Oracle("libm")
        op = o.default_params(16, args.pver, S.limcnv_for(args.pver))
        op.lparcel_pbl = int(args.parcel_pbl)
        o.convi(op)
        ch = S.make_chunks(args.ncols, args.pver, 16, p_conv=args.pconv)
        cores = os.cpu_count() or 1
        backend = o.backend()
        times = []
        for i in range(args.warmup + args.steps):
            t0 = time.perf_counter()
            r = o.conv_tend_batch(ch, nthreads=cores)
            dt = time.perf_counter() - t0
            if r["rc"]:
                raise RuntimeError("oracle Brent failure")
            if i >= args.warmup:
                times.append(dt)
        t = statistics.mean(times)
        rate = args.ncols / t
        line = {"impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config,
                "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"full step on {args.ncols} columns, CPU oracle ({backend}), "
                                           "OpenMP over pcols=16 chunks"},
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        json_out.write(json.dumps(line) + "\n"); json_out.flush()
        return

    # ---------------- B200 arm ---------------------------------------------------------------------
    import numpy as np
    import torch
    import torch.distributed as dist
    from cam_nor_physics_b200 import build, soundings as S, zm_conv as Z
    from cam_nor_physics_b200.device import DeviceTend

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    build.build()
    L = args.pver
    zp = Z.default_params(16, L, S.limcnv_for(L))
    zp.lparcel_pbl = int(args.parcel_pbl)
    Z.zm_init(zp)

    ch = S.make_chunks(args.ncols, L, 16, p_conv=args.pconv, col0=rank * args.ncols)
    dev = DeviceTend(ch)
    stream = torch.cuda.current_stream()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def one_step():
        dev.step()
        cons = dev.conservation()
        if world > 1:
            dist.all_reduce(cons)
        return cons

    for _ in range(args.warmup):
        one_step()
    dev.check()
    barrier()
    Z.lib().zm_launch_count(1)
    sampler = ClockSampler(local_rank)
    sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(args.steps):
        cons = one_step()
    e1.record(stream)
    barrier()
    launches = int(Z.lib().zm_launch_count(1))
    nfail = dev.check()
    ms = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms_total = float(ms.item())
    ms_per_step = ms_total / args.steps
    value = world * args.ncols / (ms_per_step * 1e-3)
    cons_h = cons.cpu().numpy()

    # ---- per-kernel times (profiling events between kernels; separate loop, not the timed one) -----
    Z.lib().zm_set_profiling(1)
    ktimes = {}
    nprof = min(args.steps, 10)
    for _ in range(nprof):
        dev.step()
        torch.cuda.synchronize()
        for n, t in Z.kernel_times():
            ktimes.setdefault(n, []).append(t)
    Z.lib().zm_set_profiling(0)
    kavg = {n: statistics.mean(v) for n, v in ktimes.items()}
    dom = "buoyan_dilute_pass1"
    t_dom = kavg.get(dom, float("nan")) * 1e-3
    peaks, peak_kind = _peaks()
    try:
        with open(FLOPS_FILE) as f:
            fl = json.load(f)
        flops_per_col = float(fl["buoyan_dilute_pass1_flops_per_column"])
    except Exception:
        fl, flops_per_col = {}, float("nan")
    fp64_peak = Z.fp64_peak_flops(20000)
    # the committed flop count was captured on the default workload; it does not transfer to other level
    # counts / launch parcels, so the FP64 fraction is only reported there
    flops_valid = (args.ncols == int(fl.get("ncols", -1)) and L == 32 and not args.parcel_pbl
                   and abs(args.pconv - 0.35) < 1e-12)
    if not flops_valid:
        flops_per_col = float("nan")
    achieved = flops_per_col * args.ncols / t_dom / 1e12
    alg_bytes = (6 * L + 2 + 2 * L + 6) * 8 + 12      # inputs t,q,pap,zm (L) + paph,zi (L+1) + 3; outputs tp,qstp + 6 scalars
    roofline = {"kernel": "k_buoyan_dilute<1> (dilute CAPE trigger, pass 1, all columns)",
                "bound": "fp64", "achieved": achieved if flops_valid else None, "peak": fp64_peak / 1e12,
                "unit": "TFLOP/s", "frac": achieved / (fp64_peak / 1e12) if flops_valid else None,
                "peak_source": "FP64 FMA-chain microbenchmark run live in this process (MEASURED_PEAKS.json has no "
                               "FP64 entry; SURVEY.md section 6 asks the builder to measure it)",
                "flops_per_column": flops_per_col if flops_valid else None, "flops_source": fl.get("source"),
                "ms_per_launch": t_dom * 1e3,
                "traffic": fl.get("dram_bytes_per_launch") if flops_valid else None,
                "hbm": {"achieved_gbs": alg_bytes * args.ncols / t_dom / 1e9, "peak_gbs": peaks.get("hbm_gbs"),
                        "peak_kind": peak_kind, "alg_bytes_per_column": alg_bytes},
                "kernel_ms": kavg}

    # ---- e2e: host-pointer C ABI from pinned host buffers -----------------------------------------
    st = {k: v.numpy() for k, v in dev.host_in.items()}     # pinned host memory views
    nch, pc = ch.nchunks, 16
    out = {}
    for k in Z.TEND_OUT_2D:
        out[k] = torch.zeros((nch, L, pc), dtype=torch.float64).pin_memory().numpy()
    for k in Z.TEND_OUT_2DP:
        out[k] = torch.zeros((nch, L + 1, pc), dtype=torch.float64).pin_memory().numpy()
    for k in Z.TEND_OUT_1D:
        out[k] = torch.zeros((nch, pc), dtype=torch.float64).pin_memory().numpy()
    for k in Z.TEND_OUT_INT:
        out[k] = torch.zeros((nch, pc), dtype=torch.int32).pin_memory().numpy()
    out["lengath"] = torch.zeros(nch, dtype=torch.int32).pin_memory().numpy()
    h2d = sum(v.nbytes for v in st.values()) + ch.ncol.nbytes
    d2h = sum(v.nbytes for v in out.values())
    for _ in range(2):
        Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
    barrier()
    t0 = time.perf_counter()
    nsteps_e2e = max(3, min(args.steps, 10))
    for _ in range(nsteps_e2e):
        Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)     # synchronous: returns after the D2H
    torch.cuda.synchronize()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * args.ncols * nsteps_e2e / float(te.item())
    # same call with the pbuf mass-flux fields (ZM_MU..ZM_MAXG) left in the library's device mirror for
    # zm_conv_tend_2_batch instead of being copied back -- informational, the headline e2e copies everything
    mirror_keys = ("mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt", "maxg")
    d2h_res = d2h - sum(out[k].nbytes for k in mirror_keys)
    Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out, keep_pbuf_on_device=True)
    barrier()
    t0 = time.perf_counter()
    for _ in range(nsteps_e2e):
        Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out, keep_pbuf_on_device=True)
    torch.cuda.synchronize()
    tr = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tr, op=dist.ReduceOp.MAX)
    e2e_resident = world * args.ncols * nsteps_e2e / float(tr.item())
    clocks = sampler.stop()       # sampled through the timed loop, the per-kernel loop and the e2e loop

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": config, "clocks": clocks,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                        "d2h_bytes_per_step": int(d2h), "steps": nsteps_e2e,
                        "api": "zm_conv_tend_batch (host pointers, pinned)",
                        "pbuf_resident": {"value": e2e_resident, "d2h_bytes_per_step": int(d2h_res),
                                          "note": "mu,md,du,eu,ed,dp,dsubcld,jt,maxg kept in the device mirror"}},
                "gpu_launches": launches, "roofline": roofline,
                "convective_columns_per_gpu": int(dev.out["lengath"].sum().item()),
                "brent_failures": int(nfail),
                "conservation": {"sum_pdel_g_ptend_q": cons_h[0], "sum_1000_prec_plus_rliq": cons_h[1],
                                 "water_residual_rel": float((cons_h[0] + cons_h[1]) / max(abs(cons_h[1]), 1e-300)),
                                 "sum_pdel_g_ptend_s": cons_h[2], "sum_latent": cons_h[3],
                                 "convective_columns": cons_h[4], "columns": cons_h[5]}}
        if world == 1 and not args.no_cpu_baseline:
            rate, cores, best, backend = cpu_reference_rate(args.ncols, L, args.pconv, 3, parcel_pbl=args.parcel_pbl)
            line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"full step on all {args.ncols} columns x3 (best), CPU oracle "
                                              f"({backend}) with OpenMP over pcols=16 chunks; {best*1e3:.1f} ms/step"}
        json_out.write(json.dumps(line) + "\n"); json_out.flush()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
