"""Randomised parity sweep: CUDA zm_convr / zm_conv_tend (with and without zm_org) / convtran vs the CPU oracle (portable-math flavour, bit-exact) over random seeds,
batch sizes, chunk widths, level counts, convective fractions and namelist options.  Runs on the GPU box
(`gpurun -- python scripts/parity_fuzz.py [ncases] [seed]`); the oracle is the checker, never the thing measured."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cam_nor_physics_b200 import soundings as S
from helpers import get_oracle, init_cuda, cuda_convr, assert_same, state_of, dpdry_gathered, CONVR_KEYS, TEND_KEYS

ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 24
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 7)
OPTS = [{}, {}, {"num_cin": 3}, {"no_deep_pbl": 1}, {"masterproc": 0, "dmpdz": -0.5e-3}, {"lparcel_pbl": 1},
        {"lparcel_pbl": 1, "num_cin": 5}, {"cam3": 1, "num_cin": 5}, {"capelmt": 200.0}, {"tau": 7200.0, "c0_ocn": 0.01}]
report = []
t00 = time.time()
for case in range(ncases):
    pver = int(rng.choice([24, 26, 32, 32, 32, 58, 72]))
    pcols = int(rng.choice([8, 16, 16, 16, 24, 48, 128]))
    ncols = int(rng.integers(1, 6000))
    for k in ("ZM_TEND_SUBBATCHES", "ZM_TEND_SCHEDULE"):
        os.environ.pop(k, None)
    if rng.random() < 0.12:                   # a few large batches: the pipelined host API cuts them into sub-batches
        ncols = int(rng.integers(16000, 60000))
        mode = int(rng.integers(0, 4))        # default ramp / equal parts / explicit sixteenths / unpipelined
        if mode == 1: os.environ["ZM_TEND_SUBBATCHES"] = str(int(rng.integers(2, 9)))
        if mode == 2: os.environ["ZM_TEND_SCHEDULE"] = str(rng.choice(["1,2,3,4,6", "2,2,4,8", "1,1,1,1,4,8", "8,8"]))
        if mode == 3: os.environ["ZM_TEND_SUBBATCHES"] = "1"
    pconv = float(rng.choice([0.0, 0.1, 0.35, 0.6, 1.0]))
    seed = int(rng.integers(1, 2**31))
    over = dict(OPTS[int(rng.integers(0, len(OPTS)))])
    Z = init_cuda(pcols, pver, **over)
    o, _, rc = get_oracle("pm", pcols, pver, **over)
    assert rc == 0, (rc, over)
    ch = S.make_chunks(ncols, pver, pcols, p_conv=pconv, seed=seed)
    if rng.random() < 0.3:                     # ragged chunks everywhere, not only at the end (CAM chunks differ in ncol)
        ch.ncol[...] = rng.integers(1, pcols + 1, ch.ncol.shape).astype(ch.ncol.dtype)
    perturbed = bool(rng.random() < 0.35)
    if perturbed:                             # level-wise noise: super-saturated / very dry layers, inversions, odd PBLs
        ch.t[...] = ch.t + rng.normal(0.0, 2.5, ch.t.shape)
        ch.q[...] = np.maximum(ch.q * rng.uniform(0.3, 1.7, ch.q.shape), 1e-12)
        ch.pblh[...] = rng.uniform(20.0, 4000.0, ch.pblh.shape)
        ch.tpert[...] = rng.uniform(0.0, 4.0, ch.tpert.shape)
        ch.landfrac[...] = rng.choice([0.0, 1.0, 0.37], size=ch.landfrac.shape)
    try:
        kind = case % 4
        use_org = kind == 3 and not over.get("cam3")
        if use_org:                               # organisation tracer branches (zm_org), values in [0, 1] with zeros
            over["zm_org"] = 1
            Z = init_cuda(pcols, pver, **over)
            o, _, rc = get_oracle("pm", pcols, pver, **over)
            org = np.maximum(rng.uniform(-0.3, 1.0, ch.t.shape), 0.0)
        if kind in (0, 2):                        # zm_convr alone
            ref = o.convr_batch(ch)
            out = cuda_convr(Z, ch)
            assert_same(out, ref, CONVR_KEYS, pcols, exact=True, what=f"fuzz case {case} (zm_convr)")
        elif use_org:
            st = state_of(ch); st["org"] = org
            ref = o.conv_tend_batch(ch, org=org)
            out = Z.zm_conv_tend(ch.ncol, st, ch.ztodt)
            assert_same(out, ref, TEND_KEYS + ["orgt", "org2d"], pcols, exact=True, what=f"fuzz case {case} (zm_conv_tend, zm_org)")
        else:                                     # the whole zm_conv_tend sequence
            ref = o.conv_tend_batch(ch)
            mirror = bool(rng.random() < 0.5)     # pbuf mass-flux fields stay on the device, zm_conv_tend_2 follows
            out = Z.zm_conv_tend(ch.ncol, state_of(ch), ch.ztodt, keep_pbuf_on_device=mirror)
            keys = [k for k in TEND_KEYS if k not in ("mu", "md", "du", "eu", "ed", "dp", "dsubcld", "jt", "maxg")] if mirror else TEND_KEYS
            assert_same(out, ref, keys, pcols, exact=True, what=f"fuzz case {case} (zm_conv_tend, mirror={mirror})")
            if mirror and ncols <= 3000:
                ncnst = int(rng.integers(2, 10))
                q, fracis, pdeldry = S.make_tracers(ch, ncnst)
                do = [0] + [int(x) for x in rng.integers(0, 2, ncnst - 1)]
                dry = [int(x) for x in rng.integers(0, 2, ncnst)]
                dq = Z.zm_conv_tend_2(do, q, pdeldry, fracis, ch.ztodt, dry)
                dpdry = dpdry_gathered(ch, ref, pdeldry)
                for c in range(ch.nchunks):
                    r = o.convtran(do, q[c], ref["mu"][c], ref["md"][c], ref["du"][c], ref["eu"][c], ref["ed"][c],
                                   ref["dp"][c], ref["dsubcld"][c], ref["jt"][c], ref["maxg"][c], ref["ideep"][c],
                                   ref["lengath"][c], fracis[c], dpdry[c], ch.ztodt, dry)
                    for m in range(ncnst):
                        assert np.array_equal(dq[c, m], r[m] if do[m] else np.zeros_like(r[m])), ("tend_2", case, c, m)
            if kind == 1 and ncols <= 2000:       # N4: geopotential_t (both branches) and convect_diagnostics_calc
                zvir = np.full_like(ch.t, S.ZVIR); rair = np.full_like(ch.t, S.RAIR)
                piln, rpdel = np.log(ch.pint), 1.0 / ch.pdel
                lr = bool(rng.random() < 0.5)
                zi, zm = Z.geopotential_t(ch.ncol, piln, np.log(ch.pmid), ch.pint, ch.pmid, ch.pdel, rpdel, ch.t, ch.q,
                                          rair, S.GRAVIT, zvir, dycore_lr=lr)
                dg = Z.convect_diagnostics_calc(ch.ncol, ref["mcon"], ref["dlf"], ref["rliq"], ch.pmid, ref["rprd"],
                                                ref["jctop"], ref["jcbot"])
                for c in range(ch.nchunks):
                    n = int(ch.ncol[c])
                    rzi, rzm = o.geopotential_t(n, lr, piln[c], ch.pint[c], ch.pmid[c], ch.pdel[c], rpdel[c], ch.t[c],
                                                ch.q[c], rair[c], S.GRAVIT, zvir[c])
                    assert np.array_equal(zi[c][:, :n], rzi[:, :n]) and np.array_equal(zm[c][:, :n], rzm[:, :n]), ("geopotential", case, c)
                    r = o.convect_diagnostics(n, ref["mcon"][c], ref["dlf"][c], ref["rliq"][c], ch.pmid[c], ref["rprd"][c],
                                              ref["jctop"][c], ref["jcbot"][c])
                    for k in r:
                        assert np.array_equal(dg[k][c][..., :n], r[k][..., :n]), ("convect_diagnostics", case, c, k)
        if kind == 2 and ncols <= 3000:           # convtran over a random constituent set (zeros / negatives included)
            ncnst = int(rng.integers(2, 14))
            q, fracis, pdeldry = S.make_tracers(ch, ncnst)
            q = q * rng.choice([0.0, 1.0, 1.0, 1.0, -1.0, 1e-30], size=q.shape)
            do = [0] + [int(x) for x in rng.integers(0, 2, ncnst - 1)]
            dry = [int(x) for x in rng.integers(0, 2, ncnst)]
            dpdry = dpdry_gathered(ch, ref, pdeldry)
            dq = Z.convtran(do, q, ref["mu"], ref["md"], ref["du"], ref["eu"], ref["ed"], ref["dp"], ref["dsubcld"],
                            ref["jt"], ref["maxg"], ref["ideep"], ref["lengath"], fracis, dpdry, ch.ztodt, dry,
                            dqdt=np.full_like(q, 7.25))
            for c in range(ch.nchunks):
                r = o.convtran(do, q[c], ref["mu"][c], ref["md"][c], ref["du"][c], ref["eu"][c], ref["ed"][c], ref["dp"][c],
                               ref["dsubcld"][c], ref["jt"][c], ref["maxg"][c], ref["ideep"][c], ref["lengath"][c],
                               fracis[c], dpdry[c], ch.ztodt, dry)
                for m in range(ncnst):
                    if do[m] and m >= 1:
                        assert np.array_equal(dq[c, m], r[m], equal_nan=True), ("convtran", case, c, m)
                    else:
                        assert np.all(dq[c, m] == 7.25), ("convtran untouched", case, c, m)
    except Z.ZmEndrun as e:                   # Brent did not converge: the oracle must say so too (endrun in the reference)
        assert ref["rc"] > 0, ("CUDA reports non-convergence, the oracle does not", case, str(e)[:200])
        report.append(dict(case=case, pver=pver, pcols=pcols, ncols=ncols, p_conv=pconv, seed=seed, options=over,
                           perturbed=perturbed, endrun=True))
        print(report[-1], flush=True)
        continue
    assert ref["rc"] == 0, ("the oracle reports non-convergence, CUDA does not", case)
    assert Z.lib().zm_sync_check(None) == 0
    rec = dict(case=case, pver=pver, pcols=pcols, ncols=ncols, p_conv=pconv, seed=seed, options=over, perturbed=perturbed,
               convective=int(out["lengath"].sum()), oracle_rc=int(ref["rc"]))
    report.append(rec)
    print(rec, flush=True)
print(json.dumps({"cases": len(report), "all_bit_exact": True, "seconds": round(time.time() - t00, 1),
                  "perturbed_cases": sum(1 for r in report if r["perturbed"]),
                  "endrun_cases": sum(1 for r in report if r.get("endrun")),
                  "convective_columns": sum(r.get("convective", 0) for r in report)}))
