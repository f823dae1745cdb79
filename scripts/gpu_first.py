"""First GPU sanity run: math bit-exactness, FP64 peak, convr/evap/momtran/convtran vs oracle."""
import sys, time, os, subprocess, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
from oracle_lib import Oracle

print("gfortran:", subprocess.run("which gfortran flang nvfortran; nproc; lscpu | grep 'Model name'", shell=True, capture_output=True, text=True).stdout)
rng = np.random.default_rng(0)
for fid, x, y in [(0, np.exp(rng.uniform(-30, 30, 100000)), None), (1, rng.uniform(0.5, 4, 100000), None),
                  (2, rng.uniform(-50, 50, 100000), None), (3, rng.uniform(-8, 8, 100000), None),
                  (4, rng.uniform(0.5, 40, 100000), rng.uniform(-3, 3, 100000))]:
    d = Z.math_eval(fid, x, y, device=True); h = Z.math_eval(fid, x, y, device=False)
    print("math", fid, "device!=host:", int((d != h).sum()))
print("fp64 peak TFLOP/s:", Z.fp64_peak_flops(20000) / 1e12)

L = 32; lim = S.limcnv_for(L)
o = Oracle("pm"); p = o.default_params(16, L, lim); o.convi(p)
zp = Z.default_params(16, L, lim); Z.zm_init(zp)
def cmp(name, a, b):
    a = np.asarray(a); b = np.asarray(b)
    ne = int((a != b).sum())
    if ne:
        d = np.abs(a.astype(float) - b.astype(float)); print(f"  {name}: {ne} differ, maxabs {d.max():.3e}, maxrel {(d/np.maximum(np.abs(b),1e-300))[d>0].max():.3e}")
    return ne
for ncols, pconv in [(16, 1.0), (4096, 0.5), (55296, 0.35)]:
    ch = S.make_chunks(ncols, L, 16, p_conv=pconv)
    t0 = time.time(); ref = o.convr_batch(ch, nthreads=0); tcpu = time.time() - t0
    Z.lib().zm_set_profiling(1)
    t0 = time.time()
    out = Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel, 0.5 * ch.ztodt, ch.tpert, ch.landfrac)
    tg = time.time() - t0
    print(f"ncols={ncols} pconv={pconv} triggered={int(ref['lengath'].sum())} cpu {tcpu:.3f}s gpu-call {tg:.3f}s", Z.kernel_times())
    bad = 0
    for k in ["lengath", "ideep", "cape", "prec", "jctop", "jcbot", "qtnd", "heat", "mcon", "cme", "eurt", "dlf", "pflx", "zdu", "rprd", "ql", "rliq", "rice"]:
        bad += cmp(k, out[k], ref[k])
    # gathered outputs: compare rows < lengath
    pc = 16
    mask = (np.arange(pc)[None, :] < ref["lengath"][:, None])
    for k in ["mu", "md", "du", "eu", "ed", "dp"]:
        bad += cmp(k, out[k] * mask[:, None, :], ref[k] * mask[:, None, :])
    for k in ["dsubcld", "jt", "maxg"]:
        bad += cmp(k, out[k] * mask, ref[k] * mask)
    print("  total mismatches:", bad)
    if ncols == 4096:
        # evap / momtran / convtran on this state
        nch = ch.nchunks
        t1 = ch.t + ref["heat"] * ch.ztodt / p.cpair
        q1 = np.maximum(ch.q + ref["qtnd"] * ch.ztodt, 1e-12)
        ev = Z.zm_conv_evap(ch.ncol, t1, ch.pmid, ch.pdel, q1, ch.landfrac, ref["rprd"], ch.cld, ch.ztodt, ref["prec"])
        bad = 0
        for c in range(nch):
            r = o.conv_evap(int(ch.ncol[c]), t1[c], ch.pmid[c], ch.pdel[c], q1[c], ch.landfrac[c], ref["rprd"][c], ch.cld[c], ch.ztodt, ref["prec"][c])
            for k in r: bad += int((ev[k][c] != r[k]).sum())
        print("  evap mismatches:", bad)
        winds = np.stack([ch.u, ch.v], axis=1)
        mo = Z.momtran(ch.ncol, [1, 1], winds, ref["mu"], ref["md"], ref["du"], ref["eu"], ref["ed"], ref["dp"], ref["dsubcld"], ref["jt"], ref["maxg"], ref["ideep"], ref["lengath"], ch.ztodt)
        bad = 0
        for c in range(nch):
            r = o.momtran(int(ch.ncol[c]), [1, 1], winds[c], ref["mu"][c], ref["md"][c], ref["du"][c], ref["eu"][c], ref["ed"][c], ref["dp"][c], ref["dsubcld"][c], ref["jt"][c], ref["maxg"][c], ref["ideep"][c], ref["lengath"][c], ch.ztodt)
            for k in r:
                nb = int((mo[k][c] != r[k]).sum()); bad += nb
                if nb and bad < 50: print("   mom", c, k, nb, np.abs(mo[k][c]-r[k]).max())
        print("  momtran mismatches:", bad)
        ncnst = 6
        qt, fracis, pdeldry = S.make_tracers(ch, ncnst)
        do = [0, 1, 1, 0, 1, 1]; dry = [0, 0, 1, 0, 0, 1]
        dpdry = np.zeros_like(ch.pdel)
        for c in range(nch):
            n = ref["lengath"][c]
            idx = ref["ideep"][c][:n] - 1
            dpdry[c][:, :n] = pdeldry[c][:, idx] / 100.0
        dq = Z.convtran(do, qt, ref["mu"], ref["md"], ref["du"], ref["eu"], ref["ed"], ref["dp"], ref["dsubcld"], ref["jt"], ref["maxg"], ref["ideep"], ref["lengath"], fracis, dpdry, ch.ztodt, dry)
        bad = 0
        for c in range(nch):
            r = o.convtran(do, qt[c], ref["mu"][c], ref["md"][c], ref["du"][c], ref["eu"][c], ref["ed"][c], ref["dp"][c], ref["dsubcld"][c], ref["jt"][c], ref["maxg"][c], ref["ideep"][c], ref["lengath"][c], fracis[c], dpdry[c], ch.ztodt, dry)
            bad += int((dq[c] != r).sum())
        print("  convtran mismatches:", bad, "nonzero:", int((dq != 0).sum()))
