import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
Z.zm_init(Z.default_params(16, 32, 3)); Z.lib().zm_set_profiling(1)
for pc in (1.0, 0.0, 0.35):
    ch = S.make_chunks(55296, 32, 16, p_conv=pc)
    for _ in range(2):
        out = Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel, 900.0, ch.tpert, ch.landfrac)
    kt = dict(Z.kernel_times())
    print("p_conv", pc, "triggered", int(out["lengath"].sum()), {k: round(v, 3) for k, v in kt.items() if "buoyan" in k or "plume" in k})
# sorted by type: tropical first then stable (same columns as p_conv=0.35, permuted)
ch = S.make_chunks(55296, 32, 16, p_conv=0.35)
key = ch.t[:, -1, :].reshape(-1)          # surface-layer temperature
order = np.argsort(-key, kind="stable")
import copy
c2 = copy.deepcopy(ch)
def perm2(a):
    nch, L, pc = a.shape
    flat = a.transpose(0, 2, 1).reshape(nch * pc, L)[order]
    return np.ascontiguousarray(flat.reshape(nch, pc, L).transpose(0, 2, 1))
def perm1(a):
    return np.ascontiguousarray(a.reshape(-1)[order].reshape(a.shape))
for n in ["t", "q", "pmid", "pint", "pdel", "zm", "zi"]:
    setattr(c2, n, perm2(getattr(ch, n)))
for n in ["phis", "pblh", "tpert", "landfrac"]:
    setattr(c2, n, perm1(getattr(ch, n)))
for _ in range(2):
    out = Z.zm_convr(c2.ncol, c2.t, c2.q, c2.pblh, c2.zm, c2.phis, c2.zi, c2.pmid, c2.pint, c2.pdel, 900.0, c2.tpert, c2.landfrac)
kt = dict(Z.kernel_times())
print("sorted by surface T: triggered", int(out["lengath"].sum()), {k: round(v, 3) for k, v in kt.items() if "buoyan" in k or "plume" in k})
