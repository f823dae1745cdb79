"""How much of the per-column latency is warp divergence?  Times k_buoyan_dilute on 64 copies of one
column vs 64 different columns (both one block of work: pure latency, no throughput effects)."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
Z.zm_init(Z.default_params(16, 32, 3)); Z.lib().zm_set_profiling(1)
ch = S.make_chunks(64, 32, 16, p_conv=1.0)
def run(ch, tag):
    for _ in range(3):
        out = Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel, 900.0, ch.tpert, ch.landfrac)
    print(tag, {k: round(v, 4) for k, v in Z.kernel_times()}, "lengath", out["lengath"].tolist())
run(ch, "64 different columns ")
import copy
for src in (0, 5, 17):
    c2 = copy.deepcopy(ch)
    cc, ii = divmod(src, 16)
    for name in ["t", "q", "pmid", "pint", "pdel", "zm", "zi", "u", "v", "cld"]:
        a = getattr(c2, name); a[:] = a[cc, :, ii][None, :, None]
    for name in ["phis", "pblh", "tpert", "landfrac"]:
        a = getattr(c2, name); a[:] = a[cc, ii]
    run(c2, f"64 copies of column {src}")
