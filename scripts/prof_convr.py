"""Small driver for ncu: one f09-sized zm_convr + evap + momtran step through the host API."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 55296
L = int(sys.argv[2]) if len(sys.argv) > 2 else 32
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
lim = S.limcnv_for(L)
Z.zm_init(Z.default_params(16, L, lim))
ch = S.make_chunks(ncols, L, 16, p_conv=0.35)
for r in range(reps):
    out = Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel, 0.5 * ch.ztodt, ch.tpert, ch.landfrac)
print("triggered", int(out["lengath"].sum()), "prec mean mm/day", float(out["prec"].mean() * 86400e3))
