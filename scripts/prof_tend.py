"""Driver for ncu: one f09-sized zm_conv_tend step through the host API."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 55296
L = 32
Z.zm_init(Z.default_params(16, L, S.limcnv_for(L)))
ch = S.make_chunks(ncols, L, 16, p_conv=0.35)
st = dict(t=ch.t, q=ch.q, u=ch.u, v=ch.v, pmid=ch.pmid, pint=ch.pint, pdel=ch.pdel, zm=ch.zm, zi=ch.zi,
          phis=ch.phis, pblh=ch.pblh, tpert=ch.tpert, landfrac=ch.landfrac, cld=ch.cld)
out = Z.zm_conv_tend(ch.ncol, st, ch.ztodt)
print("convective", int(out["lengath"].sum()))
