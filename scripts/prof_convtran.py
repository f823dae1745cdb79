"""Driver: f09 zm_convr then convtran over ncnst constituents (BASELINE config 4) through the host API."""
import sys, os, time, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 55296
ncnst = int(sys.argv[2]) if len(sys.argv) > 2 else 41
L = 32
Z.zm_init(Z.default_params(16, L, S.limcnv_for(L)))
ch = S.make_chunks(ncols, L, 16, p_conv=0.35)
r = Z.zm_convr(ch.ncol, ch.t, ch.q, ch.pblh, ch.zm, ch.phis, ch.zi, ch.pmid, ch.pint, ch.pdel, 0.5 * ch.ztodt, ch.tpert, ch.landfrac)
q, fracis, pdeldry = S.make_tracers(ch, ncnst)
from helpers import dpdry_gathered
dpdry = dpdry_gathered(ch, r, pdeldry)
do = [0] + [1] * (ncnst - 1); dry = [0] + [m % 2 for m in range(1, ncnst)]
for _ in range(2):
    t0 = time.time()
    dq = Z.convtran(do, q, r["mu"], r["md"], r["du"], r["eu"], r["ed"], r["dp"], r["dsubcld"], r["jt"], r["maxg"], r["ideep"], r["lengath"], fracis, dpdry, ch.ztodt, dry)
    print("convtran host call", time.time() - t0, "s; nonzero", int(np.count_nonzero(dq)))
