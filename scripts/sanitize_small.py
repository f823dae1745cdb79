"""Small end-to-end run for compute-sanitizer: zm_conv_tend + convtran + N4 kernels on 5 ragged chunks."""
import sys, os, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
from helpers import state_of, dpdry_gathered
Z.zm_init(Z.default_params(16, 32, 3))
ch = S.make_chunks(72, 32, 16, p_conv=0.7)
out = Z.zm_conv_tend(ch.ncol, state_of(ch), ch.ztodt)
q, fracis, pdeldry = S.make_tracers(ch, 5)
dq = Z.convtran([0, 1, 1, 0, 1], q, out["mu"], out["md"], out["du"], out["eu"], out["ed"], out["dp"], out["dsubcld"],
                out["jt"], out["maxg"], out["ideep"], out["lengath"], fracis, dpdry_gathered(ch, out, pdeldry), ch.ztodt, [0, 0, 1, 0, 1])
zi, zm = Z.geopotential_t(ch.ncol, np.log(ch.pint), np.log(ch.pmid), ch.pint, ch.pmid, ch.pdel, 1 / ch.pdel, ch.t, ch.q,
                          np.full_like(ch.t, S.RAIR), S.GRAVIT, np.full_like(ch.t, S.ZVIR))
print("ok", int(out["lengath"].sum()), float(np.abs(dq).max()), float(zi.max()))
