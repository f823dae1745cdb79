"""Generates tests/golden/*.npz: inputs + outputs of the CPU oracle (glibc-libm flavour, the one closest
to a gfortran build of the reference) for BASELINE config 1 (one pcols=16 chunk of tropical soundings,
pver=32) and a 4-chunk mixed case with tracers.  The reference ships no golden vectors (SURVEY.md
section 4) and cannot be compiled here, so these pin the oracle against regressions, not against the
Fortran.  Run from the repo root:  python tests/golden/make_golden.py"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S
from helpers import get_oracle, state_of, dpdry_gathered

HERE = os.path.dirname(os.path.abspath(__file__))


def make(name, ncols, pconv, ncnst):
    L, pc = 32, 16
    o, p, rc = get_oracle("libm", pc, L)
    ch = S.make_chunks(ncols, L, pc, p_conv=pconv)
    ref = o.conv_tend_batch(ch)
    cv = o.convr_batch(ch)
    assert ref["rc"] == 0 and cv["rc"] == 0
    q, fracis, pdeldry = S.make_tracers(ch, ncnst)
    do = np.array([0] + [1 if m % 3 else 0 for m in range(1, ncnst)], np.int32)
    dry = np.array([0] + [1 if m % 2 == 0 else 0 for m in range(1, ncnst)], np.int32)
    dpdry = dpdry_gathered(ch, ref, pdeldry)
    dq = np.stack([o.convtran(do, q[c], ref["mu"][c], ref["md"][c], ref["du"][c], ref["eu"][c], ref["ed"][c],
                              ref["dp"][c], ref["dsubcld"][c], ref["jt"][c], ref["maxg"][c], ref["ideep"][c],
                              ref["lengath"][c], fracis[c], dpdry[c], ch.ztodt, dry) for c in range(ch.nchunks)])
    d = {"in_" + k: v for k, v in state_of(ch).items()}
    d.update({"in_ncol": ch.ncol, "in_ztodt": ch.ztodt, "in_tracers": q, "in_fracis": fracis, "in_dpdry": dpdry,
              "in_doconvtran": do, "in_cnst_is_dry": dry})
    d.update({"tend_" + k: v for k, v in ref.items() if k != "rc"})
    d.update({"convr_" + k: v for k, v in cv.items() if k != "rc"})
    d["convtran_dqdt"] = dq
    np.savez_compressed(os.path.join(HERE, name + ".npz"), **d)
    print(name, "lengath", ref["lengath"], "bytes", os.path.getsize(os.path.join(HERE, name + ".npz")))


if __name__ == "__main__":
    make("config1_L32_pcols16", 16, 1.0, 5)
    make("mixed4_L32_pcols16", 64, 0.5, 7)
