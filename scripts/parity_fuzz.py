"""Randomised parity sweep: CUDA zm_convr / zm_conv_tend vs the CPU oracle (portable-math flavour, bit-exact) over random seeds,
batch sizes, chunk widths, level counts, convective fractions and namelist options.  Runs on the GPU box
(`gpurun -- python scripts/parity_fuzz.py [ncases] [seed]`); the oracle is the checker, never the thing measured."""
import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from cam_nor_physics_b200 import soundings as S
from helpers import get_oracle, init_cuda, cuda_convr, assert_same, state_of, CONVR_KEYS, TEND_KEYS

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
    pconv = float(rng.choice([0.0, 0.1, 0.35, 0.6, 1.0]))
    seed = int(rng.integers(1, 2**31))
    over = dict(OPTS[int(rng.integers(0, len(OPTS)))])
    Z = init_cuda(pcols, pver, **over)
    o, _, rc = get_oracle("pm", pcols, pver, **over)
    assert rc == 0, (rc, over)
    ch = S.make_chunks(ncols, pver, pcols, p_conv=pconv, seed=seed)
    if case % 2 == 0:                         # zm_convr alone / the whole zm_conv_tend sequence, alternating
        ref = o.convr_batch(ch)
        out = cuda_convr(Z, ch)
        assert_same(out, ref, CONVR_KEYS, pcols, exact=True, what=f"fuzz case {case} (zm_convr)")
    else:
        ref = o.conv_tend_batch(ch)
        out = Z.zm_conv_tend(ch.ncol, state_of(ch), ch.ztodt)
        assert_same(out, ref, TEND_KEYS, pcols, exact=True, what=f"fuzz case {case} (zm_conv_tend)")
    assert Z.lib().zm_sync_check(None) == 0
    rec = dict(case=case, pver=pver, pcols=pcols, ncols=ncols, p_conv=pconv, seed=seed, options=over,
               convective=int(out["lengath"].sum()), oracle_rc=int(ref["rc"]))
    report.append(rec)
    print(rec, flush=True)
print(json.dumps({"cases": len(report), "all_bit_exact": True, "seconds": round(time.time() - t00, 1),
                  "convective_columns": sum(r["convective"] for r in report)}))
