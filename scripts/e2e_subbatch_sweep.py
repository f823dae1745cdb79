"""e2e time of zm_conv_tend_batch (pinned host buffers) for different sub-batch counts."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 55296
L = 32
Z.zm_init(Z.default_params(16, L, S.limcnv_for(L)))
ch = S.make_chunks(ncols, L, 16, p_conv=0.35)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
st = {k: pin(getattr(ch, k)) for k in Z.TEND_IN_ORDER}
nch, pc = ch.nchunks, 16
out = {}
for k in Z.TEND_OUT_2D: out[k] = torch.zeros((nch, L, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_2DP: out[k] = torch.zeros((nch, L + 1, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_1D: out[k] = torch.zeros((nch, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_INT: out[k] = torch.zeros((nch, pc), dtype=torch.int32).pin_memory().numpy()
out["lengath"] = torch.zeros(nch, dtype=torch.int32).pin_memory().numpy()
ref = None
for nb in (1, 2, 3, 4, 6, 8, 4):
    os.environ["ZM_TEND_SUBBATCHES"] = str(nb)
    for _ in range(2): Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
    t0 = time.perf_counter()
    for _ in range(5): Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
    dt = (time.perf_counter() - t0) / 5
    chk = {k: v.copy() for k, v in out.items()}
    if ref is None: ref = chk
    same = all(np.array_equal(ref[k], chk[k]) for k in ref)
    print(np.round(Z.tend_trace(), 2))
    print(f"NB={nb}: {dt*1e3:.2f} ms/step  {ncols/dt/1e6:.2f} M col/s  identical_to_NB1={same}")
