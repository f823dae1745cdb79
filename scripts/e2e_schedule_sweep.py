"""e2e time of zm_conv_tend_batch (pinned host buffers) for explicit sub-batch schedules (sixteenths of the batch)."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
ncols, L = 55296, 32
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
scheds = sys.argv[1:] or ["1,1,2,4,4,4", "1,1,2,3,4,5", "1,2,3,4,6", "2,2,4,4,4", "1,1,1,2,3,4,4", "1,1,2,2,3,3,4", "1,2,2,3,4,4",
                          "1,1,2,4,8", "2,2,3,4,5", "1,1,2,4,4,4"]
for sc in scheds:
    os.environ["ZM_TEND_SCHEDULE"] = sc
    for _ in range(3): Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
    ts = []
    for _ in range(8):
        t0 = time.perf_counter(); Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out); ts.append(time.perf_counter() - t0)
    print(f"{sc:18s}: median {np.median(ts)*1e3:.2f} ms  min {min(ts)*1e3:.2f} ms   {ncols/np.median(ts)/1e6:.2f} M col/s", flush=True)
