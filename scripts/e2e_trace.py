"""Timeline (ms) of one pipelined zm_conv_tend_batch call on the f09 shard: rows = sub-batches; columns = inputs on
device, late inputs on device, zm_convr done, kernels done, zm_convr outputs on host, all outputs on host."""
import sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
L = 32
Z.zm_init(Z.default_params(16, L, S.limcnv_for(L)))
ch = S.make_chunks(55296, L, 16, p_conv=0.35)
pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory().numpy()
st = {k: pin(getattr(ch, k)) for k in Z.TEND_IN_ORDER}
nch, pc = ch.nchunks, 16
out = {}
for k in Z.TEND_OUT_2D: out[k] = torch.zeros((nch, L, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_2DP: out[k] = torch.zeros((nch, L + 1, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_1D: out[k] = torch.zeros((nch, pc), dtype=torch.float64).pin_memory().numpy()
for k in Z.TEND_OUT_INT: out[k] = torch.zeros((nch, pc), dtype=torch.int32).pin_memory().numpy()
out["lengath"] = torch.zeros(nch, dtype=torch.int32).pin_memory().numpy()
for _ in range(3): Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
t0 = time.perf_counter()
for _ in range(5): Z.zm_conv_tend(ch.ncol, st, ch.ztodt, out)
print("ms/step", (time.perf_counter() - t0) / 5 * 1e3)
print(np.round(Z.tend_trace(), 2))
