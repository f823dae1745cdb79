"""Lane-efficiency model of the thread-per-column CAPE kernel from the CPU oracle's Brent trace.
For groups of 32 consecutive columns (= one warp of k_buoyan_dilute<1>) compares
  lockstep : every inversion costs max-over-lanes evaluations (what the current kernel does)
  level    : lanes run freely inside one level, re-converge at level ends
  free     : lanes run freely through the whole sweep (flat state machine)
with the mean work per lane.  CPU only (oracle is test infrastructure)."""
import sys, os, ctypes as C
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from cam_nor_physics_b200 import soundings as S
from helpers import get_oracle

ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
o, p, rc = get_oracle("pm", 16, 32)
ch = S.make_chunks(ncols, 32, 16, p_conv=0.35)
cap = 4 * 40_000_000 // 10
buf = np.zeros(cap, np.int32)
o.lib.zmo_trace_set(buf.ctypes.data_as(C.POINTER(C.c_int)), C.c_int(cap))
ref = o.convr_batch(ch, nthreads=1)
n = o.lib.zmo_trace_count()
o.lib.zmo_trace_set(None, 0)
tr = buf[:4 * n].reshape(n, 4)
print("inversions traced", n, "mean evals", tr[:, 3].mean())
# split per chunk into pass 1 / pass 2, per column sequences
from collections import defaultdict
passes = {1: defaultdict(lambda: {2: [], 3: [], 4: []}), 2: defaultdict(lambda: {2: [], 3: [], 4: []})}
cur_pass = {}; seen4 = {}
for rcall, icol, lchnk, ev in tr:
    pz = cur_pass.get(lchnk, 1)
    if rcall == 2 and seen4.get(lchnk, False):
        pz = 2; cur_pass[lchnk] = 2; seen4[lchnk] = False
    if rcall == 4: seen4[lchnk] = True
    passes[pz][(lchnk, icol)][rcall].append(ev)
for pz in (1, 2):
    cols = sorted(passes[pz])
    if pz == 2:
        print("pass 2 columns", len(cols))
    tot_lock = tot_level = tot_free = tot_mean = 0.0
    for w0 in range(0, len(cols), 32):
        grp = [passes[pz][c] for c in cols[w0:w0 + 32]]
        nlev = max(len(g[2]) for g in grp)
        lock = 0.0; level = 0.0
        lane_tot = np.zeros(len(grp))
        for j in range(nlev):
            a = np.array([g[2][j] if j < len(g[2]) else 0 for g in grp])
            b = np.array([g[4][2 * j] if 2 * j < len(g[4]) else 0 for g in grp])
            c = np.array([g[4][2 * j + 1] if 2 * j + 1 < len(g[4]) else 0 for g in grp])
            act = a > 0
            lock += 2 + a.max() + b.max() + c.max()
            per = (a + b + c + 2 * act)
            level += per.max()
            lane_tot += per
        # LCL inversions: lockstep pays one max per distinct iteration index at which some lane has it (approx: each)
        lcl = np.array([sum(g[3]) for g in grp])
        lock += sum(max(g[3]) if g[3] else 0 for g in grp) * 0 + lcl.max() * min(len(grp), 4)   # ~4 distinct levels
        lane_tot += lcl
        tot_lock += lock; tot_level += level + lcl.max(); tot_free += lane_tot.max(); tot_mean += lane_tot.mean()
    print(f"pass {pz}: mean evals/lane-sweep per warp {tot_mean:.0f}; lockstep {tot_lock:.0f} (eff {tot_mean/tot_lock:.2f}); "
          f"level-sync {tot_level:.0f} (eff {tot_mean/tot_level:.2f}); free-running {tot_free:.0f} (eff {tot_mean/tot_free:.2f})")
