"""Driver for ncu: every kernel of the path once per step on the f09 shard, device-resident -- zm_conv_tend
(zm_convr + physics_update + zm_conv_evap + momtran + convtran1) with the 41-constituent stand-in, zm_conv_tend_2
(convtran2), the budget reduction, the history diagnostics, geopotential_t and convect_diagnostics_calc."""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import numpy as np, torch
from cam_nor_physics_b200 import soundings as S, zm_conv as Z
from cam_nor_physics_b200.device import DeviceTend, _ptr
ncols = int(sys.argv[1]) if len(sys.argv) > 1 else 55296
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
L = 32
os.environ["ZM_DEV_GRAPH"] = "0"          # plain launches (a replayed graph is profiled as its kernels anyway)
Z.zm_init(Z.default_params(16, L, S.limcnv_for(L)))
ch = S.make_chunks(ncols, L, 16, p_conv=0.35)
dev = DeviceTend(ch, ncnst=41)
nch = ch.nchunks
f64 = dict(dtype=torch.float64, device="cuda")
d = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
ps = d(ch.pint[:, -1, :]); piln = d(np.log(ch.pint)); rpdel = d(1.0 / ch.pdel)
rair = torch.full((nch, L, 16), 287.04, **f64); zvir = torch.full((nch, L, 16), 0.608, **f64)
zi = torch.zeros((nch, L + 1, 16), **f64); zm = torch.zeros((nch, L, 16), **f64)
diag = {k: torch.zeros((nch, 16), **f64) for k in ("freqzm", "pcont", "pconb", "rliq2", "pcnt", "pcnb")}
mu_out = torch.zeros((nch, L, 16), **f64); md_out = torch.zeros((nch, L, 16), **f64)
qc2 = torch.zeros((nch, L, 16), **f64); rprdsh = torch.zeros((nch, L, 16), **f64); rprdtot = torch.zeros((nch, L, 16), **f64)
cmfmc2 = torch.zeros((nch, L + 1, 16), **f64)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
lib = Z.lib()
for r in range(reps):
    dev.step(); dev.step2(); dev.conservation()
    o = dev.out
    lib.zm_conv_tend_diag_batch_dev(C.c_int(nch), _ptr(dev.ncol), _ptr(ps), _ptr(dev.inp["pmid"]), _ptr(o["mu"]), _ptr(o["md"]),
                                    _ptr(o["jt"]), _ptr(o["maxg"]), _ptr(o["ideep"]), _ptr(o["lengath"]), _ptr(diag["freqzm"]),
                                    _ptr(mu_out), _ptr(md_out), _ptr(diag["pcont"]), _ptr(diag["pconb"]), s)
    lib.zm_geopotential_t_batch_dev(C.c_int(nch), _ptr(dev.ncol), C.c_int(1), _ptr(piln), None, _ptr(dev.inp["pint"]),
                                    _ptr(dev.inp["pmid"]), _ptr(dev.inp["pdel"]), _ptr(rpdel), _ptr(dev.inp["t"]), _ptr(dev.inp["q"]),
                                    _ptr(rair), C.c_double(9.80616), _ptr(zvir), _ptr(zi), _ptr(zm), s)
    lib.zm_convect_diagnostics_batch_dev(C.c_int(nch), _ptr(dev.ncol), _ptr(o["mcon"]), _ptr(o["ql"]), _ptr(qc2), _ptr(o["rliq"]),
                                         _ptr(diag["rliq2"]), _ptr(dev.inp["pmid"]), _ptr(o["rprd"]), _ptr(o["jctop"]), _ptr(o["jcbot"]),
                                         _ptr(cmfmc2), _ptr(rprdsh), _ptr(rprdtot), _ptr(diag["pcnt"]), _ptr(diag["pcnb"]), s)
    torch.cuda.synchronize()
print("convective", int(dev.out["lengath"].sum().item()), "failures", dev.check())
