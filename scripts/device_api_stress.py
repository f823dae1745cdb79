"""Stress of the device-resident entry point zm_conv_tend_batch_dev: several chunk sets of random sizes are stepped in
a random interleaved order (direct call, graph capture, graph replay, replay after another set ran, new input values
behind the same pointers, a set freed and another allocated in its place, side streams).  Every step must equal the
host-pointer API (which the parity sweep ties to the oracle) bit for bit.  GPU box only."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np, torch
from cam_nor_physics_b200 import soundings as S
from cam_nor_physics_b200.device import DeviceTend
from helpers import init_cuda, state_of, assert_same, TEND_KEYS

nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 5)
Z = init_cuda(16, 32)
sets = {}


def new_set(slot):
    ncols = int(rng.choice([16 * 40, 16 * 300 - 7, 16 * 1100, 16 * 2500 + 3]))
    ch = S.make_chunks(ncols, 32, 16, p_conv=float(rng.choice([0.1, 0.35, 0.7])), seed=int(rng.integers(1, 2**31)))
    ref = Z.zm_conv_tend(ch.ncol, state_of(ch), ch.ztodt)
    sets[slot] = dict(ch=ch, ref=ref, dev=DeviceTend(ch), steps=0)


for slot in range(3):
    new_set(slot)
side = torch.cuda.Stream()
log = []
for it in range(nsteps):
    slot = int(rng.integers(0, 3))
    action = rng.choice(["step", "step", "step", "newvalues", "realloc", "sidestream"])
    s = sets[slot]
    if action == "realloc":                      # free the set, allocate another one (addresses may be reused)
        del sets[slot]; del s
        torch.cuda.synchronize(); torch.cuda.empty_cache()
        new_set(slot); s = sets[slot]
    elif action == "newvalues":                  # new inputs behind the same device pointers
        ch = S.make_chunks(s["ch"].ncols_total, 32, 16, p_conv=float(rng.choice([0.2, 0.5])), seed=int(rng.integers(1, 2**31)))
        s["ref"] = Z.zm_conv_tend(ch.ncol, state_of(ch), ch.ztodt)
        for k in Z.TEND_IN_ORDER:
            s["dev"].inp[k].copy_(torch.from_numpy(np.ascontiguousarray(getattr(ch, k))))
        s["dev"].ncol.copy_(torch.from_numpy(ch.ncol.astype("int32")))
        s["ch"] = ch
        torch.cuda.synchronize()
    dev = s["dev"]
    for v in dev.out.values():
        v.fill_(-7)
    torch.cuda.synchronize()
    if action == "sidestream":
        with torch.cuda.stream(side):
            dev.step(); nfail = dev.check()
    else:
        dev.step(); nfail = dev.check()
    torch.cuda.synchronize()
    assert nfail == 0
    got = {k: dev.out[k].cpu().numpy() for k in TEND_KEYS}
    assert_same(got, s["ref"], TEND_KEYS, 16, exact=True, what=f"device step {it} ({action}, slot {slot})")
    s["steps"] += 1
    log.append((it, slot, str(action), int(s["ch"].ncols_total), s["steps"]))
print(json.dumps({"steps": len(log), "all_equal_host_api": True, "actions": {a: sum(1 for l in log if l[2] == a) for a in
                  ("step", "newvalues", "realloc", "sidestream")}}))
