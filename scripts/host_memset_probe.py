"""Host-side bandwidth probe for the sparse return path of zm_conv_tend_batch: parallel memset of pinned memory
(what zero-filling the caller's dense output arrays costs) against the PCIe D2H of the same bytes."""
import ctypes, sys, threading, time, os
import torch
n = int(sys.argv[1]) if len(sys.argv) > 1 else 289 * 1000 * 1000
buf = torch.empty(n, dtype=torch.uint8).pin_memory()
dev = torch.empty(n, dtype=torch.uint8, device="cuda")
base = buf.data_ptr()
print("cpus", os.cpu_count())
for nt in (1, 2, 4, 8, 16, 32):
    best = 1e9
    for rep in range(4):
        per = (n + nt - 1) // nt
        ths = [threading.Thread(target=ctypes.memset, args=(base + i * per, 0, min(per, n - i * per))) for i in range(nt)]
        t0 = time.perf_counter()
        [t.start() for t in ths]; [t.join() for t in ths]
        best = min(best, time.perf_counter() - t0)
    print(f"memset {n/1e6:.0f} MB with {nt:2d} threads: {best*1e3:.2f} ms = {n/best/1e9:.1f} GB/s")
torch.cuda.synchronize()
for rep in range(3):
    t0 = time.perf_counter(); buf.copy_(dev, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"D2H {n/1e6:.0f} MB: {dt*1e3:.2f} ms = {n/dt/1e9:.1f} GB/s")
for rep in range(3):
    t0 = time.perf_counter(); dev.copy_(buf, non_blocking=True); torch.cuda.synchronize(); dt = time.perf_counter() - t0
print(f"H2D {n/1e6:.0f} MB: {dt*1e3:.2f} ms = {n/dt/1e9:.1f} GB/s")
