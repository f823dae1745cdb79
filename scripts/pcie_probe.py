"""Raw pinned-memory PCIe rates on the box (what bounds the host-pointer e2e number)."""
import torch, time
n = 36 * 1024 * 1024  # 288 MB of doubles
h = torch.zeros(n, dtype=torch.float64).pin_memory(); d = torch.zeros(n, dtype=torch.float64, device="cuda")
h2 = torch.zeros(n // 2, dtype=torch.float64).pin_memory(); d2 = torch.zeros(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(fn, reps=5):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: h.copy_(d, non_blocking=True)); print(f"D2H 288 MB: {a*1e3:.2f} ms  {0.288*1.048576/a:.1f} GB/s")
b = t(lambda: d2.copy_(h2, non_blocking=True)); print(f"H2D 144 MB: {b*1e3:.2f} ms  {0.144*1.048576/b:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2): d2.copy_(h2, non_blocking=True)
c = t(both); print(f"both concurrently: {c*1e3:.2f} ms")
# many small copies (14 MB pieces) as the library issues them
pieces = [(h[i * (n // 20):(i + 1) * (n // 20)], d[i * (n // 20):(i + 1) * (n // 20)]) for i in range(20)]
e = t(lambda: [x.copy_(y, non_blocking=True) for x, y in pieces]); print(f"D2H 20 pieces: {e*1e3:.2f} ms")
# H2D with non-zero data, and in 14 MB pieces
hr = torch.randn(n // 2, dtype=torch.float64).pin_memory()
f = t(lambda: d2.copy_(hr, non_blocking=True)); print(f"H2D 144 MB random data: {f*1e3:.2f} ms  {0.144*1.048576/f:.1f} GB/s")
m = n // 2 // 10
pcs = [(d2[i * m:(i + 1) * m], hr[i * m:(i + 1) * m]) for i in range(10)]
g = t(lambda: [x.copy_(y, non_blocking=True) for x, y in pcs]); print(f"H2D 10 pieces: {g*1e3:.2f} ms")
import numpy as np
hn = torch.from_numpy(np.random.rand(n // 2)).pin_memory()
g2 = t(lambda: d2.copy_(hn, non_blocking=True)); print(f"H2D from_numpy pinned: {g2*1e3:.2f} ms")
s3 = torch.cuda.Stream()
def on_stream():
    with torch.cuda.stream(s3): d2.copy_(hn, non_blocking=True)
g3 = t(on_stream); print(f"H2D on side stream: {g3*1e3:.2f} ms")
