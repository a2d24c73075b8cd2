"""interp2 grid kernel: rows per thread (B200_INTERP2_GRID_V) and the write-only ceiling of the GPU."""
import os, sys
sys.path.insert(0, "/root/repo")
from armadillocudalinearinterpolation_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = sys.argv[1]
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
n = 4096
x = np.linspace(0, 1, n); y = np.linspace(0, 1, n)
z = np.asfortranarray(np.random.default_rng(2234).standard_normal((n, n)))
plan = B.Interp2Plan(x, y, z)
g2 = torch.Generator(device="cuda").manual_seed(2236)
xi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
yi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
def timeit(f, n=20):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
ref = None
for v in ("1", "2", "4"):
    os.environ["B200_INTERP2_GRID_V"] = v
    ms = timeit(lambda: plan.grid(xi, yi))
    out = plan.grid(xi, yi)
    if ref is None: ref = out.clone()
    print(f"V={v}: {ms:.4f} ms  {(8e8 + 134217728 + 16e4) / ms / 1e6:.0f} GB/s algorithmic  same bits: {torch.equal(ref.view(torch.int64), out.view(torch.int64))}")
buf = torch.empty(100_000_000, dtype=torch.float64, device="cuda")
ms = timeit(lambda: buf.fill_(1.5))
print(f"torch fill_ 0.8 GB: {ms:.4f} ms  {0.8e9 / ms / 1e6:.0f} GB/s")
bufs = [torch.empty(100_000_000, dtype=torch.float64, device="cuda") for _ in range(4)]
i = [0]
def rot():
    bufs[i[0] % 4].fill_(2.5); i[0] += 1
ms = timeit(rot)
print(f"torch fill_ 0.8 GB rotating over 4 buffers: {ms:.4f} ms  {0.8e9 / ms / 1e6:.0f} GB/s")
