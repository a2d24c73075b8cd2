"""python tools/interp2_ab.py <lib.so>: headline kernel (config 2, default layout) with another build of the library."""
import sys
sys.path.insert(0, "/root/repo")
from armadillocudalinearinterpolation_b200 import _lib
if len(sys.argv) > 1: _lib.LIB_PATH = sys.argv[1]
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
n = 4096
x = np.linspace(0, 1, n); y = np.linspace(0, 1, n)
z = np.asfortranarray(np.random.default_rng(2234).standard_normal((n, n)))
nq = 100_000_000
g = torch.Generator(device="cuda").manual_seed(2235)
xq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
yq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
plan = B.Interp2Plan(x, y, z)
zq = torch.empty_like(xq)
for _ in range(3): plan.scattered(xq, yq, out=zq)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): plan.scattered(xq, yq, out=zq)
e1.record(); torch.cuda.synchronize()
print(_lib.LIB_PATH.split("/")[-1], f"{e0.elapsed_time(e1) / 20:.3f} ms per 1e8 queries", float(zq.sum()))
