"""Two banded scattered calls at BASELINE config 2 (for ncu)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
n = 4096
x = np.linspace(0, 1, n); y = np.linspace(0, 1, n)
z = np.asfortranarray(np.random.default_rng(2234).standard_normal((n, n)))
nq = 100_000_000
g = torch.Generator(device="cuda").manual_seed(2235)
xq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
yq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
plan = B.Interp2Plan(x, y, z, flags=8)
zq = torch.empty_like(xq)
for _ in range(2): plan.scattered(xq, yq, out=zq)
torch.cuda.synchronize()
print("done", float(zq[0]))
