import sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
n = 4096
x = np.linspace(0, 1, n); y = np.linspace(0, 1, n)
z = np.asfortranarray(np.random.default_rng(2234).standard_normal((n, n)))
plan = B.Interp2Plan(x, y, z)
g2 = torch.Generator(device="cuda").manual_seed(2236)
xi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
yi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
for _ in range(3): out = plan.grid(xi, yi)
torch.cuda.synchronize()
