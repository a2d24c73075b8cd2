"""Banded scattered interp2 at BASELINE config 2: per-pass times for the current environment settings."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
n = 4096
x = np.linspace(0, 1, n); y = np.linspace(0, 1, n)
z = np.asfortranarray(np.random.default_rng(2234).standard_normal((n, n)))
nq = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
g = torch.Generator(device="cuda").manual_seed(2235)
xq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
yq = torch.rand(nq, generator=g, device="cuda", dtype=torch.float64)
plan = B.Interp2Plan(x, y, z, flags=8)
zq = torch.empty_like(xq)
for _ in range(3): plan.scattered(xq, yq, out=zq)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): plan.scattered(xq, yq, out=zq)
e1.record(); torch.cuda.synchronize()
print(f"BAND_MIB={os.environ.get('B200_INTERP2_BAND_MIB')} : {e0.elapsed_time(e1) / 10:.3f} ms per {nq:.0e} queries", flush=True)
