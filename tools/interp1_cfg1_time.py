"""interp1 at BASELINE configs[0] size (1e6 knots, 1e7 queries), 8 rotating buffer pairs, as in bench.py."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
ng, ni, nbuf = 1_000_000, 10_000_000, 8
xg = np.linspace(0.0, 1.0, ng); yg = np.sin(2 * np.pi * xg)
p1 = B.Interp1Plan(xg, yg)
g1 = torch.Generator(device="cuda").manual_seed(1236)
for order in ("unsorted", "sorted"):
    qs = [torch.rand(ni, generator=g1, device="cuda", dtype=torch.float64) for _ in range(nbuf)]
    if order == "sorted": qs = [q.sort().values for q in qs]
    outs = [torch.empty_like(q) for q in qs]
    for i in range(16): p1(qs[i % nbuf], out=outs[i % nbuf])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(200): p1(qs[i % nbuf], out=outs[i % nbuf])
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 200
    print(f"mult={os.environ.get('B200_INTERP1_GRID_MULT')} {order}: {ms * 1e3:.1f} us  frac {(16 * ni + 16 * ng) / ms / 1e6 / 6537:.3f}")
