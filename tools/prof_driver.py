"""Small driver for ncu: one launch (after warm-up) of each hot kernel at bench sizes."""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
import bench
which = sys.argv[1] if len(sys.argv) > 1 else "all"
if which in ("all", "interp2"):
    plan = B.Interp2Plan(*bench.make_grid())
    g = torch.Generator(device="cuda").manual_seed(2235)
    xq = torch.rand(bench.NQ, generator=g, device="cuda", dtype=torch.float64)
    yq = torch.rand(bench.NQ, generator=g, device="cuda", dtype=torch.float64)
    zq = torch.empty_like(xq)
    for _ in range(3):
        plan.scattered(xq, yq, out=zq)
    xi = torch.rand(10_000, generator=g, device="cuda", dtype=torch.float64).sort().values
    yi = torch.rand(10_000, generator=g, device="cuda", dtype=torch.float64).sort().values
    for _ in range(2):
        plan.grid(xi, yi)
    torch.cuda.synchronize()
    del xq, yq, zq
if which in ("all", "interp1"):
    rng = np.random.default_rng(1234)
    ng, ni = 1_000_000, 10_000_000
    xg = np.cumsum(0.5 + rng.random(ng)); xg = (xg - xg[0]) / (xg[-1] - xg[0])
    yg = np.sin(2 * np.pi * xg)
    p1 = B.Interp1Plan(xg, yg)
    g1 = torch.Generator(device="cuda").manual_seed(1236)
    qs = [torch.rand(ni, generator=g1, device="cuda", dtype=torch.float64) for _ in range(4)]
    out = torch.empty_like(qs[0])
    for q in qs:
        p1(q, out=out)
    qs = [q.sort().values for q in qs]
    for q in qs:
        p1(q, out=out)
    torch.cuda.synchronize()
if which in ("all", "edm"):
    m = B.EventDrivenMap([bench.BETA], 1000, noNeurons=1024)
    for _ in range(2):
        print(m.ComputeF(bench.Z_DRIVER))
if which == "edm1":     # ONE ring alone: the serial event chain without contention (the floor of small-ensemble runs)
    m = B.EventDrivenMap([bench.BETA], 1, noNeurons=1024)
    for _ in range(3):
        print(m.ComputeF(bench.Z_DRIVER))
if which == "interp2s":   # cell-sorted queries: the straight-line affine / tile kernel serves the call
    plan = B.Interp2Plan(*bench.make_grid())
    g = torch.Generator(device="cuda").manual_seed(2235)
    xq = torch.rand(bench.NQ, generator=g, device="cuda", dtype=torch.float64)
    yq = torch.rand(bench.NQ, generator=g, device="cuda", dtype=torch.float64)
    cell = (xq * 4095).floor().to(torch.int64) * 4096 + (yq * 4095).floor().to(torch.int64)
    o = cell.argsort(); del cell
    xq, yq = xq[o].contiguous(), yq[o].contiguous(); del o
    zq = torch.empty_like(xq)
    for _ in range(3):
        plan.scattered(xq, yq, out=zq)
    torch.cuda.synchronize()
