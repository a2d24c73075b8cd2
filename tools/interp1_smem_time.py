"""interp1 small-grid shared-memory path: 1e3 knots, 1e8 queries (bench entry interp1_f64_1e3knots_1e8queries_smem)."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
NQ = 100_000_000
gs = torch.Generator(device="cuda").manual_seed(77)
q = torch.rand(NQ, generator=gs, device="cuda", dtype=torch.float64) * 6.0 - 3.0
o = torch.empty_like(q)
ref = None
for affine in ("1", "0"):
    os.environ["B200_INTERP_AFFINE"] = affine
    for name, xg in (("linspace", np.linspace(-3.0, 3.0, 1000)), ("cumsum", None)):
        if xg is None:
            r = np.random.default_rng(3); xg = np.cumsum(0.5 + r.random(1000)); xg = (xg - xg[0]) / (xg[-1] - xg[0]) * 6.0 - 3.0
        p1 = B.Interp1Plan(xg, np.sin(xg))
        for _ in range(3): p1(q, out=o)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10): p1(q, out=o)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"affine={affine} {name}: mode {p1.lookup_mode} {ms:.4f} ms  {16 * NQ / ms / 1e6:.0f} GB/s  frac {16 * NQ / ms / 1e6 / 6537:.3f}  sum {float(o.sum()):.6f}")
