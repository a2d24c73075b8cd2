"""interp1 at BASELINE configs[0] size (1e6 knots, 1e7 queries), 8 rotating buffer pairs as in bench.py:
uniform / non-uniform knots x sorted / unsorted queries, for a sweep of launch shapes (B200_INTERP1_GRID_MULT)
and with / without programmatic dependent launch (B200_INTERP1_PDL).  One subprocess per setting (the env is read once)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import os, sys
sys.path.insert(0, %r)
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
ng, ni, nbuf = 1_000_000, 10_000_000, 8
rng = np.random.default_rng(1234)
res = []
for kind in ("uniform", "nonuniform"):
    xg = np.linspace(0.0, 1.0, ng) if kind == "uniform" else np.cumsum(0.5 + rng.random(ng))
    xg = (xg - xg[0]) / (xg[-1] - xg[0]); yg = np.sin(2 * np.pi * xg)
    p1 = B.Interp1Plan(xg, yg)
    g1 = torch.Generator(device="cuda").manual_seed(1236)
    for order in ("unsorted", "sorted"):
        qs = [torch.rand(ni, generator=g1, device="cuda", dtype=torch.float64) for _ in range(nbuf)]
        if order == "sorted": qs = [q.sort().values for q in qs]
        outs = [torch.empty_like(q) for q in qs]
        for i in range(16): p1(qs[i %% nbuf], out=outs[i %% nbuf])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(200): p1(qs[i %% nbuf], out=outs[i %% nbuf])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 200
        res.append(f"{kind[:4]}/{order[:4]} {ms * 1e3:5.1f} us ({(16 * ni + 16 * ng) / ms / 1e6 / 6537:.2f})")
        del qs, outs
    p1.close()
print("  ".join(res))
''' % ROOT
settings = [dict(B200_INTERP1_PDL="0"), dict(B200_INTERP1_PDL="1")] + [dict(B200_INTERP1_GRID_MULT=str(m)) for m in (4, 8, 12, 24, 32, 64)]
if len(sys.argv) > 1 and sys.argv[1] == "quick":
    settings = settings[:2]
for st in settings:
    out = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **st), capture_output=True, text=True)
    print(f"{str(st):38s} {out.stdout.strip()} {out.stderr[-300:] if out.returncode else ''}", flush=True)
