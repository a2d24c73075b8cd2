"""interp2 tensor grid (1e4 x 1e4 outputs on the 4096^2 f64 grid): rows-per-thread / CTA-size sweep for both pass
orders (B200_INTERP2_GRID_V, B200_INTERP2_GRID_THREADS are read once per process: one subprocess per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, os
sys.path.insert(0, %r)
import numpy as np, torch
from armadillocudalinearinterpolation_b200 import _lib
if os.environ.get("B200_AB_LIB"): _lib.LIB_PATH = os.environ["B200_AB_LIB"]   # A/B of another build of the library
import armadillocudalinearinterpolation_b200 as B, bench
grid = bench.make_grid()
g2 = torch.Generator(device="cuda").manual_seed(2236)
xi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
yi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
out = []
dt = torch.float32 if os.environ.get("GRID_F32") else torch.float64
if dt == torch.float32:
    grid = tuple(np.asarray(a, np.float32) for a in grid); xi = xi.float().sort().values; yi = yi.float().sort().values
esz = 4 if dt == torch.float32 else 8
for name, flags in (("xy", 0), ("yx", B.Interp2Plan.ORDER_YX)):
    plan = B.Interp2Plan(*grid, flags=flags)
    for _ in range(5): plan.grid(xi, yi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.grid(xi, yi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out.append(f"{name} {ms:.4f} ms ({(esz * 1e8 + esz * 4096 * 4096) / ms / 1e6 / 6537:.2f})")
    plan.close()
print("  ".join(out))
''' % ROOT
quick = len(sys.argv) > 1 and sys.argv[1] == "quick"
for v in ((4,) if quick else (4, 2, 1)):
    for th in ((128, 64) if quick else (128, 256, 64)):
        env = dict(os.environ, B200_INTERP2_GRID_V=str(v), B200_INTERP2_GRID_THREADS=str(th))
        for f32 in ((0, 1) if quick else (0,)):
            if f32: env["GRID_F32"] = "1"
            r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
            print(f"V={v} threads={th} {'f32' if f32 else 'f64'}: {r.stdout.strip()} {r.stderr[-200:] if r.returncode else ''}", flush=True)
