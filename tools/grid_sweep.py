"""interp2 tensor grid (1e4 x 1e4 outputs on the 4096^2 f64 grid): rows-per-thread / CTA-size sweep for both pass
orders (B200_INTERP2_GRID_V, B200_INTERP2_GRID_THREADS are read once per process: one subprocess per setting)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B, bench
grid = bench.make_grid()
g2 = torch.Generator(device="cuda").manual_seed(2236)
xi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
yi = torch.rand(10_000, generator=g2, device="cuda", dtype=torch.float64).sort().values
out = []
for name, flags in (("xy", 0), ("yx", B.Interp2Plan.ORDER_YX)):
    plan = B.Interp2Plan(*grid, flags=flags)
    for _ in range(5): plan.grid(xi, yi)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): plan.grid(xi, yi)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    out.append(f"{name} {ms:.4f} ms ({(8e8 + 8 * 4096 * 4096) / ms / 1e6 / 6537:.2f})")
    plan.close()
print("  ".join(out))
''' % ROOT
for v in (4, 2, 1):
    for th in (128, 256, 64):
        env = dict(os.environ, B200_INTERP2_GRID_V=str(v), B200_INTERP2_GRID_THREADS=str(th))
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"V={v} threads={th}: {r.stdout.strip()} {r.stderr[-200:] if r.returncode else ''}", flush=True)
