"""Headline interp2 kernel A/B: generic shared-memory kernel (B200_INTERP2_FAST=0) vs the straight-line affine/tile
kernel always (B200_INTERP2_FAST=2) and chosen per call by the device-side locality probe (=1, the default), unsorted and cell-sorted queries; bits compared."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys
sys.path.insert(0, %r)
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B, bench
plan = B.Interp2Plan(*bench.make_grid())
g = torch.Generator(device="cuda").manual_seed(2235)
n = bench.NQ
xq = torch.rand(n, generator=g, device="cuda", dtype=torch.float64); yq = torch.rand(n, generator=g, device="cuda", dtype=torch.float64)
zq = torch.empty_like(xq)
def t(x, y, reps=10):
    for _ in range(3): plan.scattered(x, y, out=zq)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): plan.scattered(x, y, out=zq)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
a = t(xq, yq)
chk = int(zq.view(torch.int64).sum().item())
cell = (xq * 4095).floor().to(torch.int64) * 4096 + (yq * 4095).floor().to(torch.int64)
o = cell.argsort(); del cell
xs, ys = xq[o], yq[o]; del o
b = t(xs, ys)
print(f"unsorted {a:.4f} ms  cell-sorted {b:.4f} ms  checksum {chk}")
''' % ROOT
for st in (dict(B200_INTERP2_FAST="0"), dict(B200_INTERP2_FAST="2"), dict(B200_INTERP2_FAST="1")):
    r = subprocess.run([sys.executable, "-c", code], env=dict(os.environ, **st), capture_output=True, text=True)
    print(f"{str(st):32s} {r.stdout.strip()} {r.stderr[-300:] if r.returncode else ''}", flush=True)
