"""e2e with pageable host buffers: worker-thread / chunk-size sweep of the pinned-ring staging path."""
import os, subprocess, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = r'''
import sys, time, numpy as np
sys.path.insert(0, %r)
import bench, armadillocudalinearinterpolation_b200 as B
B.set_device(0)
plan = B.Interp2Plan(*bench.make_grid())
rng = np.random.default_rng(2235)
n = 100_000_000
x = rng.random(n); y = rng.random(n); z = np.empty(n)
plan.scattered(x, y, out=z)
t = time.perf_counter()
for _ in range(3): plan.scattered(x, y, out=z)
print((time.perf_counter() - t) / 3 * 1e3)
''' % ROOT
for th in (4, 8, 12, 16, 24):
    for ch in (19, 20, 21):
        env = dict(os.environ, B200_STAGE_THREADS=str(th), B200_STAGE_CHUNK_LOG2=str(ch))
        out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
        print(f"threads {th:2d} chunk 2^{ch}: {out.stdout.strip()} ms {out.stderr[-200:] if out.returncode else ''}", flush=True)
