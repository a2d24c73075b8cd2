"""First GPU sanity run: interp1/interp2/edm vs the oracle."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
from oracle import oracle_py as O

print("devices", B.device_count())
rng = np.random.default_rng(0)
for dt in (np.float64, np.float32):
    for kind in ("uniform", "general"):
        ng = 100001
        xg = np.linspace(0, 1, ng) if kind == "uniform" else np.cumsum(0.5 + rng.random(ng)); xg = (xg - xg[0]) / (xg[-1] - xg[0])
        xg = xg.astype(dt); xg = np.unique(xg)
        yg = (np.sin(2 * np.pi * xg) + 0.1 * rng.standard_normal(xg.size)).astype(dt)
        xi = rng.uniform(-0.01, 1.01, 1000003).astype(dt)
        xi[:5] = [xg[0], xg[-1], np.nan, xg[5], xg[-2]]
        plan = B.Interp1Plan(xg, yg)
        y, idx = plan(xi, return_index=True)
        yo, io = O.interp1(xg, yg, xi, nthreads=8)
        ok = np.array_equal(y.view(np.uint64 if dt == np.float64 else np.uint32), yo.view(np.uint64 if dt == np.float64 else np.uint32))
        nan_ok = np.array_equal(np.isnan(y), np.isnan(yo))
        vals_ok = np.array_equal(y[~np.isnan(y)], yo[~np.isnan(yo)])
        print(dt.__name__, kind, "mode", plan.lookup_mode, "bits", ok, "nan", nan_ok, "vals", vals_ok, "idx", np.array_equal(idx, io))
        y1 = B.interp1(xg, yg, xi)
        print("   oneshot", np.array_equal(y1[~np.isnan(y1)], yo[~np.isnan(yo)]))
    nx, ny = 300, 257
    x = np.sort(rng.random(nx)).astype(dt); y_ = np.linspace(-1, 2, ny).astype(dt)
    z = rng.standard_normal((ny, nx)).astype(dt)
    p2 = B.Interp2Plan(x, y_, z)
    xq = rng.uniform(x[0] - 0.05, x[-1] + 0.05, 200001).astype(dt); yq = rng.uniform(-1.1, 2.1, 200001).astype(dt)
    xq[:3] = [x[0], x[-1], np.nan]; yq[:3] = [y_[0], y_[-1], 0.0]; yq[3] = np.nan
    zq = p2.scattered(xq, yq, extrap=7.5)
    zo = O.interp2_scattered(x, y_, z, xq, yq, extrap=7.5, nthreads=8)
    print(dt.__name__, "interp2 scattered", np.array_equal(np.isnan(zq), np.isnan(zo)), np.array_equal(zq[~np.isnan(zq)], zo[~np.isnan(zo)]))
    xi = np.sort(rng.uniform(x[0] - 0.05, x[-1] + 0.05, 777)).astype(dt); yi = rng.uniform(-1.1, 2.1, 1001).astype(dt)
    zi = p2.grid(xi, yi, extrap=np.nan)
    zo = O.interp2_grid(x, y_, z, xi, yi, nthreads=8)
    print(dt.__name__, "interp2 grid", np.array_equal(np.isnan(zi), np.isnan(zo)), np.array_equal(zi[~np.isnan(zi)], zo[~np.isnan(zo)]))

z0 = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)
for prec in ("f64", "f32"):
    for N in (1024, 512, 200):
        m = B.EventDrivenMap([np.float32(13.0589)], 8, noNeurons=N, precision=prec)
        m.SetDebugFlag(True); m.EnableTiming(True)
        t = time.time(); f = m.ComputeF(z0); dtm = time.time() - t
        cfg = O.edm_cfg(R=8, N=N, precision=0 if prec == "f64" else 1)
        fo, a = O.edm_compute_f(cfg, z0, nthreads=8)
        ev = m.DebugFetch("event_count")
        print(prec, N, "F", f, "oracle", fo, "relerr", np.abs(f - fo).max() / np.abs(fo).max(), "events", ev[0, :3], a["event_count"][:3],
              "idx", np.array_equal(m.DebugFetch("last_index")[0], a["last_index"]), np.array_equal(m.DebugFetch("crossed_index")[0], a["crossed_index"]),
              "lift", np.nanmax(np.abs(m.DebugFetch("lift_v")[0] - a["lift_v"])), np.nanmax(np.abs(m.DebugFetch("lift_s")[0] - a["lift_s"])),
              "ms", m.LastEvolveMs(), m.LastCounters(), "wall", dtm)
# heterogeneous
m = B.EventDrivenMap([np.float32(13.0589)], 8, noNeurons=1024)
m.SetDebugFlag(True); m.SetParameterStdDev(0.5); m.SetSeed(42)
f = m.ComputeF(z0)
beta = m.DebugFetch("beta")
cfg = O.edm_cfg(R=8, N=1024, sigma=0.5, seed=42)
print("beta rng maxdiff", np.abs(beta - O.edm_beta(cfg)).max())
cfg = O.edm_cfg(R=8, N=1024, beta_ext=beta)
fo, a = O.edm_compute_f(cfg, z0, nthreads=8)
print("hetero F", f, fo, np.abs(f - fo).max(), m.DebugFetch("event_count")[0], a["event_count"], m.DebugFetch("accept")[0])
# Jacobian
m = B.EventDrivenMap([np.float32(13.0589)], 4, noNeurons=1024)
J, f0 = m.ComputeDFDU(z0, 1e-2, return_f0=True)
cfg = O.edm_cfg(R=4, N=1024)
Jo, f0o = O.edm_compute_dfdu(cfg, z0, 1e-2, nthreads=8)
print("J", J, "\nJo", Jo, "\nmax rel", np.abs(J - Jo).max() / np.abs(Jo).max())
# default ensemble timing
m = B.EventDrivenMap([np.float32(13.0589)], 1000, noNeurons=1024)
m.EnableTiming(True)
for npt in (0, 4, 8, 16, 2):
    m.SetTuning(npt)
    for _ in range(2):
        f = m.ComputeF(z0)
    print("R=1000 N=1024 npt", npt, "evolve ms", m.LastEvolveMs(), m.LastCounters(), f)
