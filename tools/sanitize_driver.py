"""Small invocations of every kernel family, for compute-sanitizer (racecheck / synccheck / memcheck) and for
NVTX-annotated traces.  Sizes are tiny on purpose: the sanitizer serialises and instruments every access.
    compute-sanitizer --tool racecheck python tools/sanitize_driver.py
Each result is still checked against the oracle, so a run under the tool is also a parity run."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import armadillocudalinearinterpolation_b200 as B  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

which = set(sys.argv[1:]) or {"edm", "profile", "interp1", "interp2"}
rng = np.random.default_rng(3)
BETA = float(np.float32(13.0589))
Z = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], np.float64)


def same_bits(a, b):
    return np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


if "edm" in which:
    # front map: homogeneous and heterogeneous ensembles, two launch shapes (128- and 256-thread CTAs), Jacobian batch
    for N, R, sigma, T in ((256, 3, 0.0, 1.0), (256, 3, 0.4, 1.0), (1024, 2, 0.0, 0.6), (2048, 2, 0.3, 0.4)):
        m = B.EventDrivenMap([BETA], R, noNeurons=N)
        m.SetModel(time_horizon=T); m.SetParameterStdDev(sigma); m.SetSeed(9); m.SetDebugFlag(True)
        f = m.ComputeF(Z)
        fo, a = O.edm_compute_f(O.edm_cfg(R=R, N=N, sigma=sigma, seed=9, time_horizon=T), Z)
        assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"]), "event sequence differs"
        assert np.max(np.abs(f - fo)) < 1e-9
        J = m.ComputeDFDU(Z, 1e-2)
        assert np.all(np.isfinite(J))
        print(f"edm N={N} R={R} sigma={sigma}: events {a['event_count'].tolist()} ok", flush=True)
        m.close()
if "profile" in which:
    for sigma in (0.0, 0.3):
        nc, R, N = 32, 2, 256
        m = B.EventDrivenMap([BETA], R, noNeurons=N)
        m.SetTimeHorizon(0.5); m.SetParameterStdDev(sigma); m.SetSeed(4); m.SetProfileMode(nc)
        u = np.concatenate([0.2 + 0.7 * np.sin(np.linspace(0, np.pi, nc)) ** 2, np.linspace(0.0, 0.4, nc)])
        f = m.ComputeF(u)
        fo, _ = O.profile_compute_f(O.edm_cfg(R=R, N=N, sigma=sigma, seed=4, time_horizon=0.5), nc, u)
        assert np.max(np.abs(f - fo)) < 1e-9
        print(f"profile map sigma={sigma} ok", flush=True)
        m.close()
if "interp1" in which:
    for kind in ("linspace", "cumsum", "small"):
        n = 1000 if kind == "small" else 50_000
        xg = np.linspace(0, 1, n) if kind != "cumsum" else np.cumsum(0.5 + rng.random(n))
        xg = (xg - xg[0]) / (xg[-1] - xg[0]); yg = np.sin(7 * xg)
        xi = rng.uniform(-0.01, 1.01, 40_003); xi[:3] = [0.0, 1.0, np.nan]
        yi, idx = B.Interp1Plan(xg, yg)(xi, extrap=-1.0, return_index=True)
        yo, io = O.interp1(xg, yg, xi, extrap=-1.0)
        assert same_bits(yi, yo) and np.array_equal(idx, io)
        print(f"interp1 {kind} ok", flush=True)
if "interp2" in which:
    import torch
    x = np.linspace(0, 1, 200); y = np.cumsum(0.5 + rng.random(150)); z = rng.standard_normal((150, 200))
    xq = rng.uniform(-0.02, 1.02, 30_001); yq = rng.uniform(y[0] - 0.5, y[-1] + 0.5, 30_001)
    ref = O.interp2_scattered(x, y, z, xq, yq, extrap=2.0)
    P = B.Interp2Plan
    for flags in (0, P.FORCE_CELLS, P.FORCE_TILES, P.NO_CELLS | P.NO_TILES):
        assert same_bits(P(x, y, z, flags=flags).scattered(xq, yq, extrap=2.0), ref)
    pb = P(x, y, z, flags=P.FORCE_CELLS | P.FORCE_BANDS)
    zb = pb.scattered(torch.from_numpy(xq).cuda(), torch.from_numpy(yq).cuda(), extrap=2.0)
    torch.cuda.synchronize()
    assert same_bits(zb.cpu().numpy(), ref)
    xi = np.sort(rng.uniform(0, 1, 300)); yi = np.sort(rng.uniform(y[0], y[-1], 200))
    assert same_bits(P(x, y, z).grid(xi, yi), O.interp2_grid(x, y, z, xi, yi))
    assert same_bits(P(x, y, z, flags=P.ORDER_YX).grid(xi, yi), O.interp2_grid(x, y, z, xi, yi, y_first=True))
    print("interp2 scattered (4 layouts + banded) and grid (both orders) ok", flush=True)
print("sanitize_driver: all ok")
