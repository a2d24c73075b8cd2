"""Generates the committed golden fixtures from the CPU oracle (oracle/liboracle.so).

    python tests/golden/make_golden.py

The reference ships no tests or vectors and cannot be built here, so these
fixtures pin OUR restatement (regression) and record the only independent pins there are:
  * the survey's separate NumPy emulation of one realisation,
    written before the oracle existed, with I = 0.9 and beta = 13.0589 as doubles;
  * numpy.interp / scipy RegularGridInterpolator second opinions (checked in tests, not stored).
Inputs are stored with the outputs so the fixtures do not depend on numpy's RNG stream.
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle_py as O  # noqa: E402


def interp_cases():
    rng = np.random.default_rng(20261018)
    out = {}
    for tag, dt in (("f64", np.float64), ("f32", np.float32)):
        # 1-D: uniform knots, general knots, two knots; queries with knot hits, ends, out of range, NaN
        for kind in ("uniform", "general", "two"):
            if kind == "uniform":
                xg = np.linspace(-1.0, 2.0, 257)
            elif kind == "general":
                xg = np.cumsum(0.5 + rng.random(300))
            else:
                xg = np.array([0.25, 0.75])
            xg = np.unique(xg.astype(dt))
            yg = (np.sin(3 * xg) + 0.1 * rng.standard_normal(xg.size)).astype(dt)
            lo, hi = float(xg[0]), float(xg[-1])
            xi = rng.uniform(lo - 0.1 * (hi - lo), hi + 0.1 * (hi - lo), 500).astype(dt)
            xi[:8] = [xg[0], xg[-1], np.nan, xg[1], xg[-2], np.nextafter(xg[0], dt(-np.inf)),
                      np.nextafter(xg[-1], dt(np.inf)), xg[xg.size // 2]]
            yi, idx = O.interp1(xg, yg, xi, extrap=-7.0)
            yi_scan, idx_scan = O.interp1(xg, yg, np.sort(xi), extrap=-7.0, scan=True)
            out[f"i1_{tag}_{kind}_xg"] = xg
            out[f"i1_{tag}_{kind}_yg"] = yg
            out[f"i1_{tag}_{kind}_xi"] = xi
            out[f"i1_{tag}_{kind}_yi"] = yi
            out[f"i1_{tag}_{kind}_idx"] = idx
        # 2-D
        x = np.unique(np.sort(rng.random(37)).astype(dt))
        y = np.linspace(-1, 1, 29).astype(dt)
        z = rng.standard_normal((y.size, x.size)).astype(dt)
        xq = rng.uniform(x[0] - 0.05, x[-1] + 0.05, 400).astype(dt)
        yq = rng.uniform(-1.1, 1.1, 400).astype(dt)
        xq[:4] = [x[0], x[-1], np.nan, x[3]]
        yq[:4] = [y[0], y[-1], 0.0, np.nan]
        # corner cases that tell the two pass orders apart (finite extrap): NaN in one coordinate and out of
        # range in the other -> the LAST pass decides; out of range in the first-pass coordinate only -> the
        # second pass blends extrap with itself
        xq[4:10] = [np.nan, x[0] - 1, x[0] - 1, x[5], np.nan, x[-1] + 1]
        yq[4:10] = [y[0] - 1, np.nan, 0.3, y[-1] + 1, y[-1] + 1, y[0] - 1]
        xi = rng.uniform(x[0] - 0.05, x[-1] + 0.05, 23).astype(dt)
        yi = rng.uniform(-1.1, 1.1, 31).astype(dt)
        out[f"i2_{tag}_x"] = x
        out[f"i2_{tag}_y"] = y
        out[f"i2_{tag}_z"] = z
        out[f"i2_{tag}_xq"] = xq
        out[f"i2_{tag}_yq"] = yq
        xi[:3] = [np.nan, x[0] - 1, x[2]]
        yi[:3] = [y[0] - 1, np.nan, y[4]]
        out[f"i2_{tag}_zq"] = O.interp2_scattered(x, y, z, xq, yq, extrap=3.5)              # default order: X then Y
        out[f"i2_{tag}_zq_yx"] = O.interp2_scattered(x, y, z, xq, yq, extrap=3.5, y_first=True)
        out[f"i2_{tag}_xi"] = xi
        out[f"i2_{tag}_yi"] = yi
        out[f"i2_{tag}_zi"] = np.ascontiguousarray(O.interp2_grid(x, y, z, xi, yi, extrap=np.nan))
        out[f"i2_{tag}_zi_yx"] = np.ascontiguousarray(O.interp2_grid(x, y, z, xi, yi, extrap=np.nan, y_first=True))
        out[f"i2_{tag}_zi_e"] = np.ascontiguousarray(O.interp2_grid(x, y, z, xi, yi, extrap=3.5))
        out[f"i2_{tag}_zi_e_yx"] = np.ascontiguousarray(O.interp2_grid(x, y, z, xi, yi, extrap=3.5, y_first=True))
    return out


def edm_cases():
    f32 = np.float32
    z_driver = [float(f32(0.3310)), float(f32(0.6914)), float(f32(1.3557))]  # Driver.cu:24
    cases = []

    def run(name, z, **kw):
        cfg = O.edm_cfg(**kw)
        f, a = O.edm_compute_f(cfg, np.array(z), nthreads=4)
        cases.append(dict(name=name, cfg={k: (float(v) if isinstance(v, float) else int(v)) for k, v in kw.items()},
                          z=list(map(float, z)), f=f.tolist(), init_index=a["init_index"].tolist(),
                          event_count=a["event_count"].tolist(), last_index=a["last_index"].tolist(),
                          crossed_index=a["crossed_index"].tolist(), last_time=a["last_time"].tolist(),
                          crossed_time=a["crossed_time"].tolist(), position=a["position"].tolist(),
                          accept=a["accept"].tolist(), mean=a["mean"].tolist(),
                          total_candidates=int(a["total_candidates"]), total_newton_its=int(a["total_newton_its"]),
                          lift_v_head=a["lift_v"][::64].tolist(), lift_s_head=a["lift_s"][::64].tolist(),
                          coupling_head=a["coupling"][::64].tolist()))

    run("driver_N1024_f64", z_driver, R=2, N=1024)
    run("driver_N512_f64", z_driver, R=2, N=512)
    run("driver_N1024_f32", z_driver, R=2, N=1024, precision=1)
    run("survey_emulation_N1024", [0.3310, 0.6914, 1.3557], R=1, N=1024, I=0.9, beta=13.0589)
    run("survey_emulation_N512", [0.3310, 0.6914, 1.3557], R=1, N=512, I=0.9, beta=13.0589)
    run("hetero_sigma0.5", z_driver, R=4, N=1024, sigma=0.5, seed=42)
    run("perturbed_c", [z_driver[0] + 1e-2, z_driver[1], z_driver[2]], R=1, N=1024)
    run("two_fronts", z_driver[:2], R=1, N=512, M=2)
    run("four_fronts", z_driver + [2.2], R=1, N=1024, M=4)
    run("four_fronts_quiet", z_driver + [2.05], R=1, N=1024, M=4)   # ring goes quiet: not accepted, F = NaN
    run("short_horizon", z_driver, R=1, N=1024, time_horizon=1.0)
    run("quirk_accept0", z_driver, R=3, N=256, quirks=1)
    # the survey's independent emulation values for the two emulation cases
    kat = dict(N1024=dict(init_index=[512, 472, 435], events=848, last_index=[793, 754, 717],
                          crossed_index=[794, 755, 718], X_T=[1.64980178, 1.42279308, 1.20343921],
                          F=[5.19822e-3, 3.35352e-3, 2.82409e-3], normF=6.80023e-3),
               N512=dict(events=421, normF=3.94916e-2))
    cfg = O.edm_cfg(R=1, N=1024)
    J, f0 = O.edm_compute_dfdu(cfg, np.array(z_driver), 1e-2, nthreads=4)
    normal = [O.normal(42, i) for i in range(8)]
    return dict(cases=cases, survey_kat=kat, jacobian_driver_N1024=dict(J=J.tolist(), f0=f0.tolist(), eps=1e-2),
                normal_seed42=normal)


if __name__ == "__main__":
    np.savez_compressed(os.path.join(HERE, "interp_golden.npz"), **interp_cases())
    with open(os.path.join(HERE, "edm_golden.json"), "w") as fh:
        json.dump(edm_cases(), fh, indent=1)
    print("wrote", os.listdir(HERE))
