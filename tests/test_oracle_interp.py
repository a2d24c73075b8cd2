"""CPU tests of the interpolation oracle: golden fixtures, Armadillo's scan vs the
order-independent search, numpy / scipy second opinions, edge cases."""
import os

import numpy as np
import pytest

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "interp_golden.npz"))


def same_bits(a, b):
    a = np.ascontiguousarray(a); b = np.ascontiguousarray(b)
    return a.shape == b.shape and np.array_equal(np.isnan(a), np.isnan(b)) and np.array_equal(a[~np.isnan(a)], b[~np.isnan(b)])


@pytest.mark.parametrize("tag", ["f64", "f32"])
@pytest.mark.parametrize("kind", ["uniform", "general", "two"])
def test_interp1_golden(oracle, tag, kind):
    g = lambda k: GOLD[f"i1_{tag}_{kind}_{k}"]
    yi, idx = oracle.interp1(g("xg"), g("yg"), g("xi"), extrap=-7.0)
    assert same_bits(yi, g("yi"))
    assert np.array_equal(idx, g("idx"))


@pytest.mark.parametrize("tag", ["f64", "f32"])
def test_interp2_golden(oracle, tag):
    g = lambda k: GOLD[f"i2_{tag}_{k}"]
    zq = oracle.interp2_scattered(g("x"), g("y"), g("z"), g("xq"), g("yq"), extrap=3.5)
    assert same_bits(zq, g("zq"))
    zi = oracle.interp2_grid(g("x"), g("y"), g("z"), g("xi"), g("yi"))
    assert same_bits(zi, g("zi"))


@pytest.mark.parametrize("dt", [np.float64, np.float32])
def test_scan_equals_search(oracle, dt):
    """Armadillo's monotone nearest-knot scan (sorted XI) and the binary search give the same
    bracket and the same bits."""
    rng = np.random.default_rng(1)
    xg = np.unique(np.cumsum(0.5 + rng.random(2000)).astype(dt))
    yg = rng.standard_normal(xg.size).astype(dt)
    xi = np.sort(rng.uniform(xg[0] - 1, xg[-1] + 1, 20000).astype(dt))
    xi[100:110] = xg[50:60]  # exact knot hits
    xi = np.sort(xi)
    a, ia = oracle.interp1(xg, yg, xi, extrap=0.5, scan=True)
    b, ib = oracle.interp1(xg, yg, xi, extrap=0.5, scan=False, nthreads=4)
    assert same_bits(a, b) and np.array_equal(ia, ib)


def test_interp1_vs_numpy(oracle):
    rng = np.random.default_rng(2)
    xg = np.cumsum(0.5 + rng.random(5000)); yg = np.sin(xg)
    xi = rng.uniform(xg[0], xg[-1], 50000)
    yi, idx = oracle.interp1(xg, yg, xi)
    ref = np.interp(xi, xg, yg)
    assert np.max(np.abs(yi - ref)) < 1e-13
    assert np.array_equal(idx, np.clip(np.searchsorted(xg, xi, side="right") - 1, 0, xg.size - 1))


def test_interp2_vs_scipy(oracle):
    from scipy.interpolate import RegularGridInterpolator
    rng = np.random.default_rng(3)
    x = np.sort(rng.random(40)); y = np.sort(rng.random(50)); z = rng.standard_normal((50, 40))
    xq = rng.uniform(x[0], x[-1], 5000); yq = rng.uniform(y[0], y[-1], 5000)
    zq = oracle.interp2_scattered(x, y, z, xq, yq)
    ref = RegularGridInterpolator((y, x), z)(np.stack([yq, xq], 1))
    assert np.max(np.abs(zq - ref)) < 1e-12
    zi = oracle.interp2_grid(x, y, z, xq[:30], yq[:20])
    ref = RegularGridInterpolator((y, x), z)(np.stack(np.meshgrid(yq[:20], xq[:30], indexing="ij"), -1))
    assert np.max(np.abs(zi - ref)) < 1e-12


def test_edges_and_errors(oracle):
    xg = np.array([0.0, 1.0, 3.0]); yg = np.array([10.0, 20.0, 40.0])
    xi = np.array([0.0, 3.0, -1e-9, 3.0000001, np.nan, 1.0, 0.5, 2.0])
    yi, idx = oracle.interp1(xg, yg, xi, extrap=-1.0)
    assert yi[0] == 10.0 and idx[0] == 0
    assert yi[1] == 40.0 and idx[1] == 2          # last knot: bracket (n-1, n-1), w = 0
    assert yi[2] == -1.0 and idx[2] == -1 and yi[3] == -1.0
    assert np.isnan(yi[4]) and idx[4] == -1
    assert yi[5] == 20.0 and idx[5] == 1 and yi[6] == 15.0 and yi[7] == 30.0
    empty, _ = oracle.interp1(xg, yg, np.array([]))
    assert empty.size == 0
    with pytest.raises(ValueError):
        oracle.interp1(np.array([0.0, 0.0, 1.0]), yg, xi)      # not strictly ascending
    with pytest.raises(ValueError):
        oracle.interp1(np.array([0.0]), np.array([1.0]), xi)   # fewer than two knots
