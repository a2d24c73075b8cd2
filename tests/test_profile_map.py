"""Profile map (BASELINE config 5: "1e3-dim coarse profile") — NEW functionality, not in the
reference.  CPU: the oracle's definition behaves like a coarse time-stepper (the travelling wave
keeps firing, the homogeneous rest state is a fixed point).  GPU: parity with the oracle."""
import numpy as np
import pytest

Z_DRIVER = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)
BETA = float(np.float32(13.0589))


def coarse_wave(oracle, N, nc, **kw):
    """The analytic travelling-wave lift (LiftKernel) sampled at the coarse knots."""
    cfg = oracle.edm_cfg(R=1, N=N, **kw)
    v, s = oracle.edm_lift(cfg, Z_DRIVER)
    xf = -3.0 + 6.0 / N * np.arange(N)
    xc = -3.0 + 6.0 / nc * np.arange(nc)
    return np.concatenate([np.interp(xc, xf, v), np.interp(xc, xf, s)])


def test_oracle_profile_map_sustains_the_wave(oracle):
    N, nc = 1024, 512
    u = coarse_wave(oracle, N, nc)
    f, a = oracle.profile_compute_f(oracle.edm_cfg(R=2, N=N, time_horizon=1.0), nc, u)
    assert a["accept"].tolist() == [1, 1]
    assert 120 <= a["event_count"][0] <= 200          # ~170 events per unit time in the reference map
    assert np.all(np.isfinite(f))
    # a horizon twice as long sees about twice as many events
    _, a2 = oracle.profile_compute_f(oracle.edm_cfg(R=1, N=N, time_horizon=2.0), nc, u)
    assert 1.7 < a2["event_count"][0] / a["event_count"][0] < 2.3


def test_oracle_rest_state_is_a_fixed_point(oracle):
    """v = I, s = 0 everywhere: nobody fires, v stays at I, F = 0."""
    N, nc = 256, 32
    I = float(np.float32(0.9))
    u = np.concatenate([np.full(nc, I), np.zeros(nc)])
    f, a = oracle.profile_compute_f(oracle.edm_cfg(R=1, N=N, time_horizon=0.7), nc, u)
    assert a["event_count"][0] == 0 and np.max(np.abs(f)) < 1e-15


def test_oracle_lift_restrict_roundtrip(oracle):
    """T -> 0: restriction of the lift returns the coarse profile when coarse knots are grid points."""
    N, nc = 512, 64
    rng = np.random.default_rng(0)
    u = np.concatenate([0.5 * rng.random(nc), 0.01 * rng.random(nc)])
    f, a = oracle.profile_compute_f(oracle.edm_cfg(R=1, N=N, time_horizon=1e-12), nc, u)
    assert a["event_count"][0] == 0 and np.max(np.abs(f)) < 1e-11


@pytest.mark.gpu
@pytest.mark.parametrize("N,nc,T,sigma", [(1024, 512, 1.0, 0.0), (1024, 500, 0.5, 0.0), (512, 64, 1.0, 0.3), (2048, 512, 0.5, 0.0), (1000, 37, 0.4, 0.0)])
def test_gpu_profile_map_matches_oracle(b200, oracle, N, nc, T, sigma):
    R = 3
    u = coarse_wave(oracle, N, nc)
    m = b200.EventDrivenMap([BETA], R, noNeurons=N)
    m.SetTimeHorizon(T); m.SetProfileMode(nc); m.SetDebugFlag(True)
    if sigma:
        m.SetParameterStdDev(sigma); m.SetSeed(9)
    f = m.ComputeF(u)
    beta = m.DebugFetch("beta")
    fo, a = oracle.profile_compute_f(oracle.edm_cfg(R=R, N=N, time_horizon=T, beta_ext=beta if sigma else None), nc, u, nthreads=3)
    assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
    assert np.array_equal(m.DebugFetch("accept")[0], a["accept"])
    lift_v = m.DebugFetch("lift_v")[0]
    cfg1 = oracle.edm_cfg(R=1, N=N, time_horizon=1e-12)
    # the lift is the same interp1 rule on both sides: bit-identical initial state
    restricted = m.DebugFetch("position")[0]
    assert np.max(np.abs(restricted - a["restricted"])) < 1e-10 * max(1.0, np.max(np.abs(a["restricted"])))
    assert np.max(np.abs(f - fo)) < 1e-10 * max(1.0, np.max(np.abs(a["restricted"])))
    assert lift_v.shape == (N,)


@pytest.mark.gpu
def test_gpu_profile_jacobian_and_mode_switch(b200, oracle):
    N, nc, R = 256, 8, 2
    u = coarse_wave(oracle, N, nc)
    m = b200.EventDrivenMap([BETA], R, noNeurons=N)
    m.SetTimeHorizon(0.5); m.SetProfileMode(nc)
    J, f0 = m.ComputeDFDU(u, 1e-3, return_f0=True)
    assert J.shape == (2 * nc, 2 * nc)
    cfg = oracle.edm_cfg(R=R, N=N, time_horizon=0.5)
    fo, _ = oracle.profile_compute_f(cfg, nc, u)
    assert np.max(np.abs(f0 - fo)) < 1e-10
    for i in (0, 5, 11):
        du = u.copy(); du[i] += 1e-3
        fi, _ = oracle.profile_compute_f(cfg, nc, du)
        assert np.max(np.abs(J[:, i] - (fi - fo) / 1e-3)) < 1e-6 * max(1.0, np.max(np.abs(J[:, i])))
    assert np.array_equal(m.ComputeDFDU(u, 1e-3, f0=m.ComputeF(u)), J)     # given-F form: same bits
    with pytest.raises(b200.B200Error):
        m.ComputeF(Z_DRIVER)                     # wrong length in profile mode
    m.SetProfileMode(0)                          # back to the reference's front map
    m.SetTimeHorizon(5.0)
    f = m.ComputeF(Z_DRIVER)
    fo, _ = oracle.edm_compute_f(oracle.edm_cfg(R=R, N=N), Z_DRIVER)
    assert np.array_equal(np.isnan(f), np.isnan(fo)) and np.allclose(f[~np.isnan(f)], fo[~np.isnan(fo)], rtol=0, atol=1e-10)
