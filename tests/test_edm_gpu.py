"""GPU parity tests of the lift -> evolve -> restrict map through the C-ABI: identical event
sequences (integer outputs bit-exact) and 1e-10 relative agreement of F, positions and the
finite-difference Jacobian with the CPU oracle (north-star tolerance for FP64)."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
RTOL = 1e-10  # BASELINE.json north_star: "within a stated relative tolerance (1e-10 in FP64)"
with open(os.path.join(os.path.dirname(__file__), "golden", "edm_golden.json")) as fh:
    GOLD = json.load(fh)
CASES = {c["name"]: c for c in GOLD["cases"]}
Z_DRIVER = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)
BETA = float(np.float32(13.0589))


def make_map(b200, cfg, **kw):
    m = b200.EventDrivenMap([cfg.get("beta", BETA)], cfg.get("R", 1000), noNeurons=cfg.get("N", 1024),
                            noFronts=cfg.get("M", 3), precision="f32" if cfg.get("precision", 0) else "f64", **kw)
    model = {k: cfg[k] for k in ("I", "time_horizon", "quirks") if k in cfg}
    if model:
        m.SetModel(**model)
    if "sigma" in cfg:
        m.SetParameterStdDev(cfg["sigma"])
    if "seed" in cfg:
        m.SetSeed(cfg["seed"])
    m.SetDebugFlag(True)
    return m


def rel(a, b):
    """max |a - b| / max |b| over the finite entries; NaN/inf patterns must coincide."""
    a = np.asarray(a, float); b = np.asarray(b, float)
    assert a.shape == b.shape and np.array_equal(np.isfinite(a), np.isfinite(b)), (a, b)
    ok = np.isfinite(b)
    if not ok.any():
        return 0.0
    return np.max(np.abs(a[ok] - b[ok])) / max(np.max(np.abs(b[ok])), 1e-300)


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden(b200, name):
    c = CASES[name]
    f32 = c["cfg"].get("precision", 0) == 1
    m = make_map(b200, c["cfg"])
    f = m.ComputeF(np.array(c["z"]))
    assert np.array_equal(m.DebugFetch("init_index")[0], np.array(c["init_index"]))
    if f32:
        # FP32 compatibility arithmetic: CUDA expf/powf vs glibc differ in the last ulp and the
        # map amplifies it; same front cells, F to 1e-3 relative
        assert np.array_equal(m.DebugFetch("last_index")[0], np.array(c["last_index"]))
        assert rel(f, c["f"]) < 1e-3
        return
    assert np.array_equal(m.DebugFetch("event_count")[0], np.array(c["event_count"]))
    assert np.array_equal(m.DebugFetch("last_index")[0], np.array(c["last_index"]))
    assert np.array_equal(m.DebugFetch("crossed_index")[0], np.array(c["crossed_index"]))
    assert np.array_equal(m.DebugFetch("accept")[0], np.array(c["accept"]))
    assert rel(m.DebugFetch("position")[0], c["position"]) < RTOL
    assert rel(m.DebugFetch("last_time")[0], c["last_time"]) < RTOL
    assert rel(m.DebugFetch("crossed_time")[0], c["crossed_time"]) < RTOL
    assert rel(m.DebugFetch("mean")[0], c["mean"]) < RTOL
    if np.all(np.isfinite(c["f"])):
        assert np.max(np.abs(f - np.array(c["f"]))) < RTOL * max(1.0, np.max(np.abs(c["mean"])))
    else:
        assert np.array_equal(np.isnan(f), np.isnan(np.array(c["f"])))
    assert rel(m.DebugFetch("lift_v")[0][::64], c["lift_v_head"]) < RTOL
    assert rel(m.DebugFetch("lift_s")[0][::64], c["lift_s_head"]) < RTOL
    assert rel(m.DebugFetch("coupling")[::64], c["coupling_head"]) < RTOL


@pytest.mark.parametrize("N,npt", [(1024, 0), (1024, 4), (1024, 16), (512, 4), (512, 8), (256, 4), (1000, 8), (96, 4), (2048, 8), (4096, 16)])
def test_vs_oracle_shapes_and_tunings(b200, oracle, N, npt):
    """Every launch shape (neurons per thread, ragged N) gives the oracle's event sequence."""
    R = 3
    m = make_map(b200, dict(R=R, N=N))
    m.SetTuning(npt)
    f = m.ComputeF(Z_DRIVER)
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=R, N=N), Z_DRIVER, nthreads=4)
    assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
    assert np.array_equal(m.DebugFetch("last_index")[0], a["last_index"])
    assert np.array_equal(m.DebugFetch("crossed_index")[0], a["crossed_index"])
    if np.all(np.isfinite(fo)):
        assert rel(m.DebugFetch("position")[0], a["position"]) < RTOL
        assert np.max(np.abs(f - fo)) < RTOL * np.max(np.abs(a["mean"]))
    else:
        assert np.array_equal(np.isnan(f), np.isnan(fo))


def test_heterogeneous_ensemble_and_rng(b200, oracle):
    R, N = 6, 1024
    m = make_map(b200, dict(R=R, N=N, sigma=0.5, seed=42))
    f = m.ComputeF(Z_DRIVER)
    beta = m.DebugFetch("beta")
    cfg = oracle.edm_cfg(R=R, N=N, sigma=0.5, seed=42)
    assert np.max(np.abs(beta - oracle.edm_beta(cfg))) < 1e-13      # same counter-based generator
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=R, N=N, beta_ext=beta), Z_DRIVER, nthreads=4)
    assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
    assert np.array_equal(m.DebugFetch("crossed_index")[0], a["crossed_index"])
    assert rel(m.DebugFetch("position")[0], a["position"]) < RTOL
    assert np.max(np.abs(f - fo)) < RTOL * np.max(np.abs(a["mean"]))
    # common random numbers: same seed -> same ensemble on every call; new seed -> a new one
    assert np.array_equal(m.ComputeF(Z_DRIVER), f)
    m.PostProcess()
    assert not np.array_equal(m.ComputeF(Z_DRIVER), f)


def test_jacobian_matches_oracle_and_column_loop(b200, oracle):
    R, N, eps = 4, 1024, 1e-2
    m = make_map(b200, dict(R=R, N=N))
    J, f0 = m.ComputeDFDU(Z_DRIVER, eps, return_f0=True)
    Jo, f0o = oracle.edm_compute_dfdu(oracle.edm_cfg(R=R, N=N), Z_DRIVER, eps, nthreads=4)
    assert np.max(np.abs(J - Jo)) < 1e-8 * np.max(np.abs(Jo))      # (df - f)/eps amplifies 1e-10 by 1/eps
    assert np.max(np.abs(f0 - f0o)) < RTOL * 2.0
    # the batched evaluation is bitwise the reference's sequential column loop (NewtonSolver.cpp:181-195)
    du = Z_DRIVER.copy()
    for i in range(3):
        if i > 0:
            du[i - 1] = Z_DRIVER[i - 1]
        du[i] += eps
        assert np.array_equal(J[:, i], (m.ComputeF(du) - f0) * eps ** -1)
    gj = GOLD["jacobian_driver_N1024"]
    assert np.max(np.abs(J - np.array(gj["J"]))) < 1e-8 * np.max(np.abs(gj["J"]))
    # the Jacobian from a residual already in hand (what NewtonSolver.cpp:110 computed before :93 asks): same bits,
    # n evaluations instead of n + 1
    assert np.array_equal(m.ComputeDFDU(Z_DRIVER, eps, f0=m.ComputeF(Z_DRIVER)), J)
    with pytest.raises(ValueError):
        m.ComputeDFDU(Z_DRIVER, eps, f0=np.zeros(2))


def test_batch_and_sharded_items_are_bitwise_identical(b200):
    """compute_f_batch == per-column compute_f, and evolving item slices separately then
    reducing the gathered positions gives the same bits as the single launch (the multi-GPU
    Jacobian relies on this)."""
    import torch
    R, N = 10, 512
    m = make_map(b200, dict(R=R, N=N, sigma=0.4, seed=3))
    zc = np.stack([Z_DRIVER, Z_DRIVER + [1e-2, 0, 0], Z_DRIVER + [0, 1e-2, 0]], axis=1)
    fb = m.ComputeFBatch(zc)
    for c in range(3):
        assert np.array_equal(fb[:, c], m.ComputeF(zc[:, c]))
    items = 3 * R
    pos = torch.zeros(items, 3, dtype=torch.float64, device="cuda")
    acc = torch.zeros(items, dtype=torch.int32, device="cuda")
    for lo, hi in ((0, 7), (7, 8), (8, 23), (23, 30)):       # ragged slices, not column aligned
        m.EvolveItemsDev(zc, lo, hi, pos[lo:hi], acc[lo:hi])
    fd = torch.zeros(3, 3, dtype=torch.float64, device="cuda")
    m.ReduceItemsDev(zc, pos, acc, fd)
    torch.cuda.synchronize()
    assert np.array_equal(fd.cpu().numpy().T, fb)


def test_setters_follow_the_reference(b200, oracle):
    m = make_map(b200, dict(R=2, N=1024))
    m.SetNoThreads(512)                                       # Driver.cu:69
    f = m.ComputeF(Z_DRIVER)
    fo, _ = oracle.edm_compute_f(oracle.edm_cfg(R=2, N=512), Z_DRIVER)
    assert np.max(np.abs(f - fo)) < RTOL * 2
    m.SetNoRealisations(5); m.SetTimeHorizon(3.0); m.SetParameters(0, 12.5)
    f = m.ComputeF(Z_DRIVER)
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=5, N=512, time_horizon=3.0, beta=12.5), Z_DRIVER)
    assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
    assert np.max(np.abs(f - fo)) < RTOL * 2
    for bad in (lambda: m.SetTimeHorizon(0.0), lambda: m.SetParameterStdDev(-1.0), lambda: m.SetParameters(3, 1.0),
                lambda: m.ComputeF(np.array([0.3, 0.6])), lambda: m.SetNoThreads(1)):
        with pytest.raises(b200.B200Error):
            bad()


def test_quiet_ring_fallback_and_flags(b200, oracle):
    """No neuron can fire: the arg-min falls on the smallest index at time 100 (Q5) and the
    realisation is not accepted; a front outside the domain raises the Q15 flag."""
    z = np.array([0.3310, 0.6914, 20.0])
    m = make_map(b200, dict(R=2, N=256))
    f = m.ComputeF(z)
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=2, N=256), z)
    assert m.LastInitClamped() and a["init_index_clamped"] == 1
    assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
    assert np.array_equal(m.DebugFetch("accept")[0], a["accept"])
    assert np.array_equal(np.isnan(f), np.isnan(fo))
    m2 = make_map(b200, dict(R=2, N=200))
    m2.EnableTiming(True)
    f = m2.ComputeF(Z_DRIVER)
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=2, N=200), Z_DRIVER)
    assert np.array_equal(m2.DebugFetch("event_count")[0], a["event_count"])
    assert np.array_equal(m2.DebugFetch("last_index")[0], a["last_index"])
    assert rel(m2.DebugFetch("lift_v")[0], a["lift_v"]) < RTOL


def test_default_ensemble_full_size(b200, oracle):
    """BASELINE config 3 at full size (R=1000, N=1024): every realisation reproduces the
    oracle's single-realisation event sequence; size-independent property: with sigma = 0 all
    realisations are identical, so the mean equals any one of them."""
    m = make_map(b200, dict(R=1000, N=1024))
    f = m.ComputeF(Z_DRIVER)
    fo, a = oracle.edm_compute_f(oracle.edm_cfg(R=1000, N=1024), Z_DRIVER, r_begin=0, r_end=1)
    ev = m.DebugFetch("event_count")[0]
    assert np.all(ev == a["event_count"][0])
    pos = m.DebugFetch("position")[0]
    assert np.all(pos == pos[0])
    assert rel(pos[0], a["position"][0]) < RTOL
    assert np.max(np.abs(f - fo)) < RTOL * 2


def test_in_process_multi_device_is_bitwise_single_device(b200):
    """b200_edm_set_devices: items split over the GPUs of one process, reduced on the first —
    same bits as one GPU (front map, heterogeneous ensemble, Jacobian, profile map)."""
    n_dev = b200.device_count()
    if n_dev < 2:
        pytest.skip("needs at least two GPUs in this process")
    devs = list(range(min(n_dev, 4)))
    zc = np.stack([Z_DRIVER, Z_DRIVER + [1e-2, 0, 0], Z_DRIVER + [0, 0, 1e-2]], axis=1)
    for sigma in (0.0, 0.4):
        a = make_map(b200, dict(R=37, N=512, sigma=sigma, seed=5))
        b = make_map(b200, dict(R=37, N=512, sigma=sigma, seed=5))
        b.SetDevices(devs)
        assert np.array_equal(a.ComputeFBatch(zc), b.ComputeFBatch(zc))
        assert np.array_equal(a.ComputeDFDU(Z_DRIVER, 1e-2), b.ComputeDFDU(Z_DRIVER, 1e-2))
        b.SetNoThreads(1024); a.SetNoThreads(1024)          # helpers follow reconfiguration
        assert np.array_equal(a.ComputeF(Z_DRIVER), b.ComputeF(Z_DRIVER))
    a = make_map(b200, dict(R=5, N=512)); b = make_map(b200, dict(R=5, N=512)); b.SetDevices(devs)
    for m in (a, b):
        m.SetTimeHorizon(0.5); m.SetProfileMode(64)
    u = np.concatenate([np.linspace(0.2, 0.95, 64), np.linspace(0.0, 0.5, 64)])
    assert np.array_equal(a.ComputeF(u), b.ComputeF(u))
    # 128 columns over <= 4 devices: whole columns per device, local reduction, one all-gather of the columns
    Ja, fa = a.ComputeDFDU(u, 1e-3, return_f0=True); Jb, fb = b.ComputeDFDU(u, 1e-3, return_f0=True)
    assert np.array_equal(Ja, Jb) and np.array_equal(fa, fb)
    uc = np.stack([u * (1 + 1e-3 * k) for k in range(40)], axis=1)
    assert np.array_equal(a.ComputeFBatch(uc), b.ComputeFBatch(uc))
    # the Jacobian from a residual already in hand (b200_edm_compute_dfdu_given_f): item mode and column mode
    c = make_map(b200, dict(R=37, N=512, sigma=0.4, seed=5)); c.SetDevices(devs)
    Jc, fc = c.ComputeDFDU(Z_DRIVER, 1e-2, return_f0=True)
    assert np.array_equal(c.ComputeDFDU(Z_DRIVER, 1e-2, f0=fc), Jc)
    assert np.array_equal(b.ComputeDFDU(u, 1e-3, f0=fb), Jb)
    # the library hands the caller's current device back, whichever device it was (the map lives on device 0)
    import torch
    torch.cuda.set_device(n_dev - 1)
    try:
        fb = b.ComputeF(u)
        assert torch.cuda.current_device() == n_dev - 1
        assert np.array_equal(fb, a.ComputeF(u))
        assert torch.cuda.current_device() == n_dev - 1
    finally:
        torch.cuda.set_device(0)


def test_drive_above_threshold_bypasses_the_candidate_filter(b200, oracle):
    """The two-stage candidate filter assumes vth - I > 0 (round-1 advisor finding): with the drive at or above
    threshold every neuron is handed to the exact predicate instead, and the event sequence is still the oracle's."""
    for I in (1.0, 1.2):
        cfg = dict(R=2, N=256, I=I, time_horizon=0.3)
        m = make_map(b200, cfg)
        f = m.ComputeF(Z_DRIVER)
        fo, a = oracle.edm_compute_f(oracle.edm_cfg(**cfg), Z_DRIVER)
        assert np.array_equal(m.DebugFetch("event_count")[0], a["event_count"])
        assert np.array_equal(m.DebugFetch("last_index")[0], a["last_index"])
        assert np.array_equal(np.isfinite(f), np.isfinite(fo))
        ok = np.isfinite(fo)
        assert np.max(np.abs(f[ok] - fo[ok]), initial=0.0) < 1e-9


def test_repeated_evaluations_are_bitwise_identical(b200):
    """compute-sanitizer is closed on this GPU pool (profiles/r2_sanitizer.md), so the race check is empirical:
    the evolve kernel's double-buffered candidate list, its one-event-late bookkeeping and the speculative event
    must give the same bits on every run — 40 evaluations of a heterogeneous ensemble at three launch shapes, front
    and profile map, every per-item output compared."""
    for N, prof in ((512, 0), (1024, 0), (2048, 0), (512, 48)):
        m = make_map(b200, dict(R=24, N=N, sigma=0.4, seed=11))
        if prof:
            m.SetTimeHorizon(0.7); m.SetProfileMode(prof)
            u = np.concatenate([0.2 + 0.7 * np.sin(np.linspace(0, np.pi, prof)) ** 2, np.linspace(0.0, 0.4, prof)])
        else:
            u = Z_DRIVER
        ref = None
        for rep in range(40):
            f = m.ComputeF(u)
            got = [f, m.DebugFetch("position"), m.DebugFetch("event_count"), m.DebugFetch("accept")]
            if not prof:
                got += [m.DebugFetch("last_index"), m.DebugFetch("crossed_time")]
            if ref is None:
                ref = got
            else:
                assert all(np.array_equal(a, b, equal_nan=True) for a, b in zip(ref, got)), (N, prof, rep)
