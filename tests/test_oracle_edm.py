"""CPU tests of the map oracle: committed golden fixtures, the survey's independent NumPy
emulation values (the only pin there is — BASELINE.md §5), FD-Newton fixed point, Q-table."""
import json
import os

import numpy as np
import pytest

with open(os.path.join(os.path.dirname(__file__), "golden", "edm_golden.json")) as fh:
    GOLD = json.load(fh)
CASES = {c["name"]: c for c in GOLD["cases"]}
Z_DRIVER = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)


@pytest.mark.parametrize("name", sorted(CASES))
def test_golden(oracle, name):
    c = CASES[name]
    cfg = oracle.edm_cfg(**c["cfg"])
    f, a = oracle.edm_compute_f(cfg, np.array(c["z"]), nthreads=2)
    tol = 1e-12 if c["cfg"].get("precision", 0) == 0 else 1e-6
    assert np.allclose(f, c["f"], rtol=0, atol=tol, equal_nan=True)
    for k in ("init_index", "event_count", "last_index", "crossed_index", "accept"):
        assert np.array_equal(a[k], np.array(c[k])), k
    assert np.allclose(a["position"], c["position"], rtol=0, atol=tol, equal_nan=True)
    assert np.allclose(a["lift_v"][::64], c["lift_v_head"], rtol=1e-12, atol=tol, equal_nan=True)
    assert np.allclose(a["coupling"][::64], c["coupling_head"], rtol=1e-12, atol=1e-15)


def test_survey_emulation_pin(oracle):
    """The oracle reproduces the survey's separately written NumPy emulation."""
    kat = GOLD["survey_kat"]["N1024"]
    cfg = oracle.edm_cfg(R=1, N=1024, I=0.9, beta=13.0589)
    f, a = oracle.edm_compute_f(cfg, np.array([0.3310, 0.6914, 1.3557]))
    assert a["init_index"].tolist() == kat["init_index"]
    assert a["event_count"][0] == kat["events"]
    assert a["last_index"][0].tolist() == kat["last_index"]
    assert a["crossed_index"][0].tolist() == kat["crossed_index"]
    assert np.allclose(a["position"][0], kat["X_T"], rtol=0, atol=5e-9)
    assert np.allclose(f, kat["F"], rtol=0, atol=5e-9)
    assert abs(np.linalg.norm(f) - kat["normF"]) < 5e-9
    cfg = oracle.edm_cfg(R=1, N=512, I=0.9, beta=13.0589)
    f, a = oracle.edm_compute_f(cfg, np.array([0.3310, 0.6914, 1.3557]))
    assert a["event_count"][0] == GOLD["survey_kat"]["N512"]["events"]
    assert abs(np.linalg.norm(f) - GOLD["survey_kat"]["N512"]["normF"]) < 5e-7


def test_newton_fixed_point_and_spectrum(oracle):
    """FD-Newton with the driver's settings (Driver.cu:28-37: tol 1e-4, eps 1e-2, <= 10 its)
    lands where the survey's emulation did: Z* ~ (0.331444, 0.695637, 1.365721), one unstable
    eigenvalue of I + J ~ (7.47, 0.765, 0.697)."""
    cfg = oracle.edm_cfg(R=1, N=1024, I=0.9, beta=13.0589)
    z = np.array([0.3310, 0.6914, 1.3557])
    f, _ = oracle.edm_compute_f(cfg, z, aux=False)
    its = 0
    while np.linalg.norm(f) > 1e-4 and its < 10:
        J, f0 = oracle.edm_compute_dfdu(cfg, z, 1e-2)
        z = z + np.linalg.solve(J, -f0)
        f, _ = oracle.edm_compute_f(cfg, z, aux=False)
        its += 1
    assert np.linalg.norm(f) <= 1e-4 and its == 8
    assert abs(np.linalg.norm(f) - 1.96e-5) < 1e-7
    assert np.allclose(z, [0.331444, 0.695637, 1.365721], atol=2e-6)
    # the survey quotes the spectrum of the Jacobian of the LAST Newton step (NewtonSolver.cpp
    # hands that matrix out through pJacobianExternal, :148-154)
    lam = np.sort(np.abs(np.linalg.eigvals(J + np.eye(3))))[::-1]
    assert np.allclose(lam, [7.4718, 0.7650, 0.6965], atol=2e-3)
    assert int(np.sum(lam > 1.0)) == 1


def test_realisations_identical_when_homogeneous(oracle):
    """sigma = 0 (reference default, EventDrivenMap.cu:105): every realisation is the same ring."""
    cfg = oracle.edm_cfg(R=3, N=256)
    _, a = oracle.edm_compute_f(cfg, Z_DRIVER)
    assert np.array_equal(a["position"][0], a["position"][1]) and np.array_equal(a["position"][1], a["position"][2])


def test_quirk_accept0_bias(oracle):
    """Q1: with the reference's CountRealisationsKernel quirk the mean is (R-1)/R of the intended one."""
    f0, a0 = oracle.edm_compute_f(oracle.edm_cfg(R=4, N=256), Z_DRIVER)
    f1, a1 = oracle.edm_compute_f(oracle.edm_cfg(R=4, N=256, quirks=1), Z_DRIVER)
    assert np.allclose(a1["mean"], a0["mean"] * 3 / 4, rtol=1e-14)


def test_front_outside_domain_is_flagged(oracle):
    """Q15: c*T_m >= L leaves the reference's index unassigned; the oracle clamps to 0 and flags it."""
    cfg = oracle.edm_cfg(R=1, N=256)
    _, a = oracle.edm_compute_f(cfg, np.array([0.3310, 0.6914, 20.0]))
    assert a["init_index_clamped"] == 1 and a["init_index"][2] == 0


def test_normal_generator(oracle):
    assert np.allclose([oracle.normal(42, i) for i in range(8)], GOLD["normal_seed42"], rtol=0, atol=1e-15)
    x = np.array([oracle.normal(7, i) for i in range(20000)])
    assert abs(x.mean()) < 0.03 and abs(x.std() - 1.0) < 0.03


def test_bounded_sample_matches_full(oracle):
    """r_begin/r_end (used by the CPU-baseline leg) evolve exactly the requested realisations."""
    cfg = oracle.edm_cfg(R=6, N=256, sigma=0.3, seed=5)
    _, full = oracle.edm_compute_f(cfg, Z_DRIVER)
    _, part = oracle.edm_compute_f(cfg, Z_DRIVER, r_begin=2, r_end=5)
    assert np.array_equal(part["position"], full["position"][2:5])
