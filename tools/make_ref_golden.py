"""Runs the UNMODIFIED reference (oracle/_ref, built by oracle/ref_build/Makefile) on a B200 and
writes its raw outputs to gpurun_out/ref_golden.npz (+ a text report).  The committed copy of that
file is tests/golden/ref_b200.npz: the pin of the CPU oracle (CPU tests) and of the product's FP32
mode (GPU tests) against the real reference.  Usage on the GPU box:  python tools/make_ref_golden.py"""
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_py as REF  # noqa: E402
from oracle import oracle_py as O  # noqa: E402

BETA = float(np.float32(13.0589))                     # Driver.cu:16
GUESS = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], np.float64)  # Driver.cu:24
out = {}
rep = []


def log(*a):
    s = " ".join(str(x) for x in a)
    print(s, flush=True)
    rep.append(s)


def case(tag, z, R, N, T=5.0, sigma=0.0, seed=42):
    t = time.time()
    f, a = REF.run(z, BETA, R, N, T, sigma, seed)
    dt = time.time() - t
    same = bool(np.all(a["lift_v"] == a["lift_v"][0]) or np.all(np.isnan(a["lift_v"]) == np.isnan(a["lift_v"][0])))
    out[f"{tag}_z"] = np.asarray(z, np.float64)
    out[f"{tag}_cfg"] = np.array([BETA, R, N, T, sigma, seed], np.float64)
    out[f"{tag}_F"] = f
    for k in ("coupling", "init_index", "last_index", "last_time", "crossed_index", "crossed_time", "accept",
              "position", "mean", "position_cf", "mean_replay"):
        out[f"{tag}_{k}"] = a[k]
    out[f"{tag}_lift_v"] = a["lift_v"][0]
    out[f"{tag}_lift_s"] = a["lift_s"][0]
    if sigma > 0:
        out[f"{tag}_beta"] = a["beta"]
    if not a["replay_equal"]:
        log(f"[{tag}] REPLAY DIFFERS from ComputeF: mean {a['mean']} vs replay {a['mean_replay']}; positions differ at "
            f"{int((a['position'] != a['position_cf']).sum())} of {a['position'].size} entries")
    log(f"[{tag}] R={R} N={N} sigma={sigma}: {dt*1e3:.1f} ms  F={f}  init={a['init_index']}  "
        f"last={a['last_index'][:,0]} crossed={a['crossed_index'][:,0]} accept_sum={int(a['accept'].sum())} "
        f"mean={a['mean']} lift_rows_identical={same} nan_in_lift={int(np.isnan(a['lift_v'][0]).sum())}")
    # quick look against the FP32 oracle (informational; the tests hold the assertions)
    cfg = O.edm_cfg(R=R, N=N, beta=BETA, sigma=sigma, seed=seed, precision=1, quirks=1, time_horizon=T,
                    beta_ext=(a["beta"].astype(np.float64) if sigma > 0 else None))
    fo, ao = O.edm_compute_f(cfg, z)
    if not a["replay_equal"]:       # heterogeneous case: the ensemble fed to the oracle is the replay's
        log(f"    oracle mean vs replay mean: {np.abs(ao['mean'] - a['mean_replay']).max():.3e}")
    log(f"    oracle f32+Q1: F={fo}  dF={np.abs(fo-f).max():.3e}  init_eq={np.array_equal(ao['init_index'], a['init_index'])} "
        f"last_eq={np.array_equal(ao['last_index'].T, a['last_index'])} crossed_eq={np.array_equal(ao['crossed_index'].T, a['crossed_index'])} "
        f"accept_eq={np.array_equal(ao['accept'], a['accept'])} "
        f"pos_maxdiff={np.nanmax(np.abs(ao['position'].T - a['position'])):.3e} "
        f"w_maxdiff={np.abs(ao['coupling'] - a['coupling']).max():.3e}")
    lv, ls = a["lift_v"][0].astype(np.float64), a["lift_s"][0].astype(np.float64)
    ok = ~np.isnan(lv) & ~np.isnan(ao["lift_v"])
    log(f"    lift: nan_eq={np.array_equal(np.isnan(lv), np.isnan(ao['lift_v']))} "
        f"v_maxdiff={np.abs(lv-ao['lift_v'])[ok].max():.3e} s_maxdiff={np.nanmax(np.abs(ls-ao['lift_s'])):.3e}")
    return f, a


case("A_default_1024", GUESS, 1000, 1024)
case("B_default_512", GUESS, 1000, 512)
case("C_offguess_1024", GUESS * np.array([1.01, 0.98, 1.03]), 8, 1024)
case("D_sigma05_1024", GUESS, 16, 1024, sigma=0.5, seed=7)
case("E_T2_768", GUESS, 5, 768, T=2.0)

f1, a1 = REF.run(GUESS, BETA, 16, 1024, 5.0, 0.5, 7)
f2, a2 = REF.run(GUESS, BETA, 16, 1024, 5.0, 0.5, 7)
log(f"[sigma repeat] two fresh maps, same seed: F equal={np.array_equal(f1, f2)} beta(replay) equal={np.array_equal(a1['beta'], a2['beta'])} "
    f"replay means equal={np.array_equal(a1['mean_replay'], a2['mean_replay'])}")
t = time.time()
flag, zs, hist, jac = REF.newton(GUESS, BETA, 1000, 1024)
log(f"[newton N=1024 R=1000 tol=1e-4 eps=1e-2] {time.time()-t:.2f} s flag={flag} z*={zs} hist={hist}")
log(f"    last jacobian=\n{jac}\n    eig(I+J)={np.linalg.eigvals(jac + np.eye(3))}")
out["N_flag"] = np.array([flag]); out["N_z"] = zs; out["N_hist"] = hist; out["N_jac"] = jac
nu = REF.unstable(zs, BETA, 1000, 1024)
log(f"[stability at z*] unstable eigenvalues = {nu}")
out["N_unstable"] = np.array([nu])
flag5, zs5, hist5, jac5 = REF.newton(GUESS, BETA, 1000, 512)
log(f"[newton N=512] flag={flag5} z={zs5} hist={hist5}")
out["N512_flag"] = np.array([flag5]); out["N512_z"] = zs5; out["N512_hist"] = hist5

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
cwd = os.path.join(ROOT, "gpurun_out", "driver_ref_cwd")
os.makedirs(cwd, exist_ok=True)
r = subprocess.run(["timeout", "120", REF.REF_DRIVER], cwd=cwd, capture_output=True, text=True)
log("[Driver_ref] rc", r.returncode)
log(r.stdout[-3000:])
log(r.stderr[-500:])
open(os.path.join(ROOT, "gpurun_out", "driver_ref_stdout.txt"), "w").write(r.stdout)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "ref_golden.npz"), **out)
open(os.path.join(ROOT, "gpurun_out", "ref_report.txt"), "w").write("\n".join(rep) + "\n")
