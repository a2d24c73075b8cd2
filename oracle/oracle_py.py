"""ctypes loader for the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product package never does.  PARITY UNPINNED — see oracle/oracle.h.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE)
            if f.endswith((".c", ".inc", ".h")) or f == "Makefile"]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


class EdmCfg(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("vth", "a1", "a2", "b1", "b2", "I", "L", "tol", "time_horizon")] + [
        ("counter_max", C.c_uint32), ("quirks", C.c_uint32),
        ("beta", C.c_double), ("sigma", C.c_double), ("seed", C.c_uint64),
        ("N", C.c_uint32), ("R", C.c_uint32), ("M", C.c_uint32), ("precision", C.c_uint32),
        ("beta_ext", C.c_void_p)]


class EdmAux(C.Structure):
    _fields_ = [("init_index", C.c_void_p), ("lift_v", C.c_void_p), ("lift_s", C.c_void_p),
                ("last_index", C.c_void_p), ("last_time", C.c_void_p),
                ("crossed_index", C.c_void_p), ("crossed_time", C.c_void_p),
                ("accept", C.c_void_p), ("position", C.c_void_p), ("event_count", C.c_void_p),
                ("mean", C.c_void_p), ("beta", C.c_void_p), ("coupling", C.c_void_p),
                ("init_index_clamped", C.c_int32),
                ("total_events", C.c_uint64), ("total_neuron_events", C.c_uint64),
                ("total_candidates", C.c_uint64), ("total_newton_its", C.c_uint64)]


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        _LIB.oracle_normal.restype = C.c_double
        _LIB.oracle_normal.argtypes = [C.c_uint64, C.c_uint64]
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _real(dtype):
    dtype = np.dtype(dtype)
    assert dtype in (np.float64, np.float32)
    return dtype, ("f64" if dtype == np.float64 else "f32"), (C.c_double if dtype == np.float64 else C.c_float)


def interp1(xg, yg, xi, extrap=np.nan, scan=False, nthreads=1, want_idx=True):
    dt, sfx, creal = _real(xg.dtype)
    xg = np.ascontiguousarray(xg, dt); yg = np.ascontiguousarray(yg, dt); xi = np.ascontiguousarray(xi, dt)
    yi = np.empty(xi.shape, dt)
    idx = np.empty(xi.shape, np.int32) if want_idx else None
    f = getattr(lib(), "oracle_interp1_" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p,
                  C.c_void_p, creal, C.c_int, C.c_int]
    rc = f(_p(xg), _p(yg), xg.size, _p(xi), xi.size, _p(yi), _p(idx), extrap, int(scan), nthreads)
    if rc:
        raise ValueError(f"oracle_interp1 rc={rc}")
    return (yi, idx) if want_idx else yi


def _order(y_first):
    lib().oracle_interp2_set_order(int(bool(y_first)))


def interp2_grid(x, y, z, xi, yi, extrap=np.nan, nthreads=1, y_first=False):
    """z: (ny, nx) array (any memory order) -> zi (nyi, nxi).  y_first: the mirrored pass order."""
    _order(y_first)
    dt, sfx, creal = _real(z.dtype)
    x = np.ascontiguousarray(x, dt); y = np.ascontiguousarray(y, dt)
    xi = np.ascontiguousarray(xi, dt); yi = np.ascontiguousarray(yi, dt)
    zf = np.asfortranarray(z, dt)
    zi = np.empty((yi.size, xi.size), dt, order="F")
    f = getattr(lib(), "oracle_interp2_grid_" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_size_t,
                  C.c_void_p, C.c_size_t, C.c_void_p, creal, C.c_int]
    rc = f(_p(x), x.size, _p(y), y.size, _p(zf), _p(xi), xi.size, _p(yi), yi.size, _p(zi), extrap, nthreads)
    if rc:
        raise ValueError(f"oracle_interp2_grid rc={rc}")
    return zi


def interp2_scattered(x, y, z, xq, yq, extrap=np.nan, nthreads=1, y_first=False):
    _order(y_first)
    dt, sfx, creal = _real(z.dtype)
    x = np.ascontiguousarray(x, dt); y = np.ascontiguousarray(y, dt)
    xq = np.ascontiguousarray(xq, dt); yq = np.ascontiguousarray(yq, dt)
    zf = np.asfortranarray(z, dt)
    zq = np.empty(xq.shape, dt)
    f = getattr(lib(), "oracle_interp2_scattered_" + sfx)
    f.restype = C.c_int
    f.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_void_p, C.c_void_p,
                  C.c_size_t, C.c_void_p, creal, C.c_int]
    rc = f(_p(x), x.size, _p(y), y.size, _p(zf), _p(xq), _p(yq), xq.size, _p(zq), extrap, nthreads)
    if rc:
        raise ValueError(f"oracle_interp2_scattered rc={rc}")
    return zq


def edm_cfg(**kw):
    c = EdmCfg()
    lib().oracle_edm_cfg_default.argtypes = [C.c_void_p]
    lib().oracle_edm_cfg_default(C.byref(c))
    keep = {}
    for k, v in kw.items():
        if k == "beta_ext":
            if v is not None:
                v = np.ascontiguousarray(v, np.float64)
                keep["beta_ext"] = v
                c.beta_ext = v.ctypes.data
        else:
            setattr(c, k, v)
    c._keep = keep
    return c


def edm_compute_f(cfg, z, r_begin=0, r_end=0, nthreads=1, aux=True):
    """Returns (f, aux dict).  aux arrays cover realisations [r_begin, r_end) (all if 0,0)."""
    z = np.ascontiguousarray(z, np.float64)
    M, N, R = cfg.M, cfg.N, cfg.R
    assert z.size == M
    nr = (r_end - r_begin) if r_end > r_begin else R
    f = np.empty(M)
    a = EdmAux()
    out = {}
    if aux:
        out = dict(init_index=np.zeros(M, np.int32), lift_v=np.zeros(N), lift_s=np.zeros(N),
                   last_index=np.zeros((nr, M), np.int32), last_time=np.zeros((nr, M)),
                   crossed_index=np.zeros((nr, M), np.int32), crossed_time=np.zeros((nr, M)),
                   accept=np.zeros(nr, np.int32), position=np.zeros((nr, M)),
                   event_count=np.zeros(nr, np.int32), mean=np.zeros(M), coupling=np.zeros(N))
        for k, v in out.items():
            setattr(a, k, v.ctypes.data)
    fn = lib().oracle_edm_compute_f
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_int]
    rc = fn(C.addressof(cfg), _p(z), _p(f), C.addressof(a), r_begin, r_end, nthreads)
    if rc:
        raise ValueError(f"oracle_edm_compute_f rc={rc}")
    out.update(init_index_clamped=a.init_index_clamped, total_events=a.total_events,
               total_neuron_events=a.total_neuron_events, total_candidates=a.total_candidates,
               total_newton_its=a.total_newton_its)
    return f, out


def edm_compute_dfdu(cfg, u, eps, nthreads=1):
    u = np.ascontiguousarray(u, np.float64)
    n = cfg.M
    jac = np.empty((n, n), order="F")
    f0 = np.empty(n)
    fn = lib().oracle_edm_compute_dfdu
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, C.c_int]
    rc = fn(C.addressof(cfg), _p(u), eps, _p(jac), _p(f0), nthreads)
    if rc:
        raise ValueError(f"oracle_edm_compute_dfdu rc={rc}")
    return jac, f0


def edm_beta(cfg):
    out = np.empty((cfg.R, cfg.N))
    fn = lib().oracle_edm_beta
    fn.restype = None
    fn.argtypes = [C.c_void_p, C.c_void_p]
    fn(C.addressof(cfg), _p(out))
    return out


def normal(seed, index):
    return lib().oracle_normal(seed, index)


def profile_compute_f(cfg, n_coarse, u, r_begin=0, r_end=0, nthreads=1):
    """Profile map (config 5): returns (f, dict(restricted, accept, event_count))."""
    u = np.ascontiguousarray(u, np.float64)
    n = 2 * n_coarse
    assert u.size == n
    nr = (r_end - r_begin) if r_end > r_begin else cfg.R
    f = np.empty(n); restricted = np.empty((nr, n)); accept = np.empty(nr, np.int32); evc = np.empty(nr, np.int32)
    fn = lib().oracle_profile_compute_f
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p,
                   C.c_uint32, C.c_uint32, C.c_int]
    rc = fn(C.addressof(cfg), n_coarse, _p(u), _p(f), _p(restricted), _p(accept), _p(evc), r_begin, r_end, nthreads)
    if rc:
        raise ValueError(f"oracle_profile_compute_f rc={rc}")
    return f, dict(restricted=restricted, accept=accept, event_count=evc)


def edm_lift(cfg, z):
    v = np.empty(cfg.N); s = np.empty(cfg.N)
    fn = lib().oracle_edm_lift
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    rc = fn(C.addressof(cfg), _p(np.ascontiguousarray(z, np.float64)), _p(v), _p(s))
    if rc:
        raise ValueError(f"oracle_edm_lift rc={rc}")
    return v, s
