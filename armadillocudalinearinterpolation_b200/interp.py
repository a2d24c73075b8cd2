"""interp1 / interp2 — the host-side mirror of arma::interp1 / arma::interp2 for this path.

    arma::interp1(X, Y, XI, YI, "*linear", extrap)        -> YI = interp1(X, Y, XI, extrap)
    arma::interp2(X, Y, Z, XI, YI, ZI, "linear", extrap)  -> ZI = interp2(X, Y, Z, XI, YI, extrap)

Same argument order and meaning as Armadillo's free functions (fn_interp1.hpp / fn_interp2.hpp;
the reference links Armadillo, Makefile:5, but vendors none of it).  numpy arrays are host
buffers and go through the host entry points of the C-ABI (include/b200_interp.h); torch CUDA
tensors stay resident in HBM and go through the *_dev entry points on torch's current stream.
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import B200_F32, B200_F64, check

_DT = {np.dtype(np.float64): B200_F64, np.dtype(np.float32): B200_F32}


def _np(a, dt=None):
    a = np.ascontiguousarray(a) if dt is None else np.ascontiguousarray(a, dtype=dt)
    if a.dtype not in _DT:
        raise TypeError(f"dtype {a.dtype} is not supported (float64 / float32 only)")
    return a


def _ptr(a):
    return C.c_void_p(a.ctypes.data)


def _is_torch(t):
    return type(t).__module__.startswith("torch")


def _torch_stream():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class Interp1Plan:
    """Grid (X, Y) resident in HBM; interpolate many query batches (b200_interp1_plan_*)."""

    def __init__(self, X, Y):
        X = _np(X)
        Y = _np(Y, X.dtype)
        if X.ndim != 1 or X.shape != Y.shape:
            raise ValueError("X and Y must be vectors of equal length")
        self.dtype = X.dtype
        self.n = X.size
        self._h = C.c_void_p()
        check(_lib.lib().b200_interp1_plan_create(_DT[X.dtype], _ptr(X), _ptr(Y), C.c_size_t(X.size), C.byref(self._h)))

    @property
    def lookup_mode(self):
        return _lib.lib().b200_interp1_plan_lookup_mode(self._h)

    def set_values(self, Y):
        Y = _np(Y, self.dtype)
        if Y.size != self.n:
            raise ValueError("Y has the wrong length")
        check(_lib.lib().b200_interp1_plan_set_values(self._h, _ptr(Y)))

    def __call__(self, XI, extrap=np.nan, return_index=False, out=None):
        if _is_torch(XI):
            return self._call_dev(XI, extrap, return_index, out)
        XI = _np(XI, self.dtype)
        YI = np.empty(XI.shape, self.dtype) if out is None else out
        idx = np.empty(XI.shape, np.int32) if return_index else None
        check(_lib.lib().b200_interp1_exec(self._h, _ptr(XI), C.c_size_t(XI.size), _ptr(YI),
                                          _ptr(idx) if return_index else None, C.c_double(extrap)))
        return (YI, idx) if return_index else YI

    def _call_dev(self, XI, extrap, return_index, out):
        import torch
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        if not XI.is_cuda or XI.dtype != tdt or not XI.is_contiguous():
            raise TypeError("device queries must be contiguous CUDA tensors of the plan's dtype")
        YI = torch.empty_like(XI) if out is None else out
        idx = torch.empty(XI.shape, dtype=torch.int32, device=XI.device) if return_index else None
        check(_lib.lib().b200_interp1_exec_dev(self._h, C.c_void_p(XI.data_ptr()), C.c_size_t(XI.numel()),
                                              C.c_void_p(YI.data_ptr()),
                                              C.c_void_p(idx.data_ptr()) if return_index else None,
                                              C.c_double(extrap), _torch_stream()))
        return (YI, idx) if return_index else YI

    def close(self):
        if self._h:
            _lib.lib().b200_interp1_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Interp2Plan:
    """Grid (X, Y, Z) resident in HBM.  Z is Y.size x X.size (rows follow Y), any memory order;
    it is stored column-major like arma::mat."""

    NO_CELLS, FORCE_CELLS, NO_BANDS, FORCE_BANDS, NO_TILES, FORCE_TILES, ORDER_YX = 1, 2, 4, 8, 16, 32, 64   # include/b200_interp.h flags

    def __init__(self, X, Y, Z, flags=0):
        X = _np(X)
        Y = _np(Y, X.dtype)
        Z = np.asarray(Z)
        if Z.shape != (Y.size, X.size):
            raise ValueError("Z must be Y.size x X.size (arma: X.n_elem == Z.n_cols, Y.n_elem == Z.n_rows)")
        Zf = np.asfortranarray(Z, dtype=X.dtype)
        self.dtype = X.dtype
        self.nx, self.ny = X.size, Y.size
        self._h = C.c_void_p()
        check(_lib.lib().b200_interp2_plan_create_ex(_DT[X.dtype], _ptr(X), C.c_size_t(X.size), _ptr(Y),
                                                    C.c_size_t(Y.size), _ptr(Zf), C.c_uint(flags), C.byref(self._h)))

    def grid(self, XI, YI, extrap=np.nan):
        """Tensor-grid queries (Armadillo's interp2 shape): returns ZI of shape (YI.size, XI.size)."""
        if _is_torch(XI):
            import torch
            ZI = torch.empty((XI.numel(), YI.numel()), dtype=XI.dtype, device=XI.device)  # column-major ZI
            check(_lib.lib().b200_interp2_grid_dev(self._h, C.c_void_p(XI.data_ptr()), C.c_size_t(XI.numel()),
                                                  C.c_void_p(YI.data_ptr()), C.c_size_t(YI.numel()),
                                                  C.c_void_p(ZI.data_ptr()), C.c_double(extrap), _torch_stream()))
            return ZI.t()
        XI = _np(XI, self.dtype)
        YI = _np(YI, self.dtype)
        ZI = np.empty((YI.size, XI.size), self.dtype, order="F")
        check(_lib.lib().b200_interp2_grid(self._h, _ptr(XI), C.c_size_t(XI.size), _ptr(YI), C.c_size_t(YI.size),
                                          _ptr(ZI), C.c_double(extrap)))
        return ZI

    def scattered(self, XQ, YQ, extrap=np.nan, out=None):
        """Scattered queries (XQ[k], YQ[k]) -> ZQ[k]."""
        if _is_torch(XQ):
            import torch
            ZQ = torch.empty_like(XQ) if out is None else out
            check(_lib.lib().b200_interp2_scattered_dev(self._h, C.c_void_p(XQ.data_ptr()), C.c_void_p(YQ.data_ptr()),
                                                       C.c_size_t(XQ.numel()), C.c_void_p(ZQ.data_ptr()),
                                                       C.c_double(extrap), _torch_stream()))
            return ZQ
        XQ = _np(XQ, self.dtype)
        YQ = _np(YQ, self.dtype)
        if XQ.shape != YQ.shape:
            raise ValueError("XQ and YQ must have the same shape")
        ZQ = np.empty(XQ.shape, self.dtype) if out is None else out
        check(_lib.lib().b200_interp2_scattered(self._h, _ptr(XQ), _ptr(YQ), C.c_size_t(XQ.size), _ptr(ZQ),
                                               C.c_double(extrap)))
        return ZQ

    def close(self):
        if self._h:
            _lib.lib().b200_interp2_plan_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def interp1(X, Y, XI, extrap=np.nan, return_index=False):
    """arma::interp1(X, Y, XI, YI, "*linear", extrap) — one-shot call (b200_interp1_f64/_f32)."""
    X = _np(X)
    Y = _np(Y, X.dtype)
    XI = _np(XI, X.dtype)
    YI = np.empty(XI.shape, X.dtype)
    idx = np.empty(XI.shape, np.int32) if return_index else None
    fn = _lib.lib().b200_interp1_f64 if X.dtype == np.float64 else _lib.lib().b200_interp1_f32
    ex = C.c_double(extrap) if X.dtype == np.float64 else C.c_float(extrap)
    check(fn(_ptr(X), _ptr(Y), C.c_size_t(X.size), _ptr(XI), C.c_size_t(XI.size), _ptr(YI),
             _ptr(idx) if return_index else None, ex))
    return (YI, idx) if return_index else YI


def interp2(X, Y, Z, XI, YI, extrap=np.nan):
    """arma::interp2(X, Y, Z, XI, YI, ZI, "linear", extrap) — one-shot call (b200_interp2_f64/_f32)."""
    X = _np(X)
    Y = _np(Y, X.dtype)
    Z = np.asarray(Z)
    if Z.shape != (Y.size, X.size):
        raise ValueError("Z must be Y.size x X.size")
    Zf = np.asfortranarray(Z, dtype=X.dtype)
    XI = _np(XI, X.dtype)
    YI = _np(YI, X.dtype)
    ZI = np.empty((YI.size, XI.size), X.dtype, order="F")
    fn = _lib.lib().b200_interp2_f64 if X.dtype == np.float64 else _lib.lib().b200_interp2_f32
    ex = C.c_double(extrap) if X.dtype == np.float64 else C.c_float(extrap)
    check(fn(_ptr(X), C.c_size_t(X.size), _ptr(Y), C.c_size_t(Y.size), _ptr(Zf), _ptr(XI), C.c_size_t(XI.size),
             _ptr(YI), C.c_size_t(YI.size), _ptr(ZI), ex))
    return ZI
