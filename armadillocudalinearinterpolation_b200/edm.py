"""EventDrivenMap — Python binding with the public surface of the reference class
(EventDrivenMap.hpp:18-51: ctor(pParameters, noReal), ComputeF, SetTimeHorizon,
SetNoRealisations, SetNoThreads, SetParameterStdDev, SetParameters, ResetSeed, SetNewSeed,
PostProcess, SetDebugFlag) over the C-ABI of include/b200_edm.h.  The C++ drop-in
(host/EventDrivenMapB200.hpp) is the product boundary; this binding exists for tests, the
bench and the one-process-per-GPU Jacobian sharding (parallel.py).
"""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import B200_F32, B200_F64, check


class Model(C.Structure):
    """b200_edm_model: parameters.hpp:1-15 as a runtime struct."""
    _fields_ = [(n, C.c_double) for n in ("vth", "a1", "a2", "b1", "b2", "I", "L", "tol", "time_horizon")] + [
        ("counter_max", C.c_uint32), ("quirks", C.c_uint32)]


QUIRK_ACCEPT0_BIAS = 1

DBG = dict(init_index=(0, np.int32), lift_v=(1, np.float64), lift_s=(2, np.float64),
           last_index=(3, np.int32), last_time=(4, np.float64), crossed_index=(5, np.int32),
           crossed_time=(6, np.float64), accept=(7, np.int32), position=(8, np.float64),
           event_count=(9, np.int32), mean=(10, np.float64), beta=(11, np.float64),
           coupling=(12, np.float64))


def _dp(a):
    return a.ctypes.data_as(C.c_void_p)


class EventDrivenMap:
    def __init__(self, parameters, noReal, noNeurons=1024, noFronts=3, precision="f64", verbose=False):
        p = np.ascontiguousarray(parameters, np.float64).ravel()
        self._L = _lib.lib()
        self._h = C.c_void_p()
        self.verbose = verbose
        self.precision = precision
        check(self._L.b200_edm_create(_dp(p), C.c_size_t(p.size), C.c_uint32(noReal), C.c_uint32(noNeurons),
                                      C.c_uint32(noFronts), B200_F64 if precision == "f64" else B200_F32,
                                      C.byref(self._h)))
        self.R, self.N, self.M = int(noReal), int(noNeurons), int(noFronts)
        self._fronts = int(noFronts)
        self._last_cols = 0

    # ---- AbstractNonlinearProblem (AbstractNonlinearProblem.hpp:11-13) ----
    def ComputeF(self, u):
        u = np.ascontiguousarray(u, np.float64).ravel()
        f = np.empty(u.size)
        check(self._L.b200_edm_compute_f(self._h, _dp(u), C.c_size_t(u.size), _dp(f)))
        self._last_cols = 1
        return f

    def PostProcess(self):
        self.SetNewSeed()

    # ---- AbstractNonlinearProblemJacobian (AbstractNonlinearProblemJacobian.hpp:11) ----
    def ComputeDFDU(self, u, eps, return_f0=False, f0=None):
        """Forward-difference Jacobian, all evaluations in one batch.  f0 = F(u), if the caller already holds it
        (the residual of a Newton iteration): the base evaluation is then not repeated; same bits."""
        u = np.ascontiguousarray(u, np.float64).ravel()
        n = u.size
        jac = np.empty((n, n), order="F")
        if f0 is not None:
            f0 = np.ascontiguousarray(f0, np.float64).ravel()
            if f0.size != n:
                raise ValueError("f0 and u differ in length")
            check(self._L.b200_edm_compute_dfdu_given_f(self._h, _dp(u), C.c_size_t(n), C.c_double(eps), _dp(f0), _dp(jac)))
            self._last_cols = n
            return (jac, f0.copy()) if return_f0 else jac
        f0 = np.empty(n)
        check(self._L.b200_edm_compute_dfdu(self._h, _dp(u), C.c_size_t(n), C.c_double(eps), _dp(jac), _dp(f0)))
        self._last_cols = n + 1
        return (jac, f0) if return_f0 else jac

    def ComputeFBatch(self, z_cols):
        """z_cols: (n, ncols) array, one evaluation point per column -> F of the same shape."""
        z = np.asfortranarray(z_cols, np.float64)
        if z.ndim == 1:
            z = z.reshape(-1, 1, order="F")
        n, ncols = z.shape
        f = np.empty((n, ncols), order="F")
        check(self._L.b200_edm_compute_f_batch(self._h, _dp(z), C.c_size_t(n), C.c_size_t(ncols), _dp(f)))
        self._last_cols = ncols
        return f

    # ---- setters (EventDrivenMap.cu:242-359) ----
    def SetTimeHorizon(self, T):
        check(self._L.b200_edm_set_time_horizon(self._h, C.c_double(T)))
        if self.verbose:
            print(f"Time horizon set to {T}")

    def SetNoRealisations(self, noReal):
        check(self._L.b200_edm_set_no_realisations(self._h, C.c_uint32(noReal)))
        self.R = int(noReal)
        if self.verbose:
            print(f"Number of realisations set to {noReal}")

    def SetNoThreads(self, noThreads):
        """The reference's thread count IS its neuron count (one thread per neuron)."""
        check(self._L.b200_edm_set_no_neurons(self._h, C.c_uint32(noThreads)))
        self.N = int(noThreads)
        if self.verbose:
            print(f"Number of threads set to {noThreads}")

    def SetParameterStdDev(self, sigma):
        check(self._L.b200_edm_set_param_stddev(self._h, C.c_double(sigma)))
        if self.verbose:
            print(f"Parameter standard deviation set to {sigma}")

    def SetParameters(self, parId, parVal):
        check(self._L.b200_edm_set_parameter(self._h, C.c_uint32(parId), C.c_double(parVal)))
        if self.verbose:
            print(f"Parameter value set to {parVal}")

    def ResetSeed(self):
        """Common random numbers are the default here: every ComputeF sees the same ensemble."""

    def SetSeed(self, seed):
        check(self._L.b200_edm_set_seed(self._h, C.c_uint64(seed)))

    def GetSeed(self):
        s = C.c_uint64()
        check(self._L.b200_edm_get_seed(self._h, C.byref(s)))
        return s.value

    def SetNewSeed(self):
        check(self._L.b200_edm_new_seed(self._h))
        if self.verbose:
            print("New seed set")

    def SetDebugFlag(self, val):
        check(self._L.b200_edm_set_debug(self._h, int(bool(val))))
        if self.verbose:
            print("Debugging on" if val else "Debugging off")

    # ---- model / tuning / introspection ----
    def GetModel(self):
        m = Model()
        check(self._L.b200_edm_get_model(self._h, C.byref(m)))
        return m

    def SetModel(self, **kw):
        m = self.GetModel()
        for k, v in kw.items():
            setattr(m, k, v)
        check(self._L.b200_edm_set_model(self._h, C.byref(m)))

    def SetProfileMode(self, n_coarse):
        """Switch to the profile map on n_coarse coarse knots (vectors of length 2 n_coarse); 0 = front map."""
        check(self._L.b200_edm_set_profile_mode(self._h, C.c_uint32(n_coarse)))
        self.M = 2 * int(n_coarse) if n_coarse else self._fronts
        self._last_cols = 0

    def SetDevices(self, device_ids):
        """Split every evaluation over these devices of this process (first = the handle's own device)."""
        ids = (C.c_int * len(device_ids))(*device_ids)
        check(self._L.b200_edm_set_devices(self._h, ids, C.c_size_t(len(device_ids))))

    def SetTuning(self, neurons_per_thread):
        check(self._L.b200_edm_set_tuning(self._h, int(neurons_per_thread)))

    def EnableTiming(self, on=True):
        check(self._L.b200_edm_enable_timing(self._h, int(bool(on))))

    def LastEvolveMs(self):
        ms = C.c_double()
        check(self._L.b200_edm_last_evolve_ms(self._h, C.byref(ms)))
        return ms.value

    def LastCounters(self):
        out = (C.c_uint64 * 4)()
        check(self._L.b200_edm_last_counters(self._h, out))
        return dict(events=out[0], candidates=out[1], newton_its=out[2], fallbacks=out[3])

    def LastInitClamped(self):
        c = C.c_int()
        check(self._L.b200_edm_last_init_clamped(self._h, C.byref(c)))
        return bool(c.value)

    def DebugFetch(self, what):
        """Arrays of the most recent evaluation (replaces the Save*() dumps, EventDrivenMap.cu:406-503)."""
        code, dt = DBG[what]
        C_, R, N, M = self._last_cols, self.R, self.N, self.M
        shape = {"init_index": (C_, M), "lift_v": (C_, N), "lift_s": (C_, N), "accept": (C_, R),
                 "event_count": (C_, R), "mean": (C_, M), "beta": (R, N), "coupling": (N,)}.get(what, (C_, R, M))
        out = np.empty(shape, dt)
        check(self._L.b200_edm_debug_fetch(self._h, code, _dp(out), C.c_size_t(out.nbytes)))
        return out

    # ---- sharded evaluation on device buffers (torch tensors) ----
    def EvolveItemsDev(self, z_cols, item_begin, item_end, pos, accept, stream=None):
        z = np.asfortranarray(z_cols, np.float64)
        n, ncols = z.shape
        check(self._L.b200_edm_evolve_items_dev(self._h, _dp(z), C.c_size_t(n), C.c_size_t(ncols),
                                                C.c_size_t(item_begin), C.c_size_t(item_end),
                                                C.c_void_p(pos.data_ptr()), C.c_void_p(accept.data_ptr()),
                                                C.c_void_p(stream or 0)))
        self._last_cols = ncols

    def ReduceItemsDev(self, z_cols, pos_all, accept_all, f_cols, stream=None):
        z = np.asfortranarray(z_cols, np.float64)
        n, ncols = z.shape
        check(self._L.b200_edm_reduce_items_dev(self._h, _dp(z), C.c_size_t(n), C.c_size_t(ncols),
                                                C.c_void_p(pos_all.data_ptr()), C.c_void_p(accept_all.data_ptr()),
                                                C.c_void_p(f_cols.data_ptr()), C.c_void_p(stream or 0)))

    def close(self):
        if self._h:
            self._L.b200_edm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
