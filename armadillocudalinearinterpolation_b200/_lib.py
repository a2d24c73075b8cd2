"""Loader for the C-ABI library (lib/libb200edm.so, declared in include/*.h).

There is no CPU path: if the library is missing or no B200 is visible, calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libb200edm.so")
_lib = None

B200_F64, B200_F32 = 0, 1

STATUS_NAMES = {0: "B200_OK", -1: "B200_ERR_INVALID_ARG", -2: "B200_ERR_CUDA", -3: "B200_ERR_NOT_SORTED",
                -4: "B200_ERR_TOO_SMALL", -5: "B200_ERR_NO_DEVICE", -6: "B200_ERR_UNSUPPORTED",
                -7: "B200_ERR_NONFINITE"}


class B200Error(RuntimeError):
    def __init__(self, status, text):
        self.status = status
        super().__init__(f"{STATUS_NAMES.get(status, status)}: {text}")


def build(verbose=False):
    """Compile csrc/*.cu for sm_100a into lib/libb200edm.so (nvcc cross-compiles without a GPU)."""
    import subprocess
    out = subprocess.run(["make", "-C", os.path.join(_HERE, "csrc"), "-j4"], capture_output=True, text=True)
    if out.returncode != 0:
        raise RuntimeError("building libb200edm.so failed:\n" + out.stdout[-4000:] + out.stderr[-4000:])
    if verbose:
        print(out.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(there is no CPU fallback)")
        _lib = C.CDLL(LIB_PATH)
        _lib.b200_last_error.restype = C.c_char_p
    return _lib


def check(status):
    if status != 0:
        raise B200Error(status, lib().b200_last_error().decode())


def device_count():
    return lib().b200_device_count()


def set_device(i):
    check(lib().b200_set_device(int(i)))


def synchronize():
    check(lib().b200_synchronize())


def bench_random_gather(table_bytes=512 << 20, n_gathers=100_000_000):
    """Machine ceiling for random 32-byte gathers (returns ms, gathers/s)."""
    ms, rate = C.c_double(), C.c_double()
    check(lib().b200_bench_random_gather(C.c_size_t(table_bytes), C.c_size_t(n_gathers), C.byref(ms), C.byref(rate)))
    return ms.value, rate.value


def bench_fp64_fma():
    """Dense FP64 FMA ceiling of the current device in TFLOP/s."""
    t = C.c_double()
    check(lib().b200_bench_fp64_fma(C.byref(t)))
    return t.value
