"""B200-native lift / evolve / restrict map and linear-interpolation kernels.

Everything computes on the GPU through lib/libb200edm.so (hand-written sm_100a CUDA behind the
C-ABI of include/*.h).  There is no CPU fallback: importing works anywhere, computing needs a B200.
"""
from ._lib import B200Error, build, device_count, set_device, synchronize  # noqa: F401
from .edm import EventDrivenMap, QUIRK_ACCEPT0_BIAS  # noqa: F401
from .interp import Interp1Plan, Interp2Plan, interp1, interp2  # noqa: F401
