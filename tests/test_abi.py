"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/*.h declares, and fails loudly (never computes on the host) when no B200 is visible."""
import ctypes as C
import glob
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = set()
    for h in glob.glob(os.path.join(ROOT, "include", "*.h")):
        text = re.sub(r"/\*.*?\*/", "", open(h).read(), flags=re.S)
        names |= set(re.findall(r"\b(b200_[a-z0-9_]+)\s*\(", text))
    return sorted(names)


def test_library_exports_every_declared_symbol(b200):
    from armadillocudalinearinterpolation_b200 import _lib
    lib = _lib.lib()
    syms = declared_symbols()
    assert len(syms) >= 40
    missing = [s for s in syms if not hasattr(lib, s)]
    assert not missing, missing


def test_version_and_error_text(b200):
    from armadillocudalinearinterpolation_b200 import _lib
    assert _lib.lib().b200_version() >= 0x000100
    assert isinstance(_lib.lib().b200_last_error(), bytes)


def test_invalid_arguments_are_rejected_without_a_gpu(b200):
    from armadillocudalinearinterpolation_b200 import _lib
    lib = _lib.lib()
    h = C.c_void_p()
    assert lib.b200_interp1_plan_create(0, None, None, C.c_size_t(3), C.byref(h)) == -1
    assert lib.b200_edm_create(None, C.c_size_t(0), 1, 2, 3, 0, C.byref(h)) == -1
    assert b"NULL" in lib.b200_last_error()
    assert lib.b200_edm_set_time_horizon(None, C.c_double(1.0)) == -1


def test_no_cpu_fallback(b200):
    """On a box without a GPU every compute entry point must fail with B200_ERR_NO_DEVICE."""
    if b200.device_count() > 0:
        pytest.skip("a GPU is visible")
    with pytest.raises(b200.B200Error) as e:
        b200.interp1(np.linspace(0, 1, 8), np.zeros(8), np.array([0.5]))
    assert e.value.status == -5
    with pytest.raises(b200.B200Error) as e:
        b200.EventDrivenMap([13.0589], 4)
    assert e.value.status == -5
    with pytest.raises(b200.B200Error):
        b200.Interp2Plan(np.arange(4.0), np.arange(3.0), np.zeros((3, 4)))


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: no product source may reference it."""
    pkg = os.path.join(ROOT, "armadillocudalinearinterpolation_b200")
    bad = []
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", ".h")) or f == "Makefile":
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # comments may cite the oracle as the parity reference; code may not include, import,
                # link or load it
                if re.search(r"oracle_py|liboracle|#include[^\n]*oracle|import oracle|from oracle|oracle_(interp|edm|normal)", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad
