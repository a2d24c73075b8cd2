"""A/B of two builds of the library on the map kernel: python tools/edm_ab.py <lib.so>."""
import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from armadillocudalinearinterpolation_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = sys.argv[1]
import armadillocudalinearinterpolation_b200 as B
import bench
print("lib", _lib.LIB_PATH)
for sigma in (0.0, 0.5):
    m = B.EventDrivenMap([bench.BETA], 1000, noNeurons=1024)
    m.SetParameterStdDev(sigma)
    for _ in range(5): f = m.ComputeF(bench.Z_DRIVER)
    B.synchronize(); t = time.perf_counter()
    for _ in range(20): f = m.ComputeF(bench.Z_DRIVER)
    wall = (time.perf_counter() - t) / 20 * 1e3
    m.EnableTiming(True)
    ms = []
    for _ in range(5):
        m.ComputeF(bench.Z_DRIVER); ms.append(m.LastEvolveMs())
    print(f"sigma={sigma} wall/ComputeF {wall:.3f} ms (timing off)  evolve {np.median(ms):.3f} ms (timing on)  F={f}", flush=True)
    m.close()
for R in (1, 148):
    m = B.EventDrivenMap([bench.BETA], R, noNeurons=1024)
    for _ in range(5): f = m.ComputeF(bench.Z_DRIVER)
    B.synchronize(); t = time.perf_counter()
    for _ in range(20): f = m.ComputeF(bench.Z_DRIVER)
    print(f"R={R} wall/ComputeF {(time.perf_counter() - t) / 20 * 1e3:.3f} ms")
    m.close()
