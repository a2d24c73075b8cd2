"""Map evaluation time vs number of rings (which build of the evolve kernel the launcher picks)."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
import bench
for sigma in (0.0, 0.5):
    for R in (1, 148, 500, 592, 593, 1000):
        m = B.EventDrivenMap([bench.BETA], R, noNeurons=1024)
        m.SetParameterStdDev(sigma)
        for _ in range(5): f = m.ComputeF(bench.Z_DRIVER)
        B.synchronize(); t = time.perf_counter()
        for _ in range(20): f = m.ComputeF(bench.Z_DRIVER)
        print(f"sigma={sigma} R={R:5d} {(time.perf_counter() - t) / 20 * 1e3:.3f} ms  F={f}", flush=True)
        m.close()
