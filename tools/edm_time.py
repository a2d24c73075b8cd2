import sys, time
import os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import armadillocudalinearinterpolation_b200 as B
import bench
for sigma in (0.0, 0.5):
    for npt in (4, 8, 16):
        m = B.EventDrivenMap([bench.BETA], 1000, noNeurons=1024)
        m.SetParameterStdDev(sigma); m.EnableTiming(True); m.SetTuning(npt)
        for _ in range(3): f = m.ComputeF(bench.Z_DRIVER)
        ms = []
        for _ in range(5):
            m.ComputeF(bench.Z_DRIVER); ms.append(m.LastEvolveMs())
        print(f"sigma={sigma} npt={npt} evolve {np.mean(ms):.3f} ms  F={f} {m.LastCounters()}", flush=True)
        m.close()
m = B.EventDrivenMap([bench.BETA], 1000, noNeurons=1024)
m.EnableTiming(True)
for _ in range(3): J = m.ComputeDFDU(bench.Z_DRIVER, 1e-2)
t = time.perf_counter()
for _ in range(5): J = m.ComputeDFDU(bench.Z_DRIVER, 1e-2)
print("jacobian (4 evals, 4000 CTAs) ms", (time.perf_counter() - t) / 5 * 1e3, "evolve", m.LastEvolveMs())
