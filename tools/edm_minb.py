import os, sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
import bench
ref = None
for sigma in (0.0, 0.5):
    for minb in ("4", "6", "8", "10"):
        os.environ["B200_EDM_MINB"] = minb
        m = B.EventDrivenMap([bench.BETA], 1000, noNeurons=1024)
        m.SetParameterStdDev(sigma); m.EnableTiming(True); m.SetTuning(8)
        for _ in range(3): f = m.ComputeF(bench.Z_DRIVER)
        ms = []
        for _ in range(5):
            m.ComputeF(bench.Z_DRIVER); ms.append(m.LastEvolveMs())
        zc = np.repeat(bench.Z_DRIVER[:, None], 4, axis=1)
        m.ComputeFBatch(zc); m.ComputeFBatch(zc); ms4 = m.LastEvolveMs()
        print(f"sigma={sigma} minb={minb} evolve 1 eval {np.mean(ms):.3f} ms   4 evals {ms4:.3f} ms ({ms4/4:.3f}/eval)  F={f}", flush=True)
        m.close()
