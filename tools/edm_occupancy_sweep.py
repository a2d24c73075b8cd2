"""Evolve-kernel time as a function of the number of rings: separates the serial chain of one ring
from the throughput limit of an SM (148 SMs x 8 resident rings)."""
import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
import bench
for R in (1, 74, 148, 296, 444, 592, 888, 1184, 2368, 4736):
    m = B.EventDrivenMap([bench.BETA], R, noNeurons=1024)
    m.EnableTiming(True)
    for _ in range(3): m.ComputeF(bench.Z_DRIVER)
    ms = []
    for _ in range(5):
        m.ComputeF(bench.Z_DRIVER); ms.append(m.LastEvolveMs())
    t = float(np.median(ms))
    print(f"R={R:5d}  rings/SM={R/148:6.2f}  evolve {t:7.3f} ms   per ring-event {t*1e-3*1.965e9/848:8.0f} cycles wall, "
          f"{t*1e-3*1.965e9*148/(848*R):8.0f} SM-cycles per ring-event", flush=True)
    m.close()
for R in (1, 1000):
    m = B.EventDrivenMap([bench.BETA], R, noNeurons=1024)
    m.EnableTiming(True)
    for _ in range(3): m.ComputeF(bench.Z_DRIVER)
    ph = m.LastPhaseCycles(); ev = 848
    print("R", R, "cycles per event, ring 0 thread 0:", {k: round(v / ev) for k, v in ph.items()}, "sum", round(sum(ph.values()) / ev))
