"""First GPU sanity run: interp1/interp2/edm vs the oracle."""
import sys, time
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
from oracle import oracle_py as O

print("devices", B.device_count())
rng = np.random.default_rng(0)
z0 = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)
# heterogeneous
m = B.EventDrivenMap([np.float32(13.0589)], 8, noNeurons=1024)
m.SetDebugFlag(True); m.SetParameterStdDev(0.5); m.SetSeed(42)
f = m.ComputeF(z0)
beta = m.DebugFetch("beta")
cfg = O.edm_cfg(R=8, N=1024, sigma=0.5, seed=42)
print("beta rng maxdiff", np.abs(beta - O.edm_beta(cfg)).max())
cfg = O.edm_cfg(R=8, N=1024, beta_ext=beta)
fo, a = O.edm_compute_f(cfg, z0, nthreads=8)
print("hetero F", f, fo, np.abs(f - fo).max(), m.DebugFetch("event_count")[0], a["event_count"], m.DebugFetch("accept")[0])
# Jacobian
m = B.EventDrivenMap([np.float32(13.0589)], 4, noNeurons=1024)
J, f0 = m.ComputeDFDU(z0, 1e-2, return_f0=True)
cfg = O.edm_cfg(R=4, N=1024)
Jo, f0o = O.edm_compute_dfdu(cfg, z0, 1e-2, nthreads=8)
print("J", J, "\nJo", Jo, "\nmax rel", np.abs(J - Jo).max() / np.abs(Jo).max())
# default ensemble timing
m = B.EventDrivenMap([np.float32(13.0589)], 1000, noNeurons=1024)
m.EnableTiming(True)
for npt in (0, 4, 8, 16, 2):
    m.SetTuning(npt)
    for _ in range(2):
        f = m.ComputeF(z0)
    print("R=1000 N=1024 npt", npt, "evolve ms", m.LastEvolveMs(), m.LastCounters(), f)
