import sys
sys.path.insert(0, "/root/repo")
import numpy as np
import armadillocudalinearinterpolation_b200 as B
z = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)
for N, npt in ((96, 4), (64, 4), (128, 4), (96, 8)):
    m = B.EventDrivenMap([float(np.float32(13.0589))], 6, noNeurons=N)
    m.SetDebugFlag(True); m.SetTuning(npt)
    for rep in range(3):
        f = m.ComputeF(z)
        print(N, npt, m.DebugFetch("event_count")[0], m.DebugFetch("last_index")[0][:2].tolist(), m.DebugFetch("accept")[0], m.LastCounters())
