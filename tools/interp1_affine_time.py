"""interp1, 1e6 linspace knots: affine path (16-byte value pairs) vs segment records (B200_INTERP_AFFINE=0)."""
import os, sys
sys.path.insert(0, "/root/repo")
import numpy as np, torch
import armadillocudalinearinterpolation_b200 as B
ng = 1_000_000
xg = np.linspace(0.0, 1.0, ng); yg = np.sin(2 * np.pi * xg)
g1 = torch.Generator(device="cuda").manual_seed(1236)
ref = {}
for ni, nbuf, reps in ((10_000_000, 8, 200), (100_000_000, 2, 10)):
    qs_u = [torch.rand(ni, generator=g1, device="cuda", dtype=torch.float64) for _ in range(nbuf)]
    for order in ("unsorted", "sorted"):
        qs = qs_u if order == "unsorted" else [q.sort().values for q in qs_u]
        outs = [torch.empty_like(q) for q in qs]
        for affine in ("1", "0"):
            os.environ["B200_INTERP_AFFINE"] = affine
            p1 = B.Interp1Plan(xg, yg)
            for i in range(nbuf * 2): p1(qs[i % nbuf], out=outs[i % nbuf])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps): p1(qs[i % nbuf], out=outs[i % nbuf])
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            key = (ni, order)
            same = ""
            if key in ref: same = "  same bits: " + str(torch.equal(ref[key].view(torch.int64), outs[0].view(torch.int64)))
            else: ref[key] = outs[0].clone()
            print(f"ni={ni:.0e} {order:8s} affine={affine}: {ms * 1e3:8.1f} us  frac {(16 * ni + 16 * ng) / ms / 1e6 / 6537:.3f}{same}", flush=True)
            del p1
    del qs_u, qs, outs
