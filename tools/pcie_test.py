"""What the host link gives: pinned H2D / D2H alone and together (the floor of the host-buffer interp paths)."""
import torch, time
n = 200_000_000
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, reps=5):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / reps
a = t(lambda: d_in.copy_(h_in, non_blocking=True))
b = t(lambda: h_out.copy_(d_out, non_blocking=True))
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
c = t(both)
print(f"H2D 1.6 GB: {a*1e3:.1f} ms ({1.6/a:.1f} GB/s)   D2H 0.8 GB: {b*1e3:.1f} ms ({0.8/b:.1f} GB/s)   both at once: {c*1e3:.1f} ms")
