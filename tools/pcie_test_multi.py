"""The host-link floor of the e2e path when N ranks move their 1.6 GB up + 0.8 GB down AT THE SAME TIME
(torchrun --nproc-per-node N tools/pcie_test_multi.py): the ranks of one box share the host memory system."""
import os, time
import torch, torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("gloo")
n = 200_000_000
h_in = torch.empty(n, dtype=torch.float64).pin_memory()
h_out = torch.empty(n // 2, dtype=torch.float64).pin_memory()
d_in = torch.empty(n, dtype=torch.float64, device="cuda")
d_out = torch.empty(n // 2, dtype=torch.float64, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(s2): h_out.copy_(d_out, non_blocking=True)
both(); torch.cuda.synchronize()
dist.barrier()
t0 = time.perf_counter()
for _ in range(5): both()
torch.cuda.synchronize()
ms = (time.perf_counter() - t0) / 5 * 1e3
t = torch.tensor([ms]); out = [torch.zeros(1) for _ in range(dist.get_world_size())]
dist.all_gather(out, t)
if dist.get_rank() == 0:
    v = [float(o) for o in out]
    w = dist.get_world_size()
    print(f"ranks {w}: 1.6 GB up + 0.8 GB down per rank, concurrently: max {max(v):.1f} ms, min {min(v):.1f} ms; "
          f"aggregate {w * 2.4 / (max(v) * 1e-3):.0f} GB/s; per-rank 1e8-query e2e floor = {max(v):.1f} ms -> {w * 1e8 / (max(v) * 1e-3):.3e} points/s")
dist.destroy_process_group()
