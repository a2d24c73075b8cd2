"""world_size-2 gloo tests (CPU) of the Jacobian sharding logic: item partition, padded
all-gather, fixed-order reduction.  The GPU engine is replaced by a host engine built on the
oracle (test infrastructure) so the exchange path runs without a GPU."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Z_DRIVER = np.array([np.float32(0.3310), np.float32(0.6914), np.float32(1.3557)], dtype=np.float64)


class OracleEngine:
    """Host stand-in for parallel.GpuEngine."""

    def __init__(self, R, N, M=3, sigma=0.0, seed=42):
        import torch
        from oracle import oracle_py as O
        self.torch, self.O = torch, O
        self.R, self.N, self.M, self.sigma, self.seed = R, N, M, sigma, seed
        self.evolved = []

    def empty(self, rows, cols=None):
        return self.torch.zeros((rows, (self.M + 1) if cols is None else cols), dtype=self.torch.float64)

    def columns(self, z_cols, out):
        for c in range(z_cols.shape[1]):
            cfg = self.O.edm_cfg(R=self.R, N=self.N, M=self.M, sigma=self.sigma, seed=self.seed)
            f, _ = self.O.edm_compute_f(cfg, z_cols[:, c])
            out[c] = self.torch.from_numpy(f)
            self.evolved.append(("col", c))

    def evolve(self, z_cols, lo, hi, out):
        self.evolved.append([])
        for k, item in enumerate(range(lo, hi)):
            col, r = divmod(item, self.R)
            cfg = self.O.edm_cfg(R=self.R, N=self.N, M=self.M, sigma=self.sigma, seed=self.seed)
            _, a = self.O.edm_compute_f(cfg, z_cols[:, col], r_begin=r, r_end=r + 1)
            out[k, :self.M] = self.torch.from_numpy(a["position"][0])
            out[k, self.M] = float(a["accept"][0])
            self.evolved[-1].append(item)

    def reduce(self, z_cols, gathered, n_items):
        g = gathered[:n_items].numpy()
        ncols = z_cols.shape[1]
        T = 5.0
        f = np.empty((self.M, ncols))
        for c in range(ncols):
            blk = g[c * self.R:(c + 1) * self.R]
            take = blk[:, self.M] == 1.0
            mean = np.zeros(self.M)
            for r in range(self.R):          # fixed order, like the oracle
                if take[r]:
                    mean += blk[r, :self.M]
            mean /= take.sum()
            U = np.concatenate([[0.0], z_cols[1:, c]])
            f[:, c] = (-z_cols[0, c]) * U - mean + z_cols[0, c] * T
        return f


def _worker(rank, world, port, R, N, sigma, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from armadillocudalinearinterpolation_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = OracleEngine(R, N, sigma=sigma)
    sj = parallel.ShardedJacobian([13.0589], R, noNeurons=N, group=dist, engine=eng, shard="items")
    J, f0 = sj.ComputeDFDU(Z_DRIVER, 1e-2, return_f0=True)
    f = sj.ComputeF(Z_DRIVER)
    q.put((rank, J, f0, f, eng.evolved))
    dist.barrier()
    dist.destroy_process_group()


def _worker_columns(rank, world, port, R, N, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from armadillocudalinearinterpolation_b200 import parallel
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    eng = OracleEngine(R, N)
    sj = parallel.ShardedJacobian([13.0589], R, noNeurons=N, group=dist, engine=eng, shard="columns")
    zc = np.stack([Z_DRIVER + [0.002 * i, 0, 0] for i in range(5)], axis=1)   # 5 columns over 2 ranks: ragged
    f = sj.ComputeFBatch(zc)
    q.put((rank, f, len(eng.evolved)))
    dist.barrier()
    dist.destroy_process_group()


def test_column_sharding_world2(oracle):
    """Many-column batches (the profile map's 1001-column Jacobian) shard by whole columns: each rank
    reduces its own columns and only the residual columns are all-gathered."""
    import torch.multiprocessing as mp
    R, N, world = 2, 512, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_columns, args=(r, world, port, R, N, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert np.array_equal(res[0][1], res[1][1])
    assert (res[0][2], res[1][2]) == (3, 2)          # ceil(5/2) columns on rank 0, the rest on rank 1
    cfg = oracle.edm_cfg(R=R, N=N)
    for c in range(5):
        fo, _ = oracle.edm_compute_f(cfg, Z_DRIVER + [0.002 * c, 0, 0])
        assert np.array_equal(res[0][1][:, c], fo)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("R,sigma", [(3, 0.0), (5, 0.4)])
def test_sharded_jacobian_world2_matches_single_process(oracle, R, sigma):
    import torch.multiprocessing as mp
    N, world = 512, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, R, N, sigma, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=300) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # every rank ends with the same Jacobian, bit for bit
    assert np.array_equal(res[0][1], res[1][1]) and np.array_equal(res[0][2], res[1][2])
    assert np.all(np.isfinite(res[0][1]))
    # the work was split without overlap: 4 columns x R items in the Jacobian call + R in ComputeF
    n_items = 4 * R
    per = (n_items + 1) // 2
    assert res[0][4][0] == list(range(0, per)) and res[1][4][0] == list(range(per, n_items))
    # and it equals the sequential column loop of the reference on one process
    cfg = oracle.edm_cfg(R=R, N=N, sigma=sigma, seed=42)
    Jo, f0o = oracle.edm_compute_dfdu(cfg, Z_DRIVER, 1e-2)
    assert np.max(np.abs(res[0][1] - Jo)) < 1e-11 * np.max(np.abs(Jo))
    assert np.allclose(res[0][2], f0o, rtol=0, atol=1e-13)
    assert np.allclose(res[0][3], f0o, rtol=0, atol=1e-13)


def test_partition_and_fd_columns():
    sys.path.insert(0, ROOT)
    from armadillocudalinearinterpolation_b200 import parallel
    for n_items, world in [(4000, 8), (10, 4), (3, 8), (1, 1), (4004, 8)]:
        per, sl = parallel.partition_items(n_items, world)
        assert len(sl) == world and sl[0][0] == 0 and sl[-1][1] == n_items
        assert all(a[1] == b[0] for a, b in zip(sl, sl[1:]))
        assert all(hi - lo <= per for lo, hi in sl)
    z = parallel.fd_columns([1.0, 2.0, 3.0], 0.5)
    assert z.shape == (3, 4) and np.array_equal(z[:, 3], [1, 2, 3]) and z[1, 1] == 2.5 and z[0, 1] == 1.0
    f = np.arange(12.0).reshape(3, 4, order="F")
    J, f0 = parallel.fd_jacobian_from_columns(f, 0.5)
    assert np.array_equal(f0, f[:, 3]) and np.array_equal(J[:, 0], (f[:, 0] - f[:, 3]) * 2.0)
