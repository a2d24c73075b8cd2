"""Finite-difference Jacobian sharded over GPUs — one process per GPU, torch.distributed.

The reference forms the Jacobian with n sequential ComputeF calls (NewtonSolver.cpp:181-195,
Stability.cpp:95-109).  Here the n+1 evaluations are one batch of (column, realisation) work
items; rank g evolves a contiguous slice of the items on its GPU (b200_edm_evolve_items_dev),
the restricted front positions + accept flags are exchanged with ONE all-gather (NCCL over
NVLink; a few hundred kB), and every rank forms the masked means in the same fixed order
(b200_edm_reduce_items_dev) — so the Jacobian is bitwise independent of the number of GPUs.
Sharding by items rather than by columns keeps all 8 GPUs busy even for n = 3.  When there are
many columns (the profile map of BASELINE config 5 has n + 1 = 1001) each rank instead owns a
contiguous slice of whole columns, reduces them locally and only the residual columns
(n doubles each) are all-gathered — the exchange shrinks from items x n to ncols x n values.
"""
import numpy as np


def partition_items(n_items, world):
    """Contiguous, equally padded slices: returns (per_rank, [(lo, hi)] * world)."""
    per = (n_items + world - 1) // world
    return per, [(min(r * per, n_items), min((r + 1) * per, n_items)) for r in range(world)]


def fd_columns(u, eps):
    """Evaluation points of the forward-difference Jacobian: columns 0..n-1 are u + eps e_i
    (NewtonSolver.cpp:184-188), column n is u itself."""
    u = np.asarray(u, np.float64).ravel()
    n = u.size
    z = np.repeat(u[:, None], n + 1, axis=1)
    z[np.arange(n), np.arange(n)] += eps
    return np.asfortranarray(z)


def fd_jacobian_from_columns(f_cols, eps):
    """J(:, i) = (F(u + eps e_i) - F(u)) * pow(eps, -1)   (NewtonSolver.cpp:194)."""
    n = f_cols.shape[0]
    return np.asfortranarray((f_cols[:, :n] - f_cols[:, n:n + 1]) * eps ** -1), f_cols[:, n].copy()


class GpuEngine:
    """Evolve / reduce on this rank's B200 through the C-ABI."""

    def __init__(self, parameters, noReal, noNeurons=1024, noFronts=3, precision="f64"):
        import torch
        from .edm import EventDrivenMap
        self.torch = torch
        self.map = EventDrivenMap(parameters, noReal, noNeurons=noNeurons, noFronts=noFronts, precision=precision)
        self.R, self.M = int(noReal), int(noFronts)
        self.device = torch.device("cuda", torch.cuda.current_device())

    def _stream(self):
        return self.torch.cuda.current_stream().cuda_stream

    def evolve(self, z_cols, lo, hi, out):
        """out: (per, M + 1) float64 device tensor; columns 0..M-1 positions, column M accept."""
        t = self.torch
        n = hi - lo
        pos = t.empty((max(n, 1), self.M), dtype=t.float64, device=self.device)
        acc = t.empty((max(n, 1),), dtype=t.int32, device=self.device)
        if n:
            self.map.EvolveItemsDev(z_cols, lo, hi, pos, acc, stream=self._stream())
            out[:n, :self.M] = pos[:n]
            out[:n, self.M] = acc[:n].to(t.float64)

    def reduce(self, z_cols, gathered, n_items):
        t = self.torch
        pos = gathered[:n_items, :self.M].contiguous()
        acc = gathered[:n_items, self.M].to(t.int32).contiguous()
        ncols = z_cols.shape[1]
        f = t.empty((ncols, self.M), dtype=t.float64, device=self.device)
        self.map.ReduceItemsDev(z_cols, pos, acc, f, stream=self._stream())
        return f.cpu().numpy().T  # (n, ncols)

    def empty(self, rows, cols=None):
        return self.torch.zeros((rows, (self.M + 1) if cols is None else cols), dtype=self.torch.float64, device=self.device)

    def columns(self, z_cols, out):
        """Whole columns on this rank: out[c, :] = F(z_cols[:, c]) (host-buffer batch call)."""
        if z_cols.shape[1]:
            f = self.map.ComputeFBatch(z_cols)           # (n, ncols) column-major = [ncols][n] in memory
            out[:z_cols.shape[1]].copy_(self.torch.from_numpy(f.T))

    def set_profile_mode(self, n_coarse):
        self.map.SetProfileMode(n_coarse)
        self.M = self.map.M


class ShardedJacobian:
    """AbstractNonlinearProblem + AbstractNonlinearProblemJacobian over `world` ranks.

    `group` is the torch.distributed module (or None for a single process); `engine` defaults
    to the GPU engine — the CPU tests inject a host engine to exercise the sharding logic with
    the gloo backend."""

    def __init__(self, parameters, noReal, noNeurons=1024, noFronts=3, precision="f64", group=None, engine=None,
                 shard="auto"):
        self.dist = group if (group is not None and group.is_initialized()) else None
        self.world = self.dist.get_world_size() if self.dist else 1
        self.rank = self.dist.get_rank() if self.dist else 0
        self.engine = engine or GpuEngine(parameters, noReal, noNeurons, noFronts, precision)
        self.R, self.M = int(noReal), int(noFronts)
        self.shard = shard  # "items", "columns" or "auto" (columns when there are >= 2 per rank)

    def SetProfileMode(self, n_coarse):
        self.engine.set_profile_mode(n_coarse)
        self.M = self.engine.M

    def _columns_t(self, z_cols):
        """Column sharding; returns the gathered residuals as a [ncols][n] tensor (engine's device)."""
        ncols = z_cols.shape[1]
        per, slices = partition_items(ncols, self.world)
        lo, hi = slices[self.rank]
        n = z_cols.shape[0]
        local = self.engine.empty(per, n)
        self.engine.columns(z_cols[:, lo:hi], local)
        if self.world > 1:
            gathered = self.engine.empty(per * self.world, n)
            self.dist.all_gather_into_tensor(gathered, local)
        else:
            gathered = local
        return gathered[:ncols]

    def _columns(self, z_cols):
        return np.asfortranarray(self._columns_t(z_cols).cpu().numpy().T)

    def _use_columns(self, ncols):
        return self.shard == "columns" or (self.shard == "auto" and ncols >= 2 * self.world and self.world > 1)

    def ComputeFBatch(self, z_cols):
        z_cols = np.asfortranarray(z_cols, np.float64)
        ncols = z_cols.shape[1]
        if self._use_columns(ncols):
            return self._columns(z_cols)
        n_items = ncols * self.R
        per, slices = partition_items(n_items, self.world)
        lo, hi = slices[self.rank]
        local = self.engine.empty(per)
        self.engine.evolve(z_cols, lo, hi, local)
        if self.world > 1:
            gathered = self.engine.empty(per * self.world)
            self.dist.all_gather_into_tensor(gathered, local)
        else:
            gathered = local
        return self.engine.reduce(z_cols, gathered, n_items)

    def ComputeF(self, u):
        return self.ComputeFBatch(np.asarray(u, np.float64).reshape(-1, 1))[:, 0]

    def ComputeDFDU(self, u, eps, return_f0=False):
        z = fd_columns(u, eps)
        n = z.shape[0]
        if self._use_columns(n + 1):
            # difference on the engine's device (same two IEEE operations as NewtonSolver.cpp:194),
            # one transfer of the finished Jacobian
            g = self._columns_t(z)
            Jt = (g[:n] - g[n]) * eps ** -1
            J = np.asfortranarray(Jt.cpu().numpy().T)
            return (J, g[n].cpu().numpy().copy()) if return_f0 else J
        f_cols = self.ComputeFBatch(z)
        J, f0 = fd_jacobian_from_columns(f_cols, eps)
        return (J, f0) if return_f0 else J
