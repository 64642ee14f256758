"""Angle sharding across GPUs: the single-box replacement for the reference's mpi4py solvers.

The reference's only parallelism is data parallelism over projection angles with the volume
replicated on every rank (recon/sirt_mpi.py:36-72, cgls_mpi.py:36-60, regularized_mpi.py:53-75):
    my_index = np.array_split(np.arange(n_proj), size)[rank]
    forward:   each rank projects its own views, nothing is exchanged
    back:      comm.Allreduce(my_back_proj, rec, op=MPI.SUM)            (sirt_mpi.py:103)
    norms:     comm.allreduce(my_norm**2, op=MPI.SUM)                  (sirt_mpi.py:110)
Here: one process per GPU (torchrun), torch.distributed with NCCL over NVLink (gloo on CPU in the
tests), the same array_split sharding, one all-reduce(sum) of the fp32 volume after the local
backprojection, and a zero-padded all-reduce that assembles per-view results (each view is owned
by exactly one rank, the trick examples/mpi_reconstruct.py:41 uses for the projections).
"""
import numpy as np
import torch
import torch.distributed as dist

from .projection_operators import ProjectionMatrix, normalise_poses


def shard_views(n_proj, world, rank):
    """np.array_split(np.arange(n_proj), world)[rank]: the first n_proj % world ranks get one extra."""
    return np.array_split(np.arange(n_proj), world)[rank]


class ShardedProjector(object):
    """A, A^T and the per-view gradient with views sharded over the ranks of ``group``.

    Every rank passes the *global* pose arrays; the rank keeps its shard (cor_shift rows included,
    sirt_mpi.py:46-49).  Volumes are replicated, projections stay sharded:
        forward(vol)  -> local projections (my_n_proj, ndx, ndz)
        adjoint(y_local) -> A^T y summed over all ranks (replicated)
        proj_grad(vol, meas_local) -> dict with global (n_proj, 6) grad6 / (n_proj,) cost on every rank
    """

    def __init__(self, geometry, alpha=None, beta=None, phi=None, xyz_shift=None, precision=np.float32,
                 group=None, device=None, backend_factory=None):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.geometry = geometry
        self.n_proj, angles, xyz = normalise_poses(geometry, alpha, beta, phi, xyz_shift)
        self.my_index = shard_views(self.n_proj, self.world, self.rank)
        self.my_n_proj = int(np.size(self.my_index))
        self.angles, self.xyz_shift = angles, xyz
        cor = np.asarray(geometry.cor_shift, dtype=np.float64).reshape(-1, 3)
        self._my_cor = cor[self.my_index]
        backend = backend_factory(geometry) if backend_factory is not None else None
        self.pm = ProjectionMatrix(geometry, precision=precision, device=device, backend=backend)
        self.op = None
        if self.my_n_proj > 0:
            self._set_local_poses()

    def _set_local_poses(self):
        # projection_matrix reads geometry.cor_shift[:n]; hand it this rank's rows for the call
        saved = self.geometry.cor_shift
        try:
            self.geometry.cor_shift = self._my_cor
            a = self.angles[self.my_index]
            self.op = self.pm.projection_matrix(phi=a[:, 0], alpha=a[:, 1], beta=a[:, 2],
                                                xyz_shift=self.xyz_shift[self.my_index])
        finally:
            self.geometry.cor_shift = saved

    def _allreduce(self, t):
        if self.world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
        return t

    def forward(self, vol):
        """Local rows of A vol (torch tensor in -> torch tensor out on the same device)."""
        return self.op._backend.forward(vol)

    def adjoint(self, y_local, out=None):
        """sum over ranks of A_rank^T y_rank, replicated on every rank (sirt_mpi.py:101-103)."""
        v = self.op._backend.adjoint(y_local, out=out)
        return self._allreduce(v)

    def residual_norm2(self, res_local):
        """sum over ranks of ||res||^2 as a float64 scalar tensor (sirt_mpi.py:110)."""
        s = (res_local.double() ** 2).sum().reshape(1)
        return self._allreduce(s)[0]

    def proj_grad(self, vol, meas_local, want_dproj=False):
        """Per-view 6-DOF gradients of all views on every rank: local views are computed, written into a
        zero (n_proj, 7) float64 table at their global rows, and the table is all-reduced."""
        be = self.pm._get_backend()
        out = be.proj_grad(vol, meas=meas_local, want_proj=True, want_dproj=want_dproj)
        table = torch.zeros((self.n_proj, 7), dtype=torch.float64, device=out["grad6"].device)
        idx = torch.as_tensor(self.my_index, device=table.device, dtype=torch.long)
        table[idx, :6] = out["grad6"]
        table[idx, 6] = out["cost"]
        self._allreduce(table)
        out["grad6_all"] = table[:, :6]
        out["cost_all"] = table[:, 6]
        return out
