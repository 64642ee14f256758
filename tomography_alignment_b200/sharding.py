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


def shard_views(n_proj, world, rank, mode="contiguous"):
    """Views of rank ``rank``.
    "contiguous"  np.array_split(np.arange(n_proj), world)[rank], the reference's split (recon/sirt_mpi.py:40): the first
                  n_proj % world ranks get one extra view;
    "interleaved" views rank, rank + world, ...: every rank sees the whole angular range.  The cost of a view depends on its
                  angle (rays at 45 degrees cross more tile rows than axis-parallel ones: +6 % on the backprojector), so contiguous
                  blocks of a half turn are unevenly expensive and every all-reduce waits for the slowest rank: measured 2.7 ms
                  of 88 ms at 8 GPUs (profiles/r2_shard_balance.json).  Same sums either way."""
    if mode == "interleaved":
        return np.arange(rank, n_proj, world)
    if mode != "contiguous":
        raise ValueError("shard mode must be 'contiguous' or 'interleaved'")
    return np.array_split(np.arange(n_proj), world)[rank]


def check_world(n_proj, world):
    """Every rank must own at least one view: a rank with an empty shard would raise on its own while its peers wait in
    the next collective.  All ranks know n_proj and the world size, so all of them raise together."""
    if world > n_proj:
        raise ValueError("%d ranks for %d views: every rank needs at least one view (use a smaller process group)"
                         % (world, n_proj))


def adjoint_allreduce(backend, y_local, out=None, group=None, n_slabs=1, reduce=True):
    """vol = sum over ranks of A_rank^T y_rank with the collective hidden behind the kernel: the volume is backprojected in
    ``n_slabs`` x-slabs (tomo_back_adjoint_slab; an x-slab of [nx][ny][nz] is contiguous) and the all-reduce of slab k is
    queued asynchronously as soon as its kernel is, so NCCL sums slab k over NVLink while slab k+1 is computed.
    Replaces the blocking comm.Allreduce(my_back_proj, rec) of recon/sirt_mpi.py:103.
    Returns (volume, works); the volume is complete once every ``work.wait()`` has been called (stream-ordered, the host
    does not block).  ``group=None`` is the default process group; ``reduce=False``: no collective, one launch."""
    reduce = reduce and dist.is_initialized() and dist.get_world_size(group) > 1
    if not reduce or not hasattr(backend, "slabs"):
        v = backend.adjoint(y_local, out=out)
        if reduce:
            dist.all_reduce(v, op=dist.ReduceOp.SUM, group=group)
        return v, []
    if out is None:
        out = torch.empty(backend.vol_shape, dtype=torch.float32, device=backend.device)
    y_local = backend._as_proj(y_local)
    vol3 = out.reshape(backend.vol_shape)
    works = []
    # Slab launches alternate between two side streams: a slab is only ~3 waves of tiles, so on one stream every launch would
    # end with a mostly empty wave (measured: +10 % on the adjoint at 8 slabs); on two streams the next slab's tiles fill the
    # SMs the previous one is draining.  The all-reduce of a slab is queued under its stream, i.e. behind its kernel only.
    slabs = backend.slabs(n_slabs)
    if len(slabs) == 1:
        # one launch on the caller's stream: NCCL's stream waits for it, the caller's stream does not wait for NCCL, so whatever
        # the caller queues next (bench.py: the gradient kernel) runs while the volume is summed.  On a side stream the two
        # kernels would share the SMs and finish together, leaving the all-reduce exposed (measured: +2.4 ms at 8 GPUs)
        backend.adjoint(y_local, out=out)
        return out, [dist.all_reduce(vol3, op=dist.ReduceOp.SUM, group=group, async_op=True)]
    cur = torch.cuda.current_stream(backend.device)
    pool = backend.slab_streams()
    ready = torch.cuda.Event()
    ready.record(cur)                       # y_local / out as the caller's stream left them
    first = None
    for k, (x0, x1) in enumerate(slabs):
        st = pool[k % len(pool)]
        st.wait_event(ready)
        if first is not None:
            st.wait_event(first)            # the first slab fills the separable adjoint's workspace for the others
        with torch.cuda.stream(st):
            backend.adjoint(y_local, out=out, x_range=(x0, x1))
            if first is None:
                first = torch.cuda.Event()
                first.record(st)
            works.append(dist.all_reduce(vol3[x0:x1], op=dist.ReduceOp.SUM, group=group, async_op=True))
    return out, works


class SharedHostBuffer(object):
    """A host array all ranks of the box see (POSIX shared memory), page-locked in every process so that each rank's
    device<->host copies into its own part are asynchronous.  The single-box counterpart of what the reference's MPI scripts
    do with Allreduce / Gather into rank 0's numpy arrays (examples/mpi_reconstruct.py:41): each rank moves only its 1/N of
    the bytes over its own PCIe link and rank 0 (or any rank) reads the whole array from host memory."""

    def __init__(self, name, shape, dtype=torch.float32, group=None):
        import os
        self.path = os.path.join("/dev/shm", name)
        rank = dist.get_rank(group) if dist.is_initialized() else 0
        n = int(np.prod(shape))
        if rank == 0:
            if os.path.exists(self.path):
                os.remove(self.path)
            t = torch.from_file(self.path, shared=True, size=n, dtype=dtype)
        if dist.is_initialized():
            dist.barrier(group)
        if rank != 0:
            t = torch.from_file(self.path, shared=True, size=n, dtype=dtype)
        self.tensor = t.reshape(shape)
        self.pinned = False
        if torch.cuda.is_available():
            rc = torch.cuda.cudart().cudaHostRegister(t.data_ptr(), t.numel() * t.element_size(), 0)
            self.pinned = int(rc) == 0
            if not self.pinned:
                # registration can be refused (locked-memory limits, very large buffers): the buffer still works, copies are
                # then staged by the driver.  The failed call leaves a non-sticky error behind that the next CUDA call would
                # report: fetch it.
                try:
                    import ctypes
                    ctypes.CDLL("libcudart.so.12").cudaGetLastError()
                except OSError:
                    pass
        self._owner = rank == 0

    def close(self):
        import os
        if self.pinned:
            torch.cuda.cudart().cudaHostUnregister(self.tensor.data_ptr())
            self.pinned = False
        if self._owner and os.path.exists(self.path):
            os.remove(self.path)


class ShardedProjector(object):
    """A, A^T and the per-view gradient with views sharded over the ranks of ``group``.

    Every rank passes the *global* pose arrays; the rank keeps its shard (cor_shift rows included,
    sirt_mpi.py:46-49).  Volumes are replicated, projections stay sharded:
        forward(vol)  -> local projections (my_n_proj, ndx, ndz)
        adjoint(y_local) -> A^T y summed over all ranks (replicated)
        proj_grad(vol, meas_local) -> dict with global (n_proj, 6) grad6 / (n_proj,) cost on every rank
    """

    def __init__(self, geometry, alpha=None, beta=None, phi=None, xyz_shift=None, precision=np.float32,
                 group=None, device=None, backend_factory=None, shard="contiguous"):
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.geometry = geometry
        self.n_proj, angles, xyz = normalise_poses(geometry, alpha, beta, phi, xyz_shift)
        check_world(self.n_proj, self.world)
        self.my_index = shard_views(self.n_proj, self.world, self.rank, shard)
        self.my_n_proj = int(np.size(self.my_index))
        self.angles, self.xyz_shift = angles, xyz
        cor = np.asarray(geometry.cor_shift, dtype=np.float64).reshape(-1, 3)
        self._my_cor = cor[self.my_index]
        backend = backend_factory(geometry) if backend_factory is not None else None
        self.pm = ProjectionMatrix(geometry, precision=precision, device=device, backend=backend)
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
        self.op._bind()
        return self.op._backend.forward(vol)

    def adjoint(self, y_local, out=None, n_slabs=1, wait=True):
        """sum over ranks of A_rank^T y_rank, replicated on every rank (sirt_mpi.py:101-103).

        The volume is backprojected in x-slabs and the all-reduce of a finished slab runs (on NCCL's stream, over NVLink)
        while the next slab is computed (``adjoint_allreduce``).  ``wait=False`` returns ``(volume, works)`` so the caller
        can queue independent kernels before ``for w in works: w.wait()``."""
        self.op._bind()
        v, works = adjoint_allreduce(self.op._backend, y_local, out, self.group, n_slabs, reduce=self.world > 1)
        if not wait:
            return v, works
        for w in works:
            w.wait()
        return v

    def residual_norm2(self, res_local):
        """sum over ranks of ||res||^2 as a float64 scalar tensor (sirt_mpi.py:110)."""
        s = (res_local.double() ** 2).sum().reshape(1)
        return self._allreduce(s)[0]

    def proj_grad(self, vol, meas_local, want_dproj=False):
        """Per-view 6-DOF gradients of all views on every rank: local views are computed, written into a
        zero (n_proj, 7) float64 table at their global rows, and the table is all-reduced."""
        self.op._bind()
        be = self.op._backend
        out = be.proj_grad(vol, meas=meas_local, want_proj=True, want_dproj=want_dproj)
        table = torch.zeros((self.n_proj, 7), dtype=torch.float64, device=out["grad6"].device)
        idx = torch.as_tensor(self.my_index, device=table.device, dtype=torch.long)
        table[idx, :6] = out["grad6"]
        table[idx, 6] = out["cost"]
        self._allreduce(table)
        out["grad6_all"] = table[:, :6]
        out["cost_all"] = table[:, 6]
        return out
