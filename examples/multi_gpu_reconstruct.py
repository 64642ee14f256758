"""Single-box replacement of examples/mpi_reconstruct.py: one process per GPU (torchrun), views sharded with
np.array_split, volume replicated, NCCL all-reduce of the backprojection (recon/sirt_mpi.py semantics).

    torchrun --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 examples/multi_gpu_reconstruct.py --size 128 --views 180
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomography_alignment_b200 import geometry                                  # noqa: E402
from tomography_alignment_b200.phantom import benchmark_poses, shepp3d           # noqa: E402
from tomography_alignment_b200.recon import CGLS, SIRT                           # noqa: E402
from tomography_alignment_b200.sharding import ShardedProjector                  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=128)
    ap.add_argument("--views", type=int, default=180)
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--solver", default="sirt", choices=["sirt", "cgls"])
    a = ap.parse_args()
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank = dist.get_rank() if dist.is_initialized() else 0
    n, n_proj = a.size, a.views
    geom = geometry.Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    truth = shepp3d(n)
    # every rank simulates its own views; the all-reduce below assembles them (mpi_reconstruct.py:33-41)
    sp = ShardedProjector(geom, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz, device="cuda:%d" % local)
    mine = sp.forward(torch.as_tensor(truth).cuda())
    proj = torch.zeros((n_proj, n, n), device=mine.device)
    proj[torch.as_tensor(sp.my_index, device=mine.device)] = mine
    if dist.is_initialized():
        dist.all_reduce(proj)
    cls = SIRT if a.solver == "sirt" else CGLS
    group = dist.group.WORLD if dist.is_initialized() else None
    solver = cls(geom, proj.cpu().numpy().reshape(n_proj, -1), np.array([phi, alpha, beta]).T, xyz,
                 options={"ground_truth": truth}, group=group, device="cuda:%d" % local)
    rec, err = solver.run_main_iteration(niter=a.iters) if a.solver == "cgls" else \
        solver.run_main_iteration(niter=a.iters, positivity=True)
    if rank == 0:
        print("%s on %d GPU(s): %d iterations, RMSE %.4f -> %.4f" % (a.solver, solver.world, len(err), err[0], err[-1]))
    if dist.is_initialized():
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
