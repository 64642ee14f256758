"""Per-shard kernel times of the 512^3 x 720 benchmark for contiguous (np.array_split) and interleaved view sharding over
8 ranks, measured on ONE GPU (one shard after the other): how much of the multi-GPU step is load imbalance."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tomography_alignment_b200 import Geometry, pose_table
from tomography_alignment_b200.cuda_backend import CudaBackend
from tomography_alignment_b200.phantom import benchmark_poses
n, n_proj, world = 512, 720, 8
g = Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
phi, alpha, beta, xyz = benchmark_poses(n_proj)
poses = pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift)
poses[:, 1:3] *= 0.5; poses[:, 3:6] *= 0.5
be = CudaBackend(g, "cuda:0")
torch.manual_seed(0)
vol = torch.rand((n, n, n), device="cuda")
bp = torch.empty((n, n, n), device="cuda")
def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
out = {}
for name, shards in (("contiguous", np.array_split(np.arange(n_proj), world)), ("interleaved", [np.arange(r, n_proj, world) for r in range(world)])):
    rows = []
    for idx in shards:
        be.set_poses(poses[idx])
        y = torch.rand((len(idx), n, n), device="cuda")
        proj = torch.empty_like(y)
        be.pad(vol)
        rows.append((t(lambda: be.forward(vol, out=proj)), t(lambda: be.adjoint(y, out=bp)),
                     t(lambda: be.proj_grad(vol, meas=y, want_proj=False, want_dproj=False, repad=False))))
    a = np.array(rows)
    out[name] = {"fwd_ms": a[:, 0].round(2).tolist(), "back_ms": a[:, 1].round(2).tolist(), "grad_ms": a[:, 2].round(2).tolist(),
                 "sum_of_max_ms": float(a.max(axis=0).sum()), "max_of_sum_ms": float(a.sum(axis=1).max()), "mean_sum_ms": float(a.sum(axis=1).mean())}
print(json.dumps(out))
