"""Run the tile-scatter backprojector on a small multi-tile case (under compute-sanitizer --tool racecheck)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tomography_alignment_b200 import Geometry, pose_table
from tomography_alignment_b200.cuda_backend import CudaBackend
from tomography_alignment_b200.phantom import benchmark_poses

shape, dshape, n_proj = (40, 36, 66), (44, 70), 6
g = Geometry(n_proj, np.array(shape), np.ones(3), np.array(dshape), np.ones(2))
phi, alpha, beta, xyz = benchmark_poses(n_proj)
phi = np.array([0.0, 0.5, 0.785, 1.3, 2.2, 3.0])
be = CudaBackend(g, "cuda:0")
be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
y = torch.rand((n_proj,) + dshape, device="cuda")
a = be.adjoint(y)
b = be.adjoint(y, gather=True)
torch.cuda.synchronize()
print("rel diff tile vs gather: %.3e" % ((a - b).norm() / b.norm()).item())
