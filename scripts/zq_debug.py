import numpy as np, sys, torch
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from helpers import make_geoms
from tomography_alignment_b200 import pose_table
from tomography_alignment_b200.cuda_backend import CudaBackend
from tomography_alignment_b200.phantom import benchmark_poses
n, sel = 256, [3, 100, 181, 300]
g, og = make_geoms((n, n, n), (n, n), len(sel))
phi, alpha, beta, xyz = (a[sel] for a in benchmark_poses(360))
vol_d = torch.rand((n, n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
res = {}
for zq in (True, False):
    be = CudaBackend(g, "cuda:0", zquad=zq)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    o = be.proj_grad(vol_d)
    res[zq] = (o["proj"].cpu().numpy(), o["dproj"].cpu().numpy().reshape(len(sel), 6, n, n))
for k in range(len(sel)):
    dp = res[True][0][k] - res[False][0][k]
    print("view", k, "proj max diff %.3e" % np.abs(dp).max())
    for c in range(6):
        d = res[True][1][k, c] - res[False][1][k, c]
        a = np.abs(d); i = np.unravel_index(a.argmax(), a.shape)
        nb = (a > 1e-3 * np.abs(res[False][1][k, c]).max()).sum()
        print("   comp", c, "max abs diff %.3e at %s ref %.3e  rms diff %.2e rms ref %.2e  n_bad %d" % (a.max(), i, res[False][1][k, c][i], np.sqrt((d ** 2).mean()), np.sqrt((res[False][1][k, c] ** 2).mean()), nb))
