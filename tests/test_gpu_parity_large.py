"""Parity at the BASELINE.json sizes: the CUDA path, through the C ABI, against the CPU
oracle -- not against itself -- at 256^3 (configs 1-2), 512^3 (configs 3 / headline) and 1024^3 (config 4).

  256^3    8 views of benchmark_poses(360): full projections, full backprojection, full gradient images
  512^3    4 views of benchmark_poses(720): the same, full outputs (oracle loops spread over the host cores)
  1024^3   3 views of benchmark_poses(1500): a seeded subset of rays (forward, gradient) and of voxels (adjoint);
           the oracle's subset entry points are the same loops (tests/test_oracle_identities.py pins them to the
           full-view ones entry for entry)

Tolerances are BASELINE.json's: relative L2 <= 1e-5 projections / backprojections, <= 1e-4 gradients.  Volumes and
projections are uniform random float32 (no smoothness to hide weight errors behind); the float32 marching state of the
kernels (re-based every 64 samples; tile-local positions in the scatter adjoint) is what grows with size."""
import numpy as np
import pytest
import torch

from oracle import oracle as O
from tomography_alignment_b200 import pose_table
from tomography_alignment_b200.phantom import benchmark_poses
from helpers import make_geoms, rel_l2

pytestmark = pytest.mark.gpu

TOL_PROJ, TOL_GRAD = 1e-5, 1e-4


def _setup(n, n_views_total, sel):
    from tomography_alignment_b200.cuda_backend import CudaBackend
    g, og = make_geoms((n, n, n), (n, n), len(sel))
    phi, alpha, beta, xyz = (a[sel] for a in benchmark_poses(n_views_total))
    be = CudaBackend(g, "cuda:0")
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    return g, og, be, (phi, alpha, beta, xyz)


def _random_inputs(n, n_views, seed):
    gen = torch.Generator(device="cuda").manual_seed(seed)
    vol = torch.rand((n, n, n), device="cuda", generator=gen)
    y = torch.rand((n_views, n, n), device="cuda", generator=gen)
    return vol, y


def _check_full(n, n_total, sel, seed):
    g, og, be, (phi, alpha, beta, xyz) = _setup(n, n_total, sel)
    vol_d, y_d = _random_inputs(n, len(sel), seed)
    vol, y = vol_d.cpu().numpy(), y_d.cpu().numpy()
    all_rays = np.arange(og.n_det)
    errs = {}
    # forward + gradient images + fused residual gradients
    out = be.proj_grad(vol_d, meas=y_d)
    fwd = be.forward(vol_d)
    for k in range(len(sel)):
        p, gr = O.forward_proj_grad_rays(og, alpha[k], beta[k], phi[k], xyz[k], og.cor_shift[k], vol, all_rays)
        errs["forward[%d]" % k] = (rel_l2(fwd[k].cpu().numpy(), p), TOL_PROJ)
        errs["proj(grad kernel)[%d]" % k] = (rel_l2(out["proj"][k].cpu().numpy(), p), TOL_PROJ)
        errs["dproj[%d]" % k] = (rel_l2(out["dproj"][k].cpu().numpy(), gr), TOL_GRAD)
        res = y[k].ravel().astype(np.float64) - p
        errs["grad6[%d]" % k] = (rel_l2(out["grad6"][k].cpu().numpy(), -gr @ res), TOL_GRAD)
        cost = 0.5 * res @ res
        errs["cost[%d]" % k] = (abs(out["cost"][k].item() - cost) / cost, TOL_PROJ)
    del out, fwd
    # exact adjoint: tile-scatter kernel against the oracle's full scatter
    ref = O.adjoint_views_parallel(og, alpha, beta, phi, xyz, y)
    errs["adjoint"] = (rel_l2(be.adjoint(y_d).cpu().numpy(), ref), TOL_PROJ)
    print("\n%d^3:" % n, {k: "%.2e" % v[0] for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v[0] <= v[1]}
    assert not bad, bad


def test_256_cubed_full_outputs_vs_oracle():
    """BASELINE configs 1-2 (256^3 x 360): 8 views spread over the half turn, every output element compared."""
    _check_full(256, 360, [0, 45, 90, 135, 179, 180, 270, 359], seed=256)


def test_512_cubed_full_outputs_vs_oracle():
    """BASELINE config 3 / the headline workload (512^3 x 720): 4 seeded views, every output element compared."""
    sel = sorted(np.random.default_rng(512).choice(720, size=4, replace=False).tolist())
    _check_full(512, 720, sel, seed=512)


def test_1024_cubed_seeded_subsets_vs_oracle():
    """BASELINE config 4 (1024^3 x 1500): 3 seeded views; 8192 seeded rays per view for the forward projector and the
    gradient images, 8192 seeded voxels (plus the 8 volume corners and face centres) for the exact adjoint."""
    n = 1024
    rng = np.random.default_rng(1024)
    sel = sorted(rng.choice(1500, size=3, replace=False).tolist())
    g, og, be, (phi, alpha, beta, xyz) = _setup(n, 1500, sel)
    vol_d, y_d = _random_inputs(n, len(sel), 1024)
    vol, y = vol_d.cpu().numpy(), y_d.cpu().numpy()
    errs = {}
    fwd = be.forward(vol_d)
    out = be.proj_grad(vol_d, meas=y_d)
    for k in range(len(sel)):
        rays = np.sort(rng.choice(og.n_det, size=8192, replace=False))
        p, gr = O.forward_proj_grad_rays(og, alpha[k], beta[k], phi[k], xyz[k], og.cor_shift[k], vol, rays)
        rays_d = torch.as_tensor(rays, device="cuda")
        errs["forward[%d]" % k] = (rel_l2(fwd[k].reshape(-1)[rays_d].cpu().numpy(), p), TOL_PROJ)
        errs["proj(grad kernel)[%d]" % k] = (rel_l2(out["proj"][k].reshape(-1)[rays_d].cpu().numpy(), p), TOL_PROJ)
        errs["dproj[%d]" % k] = (rel_l2(out["dproj"][k][:, rays_d].cpu().numpy(), gr), TOL_GRAD)
    del fwd, out
    voxels = rng.choice(og.n_vox, size=8192, replace=False)
    m = n - 1
    special = [(x * n + yy) * n + z for x in (0, m) for yy in (0, m) for z in (0, m)] + \
              [(x * n + yy) * n + z for (x, yy, z) in ((0, n // 2, n // 2), (m, n // 2, n // 2), (n // 2, 0, n // 2),
                                                       (n // 2, m, n // 2), (n // 2, n // 2, 0), (n // 2, n // 2, m))]
    voxels = np.unique(np.concatenate([voxels, np.array(special)]))
    ref = O.adjoint_voxels(og, alpha, beta, phi, xyz, y, voxels)
    got = be.adjoint(y_d).reshape(-1)[torch.as_tensor(voxels, device="cuda")].cpu().numpy()
    errs["adjoint"] = (rel_l2(got, ref), TOL_PROJ)
    assert np.count_nonzero(ref) > 4096
    print("\n1024^3:", {k: "%.2e" % v[0] for k, v in errs.items()})
    bad = {k: v for k, v in errs.items() if not v[0] <= v[1]}
    assert not bad, bad
