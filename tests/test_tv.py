"""TV proximal step (tv_denoise.py): the numpy oracle against the reference's own module (golden fixture), the
loop logic of tomography_alignment_b200.tv_denoise on the CPU with numpy stand-ins for the two kernels, and
(gpu tier) the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

from oracle import tv_oracle as TVO
from tomography_alignment_b200 import tv_denoise as TV

GOLD = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_numpy_cases.npz"))
CASES = [(0.05, 7), (0.5, 20), (2.0, 60)]


class NumpyTvOps(object):
    """Test stand-in for CudaTvOps: the two fused kernels written with the oracle's numpy stencils."""
    device = torch.device("cpu")

    def dual_error(self, weight, p, im, err):
        err.copy_(torch.as_tensor(np.float32(weight) * TVO.div(p.numpy()) - im.numpy()))

    def dual_update(self, inv_fw, t_factor, err, aux, gim):
        a = aux.numpy() + TVO.gradient(err.numpy()) * np.float32(inv_fw)
        tmp = a / np.maximum(np.sqrt(np.sum(a ** 2, 0)), 1.)
        aux.copy_(torch.as_tensor(((1 + t_factor) * tmp - t_factor * gim.numpy()).astype(np.float32)))
        gim.copy_(torch.as_tensor(tmp.astype(np.float32)))


def test_oracle_matches_reference_tv_module():
    im = GOLD["tv/im"]
    assert np.array_equal(TVO.div(GOLD["tv/div_in"]), GOLD["tv/div"])
    assert np.array_equal(TVO.gradient(im), GOLD["tv/gradient"])
    assert abs(TVO.tv_norm_3d(im) - float(GOLD["tv/tv_norm_3d"])) < 1e-4
    for w, nit in CASES:
        np.testing.assert_allclose(TVO.denoise_fista(im, weight=w, niter=nit), GOLD["tv/denoise_w%g_n%d" % (w, nit)],
                                   rtol=0, atol=2e-6)


@pytest.mark.parametrize("w,nit", CASES)
def test_denoise_fista_loop_logic_on_cpu(w, nit):
    im = GOLD["tv/im"]
    out = TV.denoise_fista(im, weight=w, niter=nit, ops=NumpyTvOps())
    assert isinstance(out, np.ndarray) and out.dtype == np.float32
    np.testing.assert_allclose(out, GOLD["tv/denoise_w%g_n%d" % (w, nit)], rtol=0, atol=5e-6)
    assert abs(TV.tv_norm_3d(im) - float(GOLD["tv/tv_norm_3d"])) < 1e-4


def test_fista_tv_reconstruction_loop_on_cpu():
    """Host logic of recon.RegularizedRecon.run_fista (regularized.py:57-154) with emulated operators and numpy TV ops."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    from helpers import EmuBackend, make_geoms
    from tomography_alignment_b200 import pose_table
    from tomography_alignment_b200.phantom import shepp3d
    from tomography_alignment_b200.recon import RegularizedRecon
    n, n_proj = 12, 8
    g, _ = make_geoms((n, n, n), (n, n), n_proj)
    phi = np.linspace(0, np.pi, n_proj)
    angles = np.array([phi, 0 * phi, 0 * phi]).T
    truth = shepp3d(n)
    be = EmuBackend(g)
    be.set_poses(pose_table(angles, np.zeros((n_proj, 3)), g.cor_shift))
    b = be.forward(truth).numpy().reshape(n_proj, -1)
    r = RegularizedRecon(g, b, angles, np.zeros((n_proj, 3)), options={"ground_truth": truth}, backend=EmuBackend(g),
                         tv_ops=NumpyTvOps())
    rec, err = r.run_fista(niter=6, hyper=4.0 * n * n_proj, beta_tv=2.0, niter_tv=10)
    assert rec.shape == (n ** 3,) and rec.dtype == np.float32 and len(err) == 6 and np.all(np.diff(err) < 0)
    assert np.all(r.total_cost[:6] >= r.data_fidelity_cost[:6])


@pytest.mark.gpu
@pytest.mark.parametrize("w,nit", CASES)
def test_denoise_fista_cuda_kernels(w, nit):
    im = GOLD["tv/im"]
    out = TV.denoise_fista(im, weight=w, niter=nit)
    np.testing.assert_allclose(out, GOLD["tv/denoise_w%g_n%d" % (w, nit)], rtol=0, atol=5e-6)
    dev = TV.denoise_fista(torch.as_tensor(im).cuda(), weight=w, niter=nit)
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), out)


@pytest.mark.gpu
def test_tv_kernels_against_numpy_stencils_ragged_shape():
    rng = np.random.default_rng(0)
    shape = (7, 33, 45)
    ops = TV.CudaTvOps("cuda:0")
    p = rng.standard_normal((3,) + shape).astype(np.float32)
    im = rng.standard_normal(shape).astype(np.float32)
    err = torch.empty(shape, device="cuda")
    ops.dual_error(0.7, torch.as_tensor(p).cuda(), torch.as_tensor(im).cuda(), err)
    np.testing.assert_allclose(err.cpu().numpy(), np.float32(0.7) * TVO.div(p) - im, rtol=0, atol=2e-6)
    aux, gim = torch.as_tensor(p).cuda().clone(), torch.as_tensor(p[::-1].copy()).cuda()
    ref = NumpyTvOps()
    aux_c, gim_c = torch.as_tensor(p).clone(), torch.as_tensor(p[::-1].copy())
    ops.dual_update(0.3, 0.4, err, aux, gim)
    ref.dual_update(0.3, 0.4, err.cpu(), aux_c, gim_c)
    np.testing.assert_allclose(aux.cpu().numpy(), aux_c.numpy(), rtol=0, atol=3e-6)
    np.testing.assert_allclose(gim.cpu().numpy(), gim_c.numpy(), rtol=0, atol=3e-6)


@pytest.mark.gpu
def test_fista_tv_reconstruction_on_gpu():
    """recon.RegularizedRecon.run_fista (regularized.py:57-154) on a 32^3 phantom: error decreases, TV-regularised
    result is closer to the piecewise-constant phantom than plain SIRT at the same number of operator applications."""
    from tomography_alignment_b200 import Geometry, pose_table
    from tomography_alignment_b200.cuda_backend import CudaBackend
    from tomography_alignment_b200.phantom import shepp3d
    from tomography_alignment_b200.recon import RegularizedRecon
    n, n_proj = 32, 20
    g = Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
    phi = np.linspace(0, np.pi, n_proj)
    angles = np.array([phi, 0 * phi, 0 * phi]).T
    truth = shepp3d(n)
    be = CudaBackend(g, "cuda:0")
    be.set_poses(pose_table(angles, np.zeros((n_proj, 3)), g.cor_shift))
    b = be.forward(torch.as_tensor(truth)).cpu().numpy().reshape(n_proj, -1)
    r = RegularizedRecon(g, b, angles, np.zeros((n_proj, 3)), options={"ground_truth": truth}, device="cuda:0")
    rec, err = r.run_fista(niter=25, hyper=4.0 * n * n_proj, beta_tv=2.0, niter_tv=20)
    assert rec.shape == (n ** 3,) and len(err) >= 3 and err[-1] < err[0], err
    assert np.all(np.isfinite(r.total_cost[:len(err)]))
