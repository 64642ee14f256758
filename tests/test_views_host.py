"""tomo_views_compute_host (csrc/views.cpp) against the numpy setup of the oracle, which is checked
bit for bit against the reference's numpy (test_oracle_golden.py)."""
import ctypes

import numpy as np
import pytest

from oracle import oracle as O
from tomography_alignment_b200 import _lib, pose_table
from tomography_alignment_b200.projection_operators import full_pose_table, reference_sample_counts
from helpers import make_geoms, random_poses, _P

V = dict(P00=0, U=3, W=6, D=9, N=12, RLEN=13, INVD=14, M=17, E=26, F=35, H=44, K=53, LINV=62, RB=71, VROT=74, VTR=83)


def views_for(g, poses):
    L = _lib.load()
    poses = full_pose_table(g, poses)            # appends the reference's sample count (numpy-evaluated)
    out = np.zeros((poses.shape[0], _lib.VIEW_STRIDE))
    cg = g.to_c()
    _lib.check(L.tomo_views_compute_host(ctypes.byref(cg), _P(poses), poses.shape[0], _P(out)), "views")
    return out


@pytest.mark.parametrize("shape,dshape,cor,step", [((9, 8, 7), (9, 7), 0.0, 1.0), ((8, 8, 8), (10, 6), 0.6, 1.0),
                                                   ((8, 8, 8), (8, 8), -0.3, 0.5)])
def test_lattice_and_derivative_tables(shape, dshape, cor, step):
    n_proj = 4
    g, og = make_geoms(shape, dshape, n_proj, cor=[cor, 0, 0], step=step)
    phi, alpha, beta, xyz = random_poses(n_proj, 11, tilt=0.1, phis=[0.0, 0.8, 1.9, np.pi])
    views = views_for(g, pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    ix, iz = np.meshgrid(np.arange(dshape[0]), np.arange(dshape[1]), indexing="ij")
    ix, iz = ix.ravel(), iz.ravel()
    for i in range(n_proj):
        vs = O.ViewSetup(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i])
        v = views[i]
        p0 = v[V["P00"]:V["P00"] + 3, None] + v[V["U"]:V["U"] + 3, None] * ix + v[V["W"]:V["W"] + 3, None] * iz
        np.testing.assert_allclose(p0, vs.p0, rtol=0, atol=1e-12)
        np.testing.assert_allclose(v[V["D"]:V["D"] + 3, None] * np.ones_like(ix), vs.r_hat * vs.step_size, rtol=0, atol=1e-14)
        assert int(v[V["N"]]) == vs.n                             # int(r_length/step) sits on a rounding edge: passed in
        assert v[V["RLEN"]] == vs.r_length0
        der = vs.der()                                            # (9, 3, n_rays)
        for k in range(3):
            np.testing.assert_allclose(v[V["M"] + 3 * k:V["M"] + 3 * k + 3, None] * np.ones_like(ix), der[k], atol=1e-14)
            e = v[V["E"] + 3 * k:V["E"] + 3 * k + 3, None] + v[V["F"] + 3 * k:V["F"] + 3 * k + 3, None] * ix \
                + v[V["H"] + 3 * k:V["H"] + 3 * k + 3, None] * iz
            np.testing.assert_allclose(e, der[3 + k], rtol=0, atol=1e-11)
            np.testing.assert_allclose(v[V["K"] + 3 * k:V["K"] + 3 * k + 3, None] * np.ones_like(ix),
                                       der[6 + k] * vs.step_size / vs.r_length0, rtol=0, atol=1e-13)
        Lm = np.array([v[V["U"]:V["U"] + 3], v[V["W"]:V["W"] + 3], v[V["D"]:V["D"] + 3]]).T
        np.testing.assert_allclose(v[V["LINV"]:V["LINV"] + 9].reshape(3, 3) @ Lm, np.eye(3), atol=1e-12)
        rot = O.rot_y(beta[i]) @ O.rot_x(alpha[i]) @ O.rot_z(phi[i])
        np.testing.assert_allclose(v[V["VROT"]:V["VROT"] + 9].reshape(3, 3), rot, atol=1e-15)
        np.testing.assert_allclose(v[V["VTR"]:V["VTR"] + 3], O.rot_y(beta[i]) @ xyz[i], atol=1e-15)


def test_argument_errors_do_not_touch_cuda():
    L = _lib.load()
    g, _ = make_geoms((4, 4, 4), (4, 4), 1)
    cg = g.to_c()
    poses = np.zeros((1, _lib.POSE_STRIDE))
    out = np.zeros((1, _lib.VIEW_STRIDE))
    assert L.tomo_views_compute_host(ctypes.byref(cg), None, 1, _P(out)) == -1
    assert L.tomo_views_compute_host(ctypes.byref(cg), _P(poses), 0, _P(out)) == -1
    cg.step_size = 0.0
    assert L.tomo_views_compute_host(ctypes.byref(cg), _P(poses), 1, _P(out)) == -2
    assert b"step_size" in L.tomo_last_error()
    with pytest.raises(_lib.TomoError):
        _lib.check(-2, "tomo_views_compute_host")
    # null-pointer checks come before any CUDA call
    assert L.tomo_forward(ctypes.byref(cg), None, 1, None, None, None) == -1
    assert L.tomo_back_adjoint(ctypes.byref(cg), None, 1, None, None, 0, None) == -1
    assert L.tomo_proj_grad(ctypes.byref(cg), None, 1, None, None, None, None, None, None, None, 0, None) == -1
    assert L.tomo_pad_volume(ctypes.byref(cg), None, None, None) == -1


def test_kernel_selection_flags():
    """Per-view flags the kernels dispatch on: separable (no tilt), tile-scatter colour count, TMA box fit of the
    voxel-driven backprojector; the table-wide counts are replicated in every record (sub-tables stay valid)."""
    NCOL, NUNCOL, SEP, NSEP, VBOK, NVBIG = 86, 87, 146, 147, 148, 149
    n_proj = 6
    g, og = make_geoms((32, 32, 32), (32, 32), n_proj)
    phi = np.array([0.0, 0.5, 1.0, 1.5, 2.0, 2.5])
    alpha = np.array([0.0, 0.0, 0.01, -0.02, 0.1, 0.0])
    beta = np.array([0.0, 0.0, 0.02, 0.01, 0.4, 1.5])          # view 4: z leaks into x' (box too narrow); last view: rays almost along z
    views = views_for(g, pose_table(np.array([phi, alpha, beta]).T, np.zeros((n_proj, 3)), g.cor_shift))
    assert list(views[:, SEP]) == [1, 1, 0, 0, 0, 0] and np.all(views[:, NSEP] == 2)
    assert np.all(views[:4, NCOL] >= 2) and views[5, NCOL] == 0 and np.all(views[:, NUNCOL] == np.sum(views[:, NCOL] == 0))
    # brick footprint: 16 x 16 x 32 voxels under Ry Rx Rz, box 32 x 44 with the z' start rounded down to a multiple of 4
    for i in range(n_proj):
        R = O.rot_y(beta[i]) @ O.rot_x(alpha[i]) @ O.rot_z(phi[i])
        sx = np.abs(R[0]) @ np.array([15, 15, 31.0])
        sz = np.abs(R[2]) @ np.array([15, 15, 31.0])
        assert views[i, VBOK] == float(sx + 3.01 <= 32 and sz + 6.01 <= 44)
    assert list(views[:, VBOK]) == [1, 1, 1, 1, 0, 0] and np.all(views[:, NVBIG] == 2)
    # anisotropic voxels scale the footprint
    g2, _ = make_geoms((32, 32, 32), (32, 32), 1, vox_pix=[2.5, 1.0, 1.0])
    v2 = views_for(g2, pose_table(np.array([[0.0, 0.0, 0.0]]), np.zeros((1, 3)), g2.cor_shift))
    assert v2[0, VBOK] == 0 and v2[0, NVBIG] == 1


@pytest.mark.parametrize("shape,dshape,step", [((16, 16, 16), (16, 16), 1.0), ((33, 9, 35), (33, 35), 1.0), ((5, 40, 3), (5, 3), 1.0),
                                               ((24, 20, 18), (24, 18), 0.5), ((12, 12, 12), (18, 9), 1.7), ((7, 6, 5), (1, 1), 1.0),
                                               ((64, 64, 64), (64, 64), 1.0)])
def test_sample_count_is_the_reference_numpy_value(shape, dshape, step):
    """n = int(r_length[0]/step_size) (ray_voxel_utilities.py:88) sits on an integer edge; the value handed to the C ABI must be the
    one the reference's full-width numpy evaluation produces (oracle.ViewSetup restates those calls one for one), for every pose --
    and when the caller leaves it out (n_samples = 0) the library's own float64 evaluation may differ by at most one."""
    n_proj = 60
    cor = np.zeros((n_proj, 3))
    cor[:, 0] = np.random.default_rng(5).uniform(-1.5, 1.5, n_proj)
    g, og = make_geoms(shape, dshape, n_proj, cor=cor, step=step)
    phi, alpha, beta, xyz = random_poses(n_proj, 23, tilt=0.3, shift=3.0)
    alpha[::3] = 0.0
    beta[::4] = 0.0
    phi[:4] = [0.0, np.pi / 2, np.pi, np.pi / 4]
    poses = pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift)
    cnt = reference_sample_counts(g, poses)
    views = views_for(g, poses)
    legacy = np.zeros((n_proj, _lib.POSE_STRIDE))
    legacy[:, :9] = poses
    views0 = views_for(g, legacy)
    edge = 0
    for i in range(n_proj):
        vs = O.ViewSetup(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i])
        assert int(cnt[i, 0]) == vs.n and cnt[i, 1] == vs.r_length0
        assert int(views[i, V["N"]]) == vs.n
        assert abs(int(views0[i, V["N"]]) - vs.n) <= 1
        edge += int(views0[i, V["N"]]) != vs.n
    print("views whose library-side count differs from numpy's:", edge, "of", n_proj)
