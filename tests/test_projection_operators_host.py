"""Host logic of the drop-in boundary (tomography_alignment_b200/projection_operators.py) without a GPU:
the operator is backed by the oracle here (tests only), so what is tested is the argument handling, the
scipy unbound-method idioms the reference's solvers use, masking, dtypes and error behaviour."""
import numpy as np
import pytest
from scipy import sparse

from oracle import oracle as O
from tomography_alignment_b200 import Geometry, ProjectionMatrix, normalise_poses
from tomography_alignment_b200 import _lib
from helpers import OracleBackend, make_geoms, random_poses, rel_l2


def make_pm(shape=(8, 9, 7), dshape=(8, 7), n_proj=4, cor=None):
    g, og = make_geoms(shape, dshape, n_proj, cor=cor)
    return g, og, ProjectionMatrix(g, precision=np.float32, backend=OracleBackend(g))


def test_geometry_matches_reference_recipe():
    g, og = make_geoms((6, 8, 10), (6, 10), 3, cor=[0.5, 0, 0])
    for name in ("vox_origin", "source_centers", "det_centers", "det_orig", "vox_centers", "cor_shift", "vox_size",
                 "det_size"):
        assert np.array_equal(getattr(g, name), getattr(og, name)), name
    assert g.n_vox == 480 and g.n_det == 60 and g.cor_shift.shape == (3, 3)
    c = g.to_c()
    assert (c.nx, c.ny, c.nz, c.ndx, c.ndz) == (6, 8, 10, 6, 10)
    assert c.src_y == -8.0 and c.det_y == 8.0 and c.det_x0 == -2.5 and c.det_z0 == -4.5


def test_default_poses_and_side_effects():
    g, og, pm = make_pm()
    A = pm.projection_matrix()
    assert pm.n_proj == 4 and A.shape == (4 * g.n_det, g.n_vox)
    np.testing.assert_array_equal(pm.angles[:, 0], np.linspace(0.0, np.pi, 4))
    assert np.all(pm.angles[:, 1:] == 0) and pm.xyz_shift.shape == (4, 3) and pm.voxel_mask is None
    # single view is promoted to length-1 arrays (projection_operators.py:43-48)
    A1 = pm.projection_matrix(phi=0.3, alpha=0.01, beta=-0.02, xyz_shift=np.array([0.1, 0.0, 0.2]))
    assert pm.n_proj == 1 and A1.shape == (g.n_det, g.n_vox) and pm.angles.shape == (1, 3)
    n, ang, xyz = normalise_poses(g, phi=np.array([0.1, 0.2]))
    assert n == 2 and ang.shape == (2, 3) and xyz.shape == (2, 3)


def test_scipy_unbound_method_idioms():
    """recon/sirt.py:33-34,59-61: csr_matrix.dot(A, x); csc_matrix.dot(csr_matrix.transpose(A), y)."""
    g, og, pm = make_pm(cor=[0.3, 0, 0])
    phi, alpha, beta, xyz = random_poses(4, 3)
    A = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    og.cor_shift = g.cor_shift
    ref = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr(np.float64)
    rng = np.random.default_rng(0)
    x = rng.random(g.n_vox).astype(np.float32)
    y = rng.random(4 * g.n_det).astype(np.float32)
    ax = sparse.csr_matrix.dot(A, x)
    assert isinstance(ax, np.ndarray) and ax.dtype == np.float32 and ax.shape == (4 * g.n_det,)
    assert rel_l2(ax, ref @ x) < 1e-6
    assert ax.reshape(4, -1).shape == (4, g.n_det)
    aty = sparse.csc_matrix.dot(sparse.csr_matrix.transpose(A), y)
    assert isinstance(aty, np.ndarray) and aty.dtype == np.float32 and aty.shape == (g.n_vox,)
    assert rel_l2(aty, ref.T @ y) < 1e-6
    aty += 1.0                                   # in-place ops on the result must work (sirt.py:63-64)
    # float64 vectors (scipy line searches hand these in) give float64 back
    assert sparse.csr_matrix.dot(A, x.astype(np.float64)).dtype == np.float64
    # other spellings
    assert rel_l2(A @ x, ref @ x) < 1e-6 and rel_l2(A.T @ y, ref.T @ y) < 1e-6 and rel_l2(A.dot(x), ref @ x) < 1e-6
    assert A.T.shape == (g.n_vox, 4 * g.n_det) and A.T.T.shape == A.shape
    with pytest.raises(ValueError):
        A @ np.zeros(5, np.float32)
    with pytest.raises(ValueError):
        A.T @ np.zeros(5, np.float32)


def test_voxel_mask_drops_columns():
    g, og, pm = make_pm()
    phi, alpha, beta, xyz = random_poses(4, 4)
    mask = np.random.default_rng(1).random(tuple(g.vox_shape)) > 0.4
    A = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz, voxel_mask=mask)
    ref = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr(np.float64, voxel_mask=mask)
    rng = np.random.default_rng(2)
    x, y = rng.random(g.n_vox).astype(np.float32), rng.random(4 * g.n_det).astype(np.float32)
    assert rel_l2(A @ x, ref @ x) < 1e-6
    assert rel_l2(A.T @ y, ref.T @ y) < 1e-6
    assert np.all((A.T @ y)[~mask.ravel()] == 0)


def test_entire_object_masked(capsys):
    g, og, pm = make_pm()
    A = pm.projection_matrix(voxel_mask=np.zeros(tuple(g.vox_shape), bool))
    assert "entire object is masked" in capsys.readouterr().out        # projection_operators.py:64
    assert np.all(A @ np.ones(g.n_vox, np.float32) == 0)
    assert np.all(A.T @ np.ones(4 * g.n_det, np.float32) == 0)


def test_projection_gradient_signature_and_order():
    g, og, pm = make_pm(cor=[0.2, 0, 0])
    rec = np.random.default_rng(3).random(tuple(g.vox_shape))
    proj, grad = pm.projection_gradient(rec, alpha=0.01, beta=-0.015, phi=0.6, xyz_shift=np.array([0.2, 0.0, -0.3]),
                                        cor_shift=g.cor_shift[0])
    assert proj.shape == (g.n_det,) and grad.shape == (6, g.n_det)
    assert proj.dtype == np.float32 and grad.dtype == np.float32
    p, gr = O.projection_gradient(og, rec, 0.01, -0.015, 0.6, np.array([0.2, 0.0, -0.3]), g.cor_shift[0])
    assert rel_l2(proj, p) < 1e-6 and rel_l2(grad, gr) < 1e-6


def test_sirt_iterations_match_csr_pipeline():
    """A few SIRT iterations written exactly like recon/sirt.py:26-67 (W, V, residual, update, positivity)
    give the same iterates with the matrix-free operator as with the reference's CSR matrix."""
    g, og, pm = make_pm((10, 10, 10), (10, 10), 6)
    phi, alpha, beta, xyz = random_poses(6, 5)
    A = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    R = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr(np.float32)
    truth = np.random.default_rng(4).random(g.n_vox).astype(np.float32)
    b = (R @ truth).reshape(6, -1)

    def sirt(M, niter=4):
        W = sparse.csr_matrix.dot(M, np.ones((g.n_vox,), dtype=np.float32))
        V = sparse.csc_matrix.dot(sparse.csr_matrix.transpose(M), np.ones((6 * g.n_det,), dtype=np.float32))
        V[V == 0.] = np.inf
        W[W == 0.] = np.inf
        V, W = 1. / V, 1. / W
        rec = np.zeros(g.n_vox, np.float32)
        for _ in range(niter):
            res = sparse.csr_matrix.dot(M, rec)
            res = b - res.reshape(6, -1)
            bp = sparse.csc_matrix.dot(sparse.csr_matrix.transpose(M), W * res.ravel())
            bp *= V
            rec += bp
            rec[rec < 0.] = 0.
        return rec
    assert rel_l2(sirt(A), sirt(R)) < 1e-5


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    g, _ = make_geoms((4, 4, 4), (4, 4), 2)
    pm = ProjectionMatrix(g)
    with pytest.raises(_lib.TomoError, match="no CPU fallback"):
        pm.projection_matrix()
    with pytest.raises(_lib.TomoError, match="no CPU fallback"):
        pm.projection_gradient(np.zeros((4, 4, 4)), 0.0, 0.0, 0.0, np.zeros(3), np.zeros(3))


def test_package_never_imports_the_oracle():
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tomography_alignment_b200")
    for dp, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, flags=re.M), f
                assert "libtomo_oracle" not in src and "libtomo_emu" not in src, f


def test_operators_from_one_projection_matrix_are_independent():
    """The reference returns independent CSR matrices (projection_operators.py:54-76): a second projection_matrix() call,
    a projection_gradient() call or a different view count must not re-pose an operator returned earlier."""
    g, og, pm = make_pm(n_proj=4)
    phi, alpha, beta, xyz = random_poses(4, 3)
    x = np.random.default_rng(0).random(g.n_vox).astype(np.float32)
    A1 = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    r1 = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    y1 = A1.dot(x)
    A2 = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi + 0.3, xyz_shift=xyz)
    r2 = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi + 0.3, xyz_shift=xyz)
    A3 = pm.projection_matrix(phi=phi[:2], alpha=alpha[:2], beta=beta[:2], xyz_shift=xyz[:2])      # fewer views
    assert A1.shape == (4 * g.n_det, g.n_vox) and A3.shape == (2 * g.n_det, g.n_vox)
    for _ in range(2):                                 # interleaved applications
        assert np.array_equal(A1.dot(x), y1) and A1.dot(x).shape == (4 * g.n_det,)
        assert rel_l2(A2.dot(x), r2.forward(x).ravel()) < 1e-6
        assert rel_l2(A3.dot(x), r1.forward(x)[:2].ravel()) < 1e-6
        yy = np.random.default_rng(1).random(4 * g.n_det).astype(np.float32)
        assert rel_l2(sparse.csc_matrix.dot(sparse.csr_matrix.transpose(A1), yy), r1.adjoint(yy.reshape(4, -1))) < 1e-6
        assert rel_l2(A2.T.dot(yy), r2.adjoint(yy.reshape(4, -1))) < 1e-6
        pm.projection_gradient(x, alpha[0], beta[0], phi[0] + 1.0, xyz[0], g.cor_shift[0])    # re-poses the backend as well
    assert np.array_equal(A1.dot(x), y1)
