"""Known-answer identities that pin the semantics of the live path."""
import numpy as np
import pytest

from oracle import oracle as O
from helpers import make_geoms, random_poses


def test_phi0_projection_is_sum_over_y():
    """phi = alpha = beta = 0, t = 0: samples sit at integer x, z and half-integer y, so
    proj[ix, iz] = sum_y rec[ix, y, iz] exactly (both border half-weights included)."""
    _, og = make_geoms((12, 10, 9), (12, 9), 1)
    rec = np.random.default_rng(0).random((12, 10, 9))
    op = O.OracleOperator(og, phi=np.array([0.0]))
    np.testing.assert_allclose(op.forward(rec)[0].reshape(12, 9), rec.sum(axis=1), rtol=0, atol=1e-12)


def test_adjointness():
    _, og = make_geoms((10, 11, 12), (10, 12), 5, cor=[0.3, 0, 0])
    phi, alpha, beta, xyz = random_poses(5, 1)
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    rng = np.random.default_rng(2)
    x, y = rng.random(og.n_vox), rng.random((5, og.n_det))
    assert abs(np.vdot(op.forward(x), y) - np.vdot(x, op.adjoint(y))) < 1e-10 * abs(np.vdot(x, op.adjoint(y)))


def test_row_sums_are_chord_lengths():
    """A 1 = in-volume chord length per ray: ny at phi = 0 (this is SIRT's W, recon/sirt.py:33)."""
    _, og = make_geoms((8, 14, 8), (8, 8), 1)
    op = O.OracleOperator(og, phi=np.array([0.0]))
    np.testing.assert_allclose(op.forward(np.ones(og.n_vox))[0], 14.0, rtol=0, atol=1e-12)


def test_gradient_matches_finite_differences():
    """Central differences as in utilities/alignment_functions.py:225-241,424-445, on a smooth volume."""
    _, og = make_geoms((14, 14, 14), (14, 14), 1)
    c = np.arange(14) - 6.5
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    rec = np.exp(-(X ** 2 + 1.3 * Y ** 2 + 0.8 * Z ** 2) / 18.0)
    th0 = np.array([0.21, -0.13, 0.33, 0.7, 0.012, -0.017])       # tx ty tz phi alpha beta

    def proj(th):
        p, _ = O.forward_proj_grad(og, th[4], th[5], th[3], th[:3], np.zeros(3), rec)
        return p
    _, g = O.forward_proj_grad(og, th0[4], th0[5], th0[3], th0[:3], np.zeros(3), rec)
    for k in range(6):
        e = np.zeros(6); e[k] = 1e-5
        fd = (proj(th0 + e) - proj(th0 - e)) / 2e-5
        # the interpolant is only piecewise smooth: compare in the L2 sense
        assert np.linalg.norm(fd - g[k]) <= 2e-3 * max(np.linalg.norm(g[k]), 1.0), k


def test_translation_along_beam_leaves_projection_unchanged():
    """examples/generate_data.py:20-23: motion along the beam does not affect the projection -- as long
    as no sample leaves the marching range.  ty shifts by an integer number of steps here."""
    _, og = make_geoms((10, 10, 10), (10, 10), 1)
    rec = np.random.default_rng(3).random((10, 10, 10))
    a = O.OracleOperator(og, phi=np.array([0.0]), xyz_shift=np.array([[0.3, 0.0, -0.2]])).forward(rec)
    b = O.OracleOperator(og, phi=np.array([0.0]), xyz_shift=np.array([[0.3, 2.0, -0.2]])).forward(rec)
    np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)


@pytest.mark.parametrize("w32", [False, True])
def test_matrix_free_equals_reference_csr(w32):
    """The reference's actual pipeline (COO -> CSR with duplicates summed, projection_operators.py:54-76)
    against the matrix-free application."""
    _, og = make_geoms((8, 9, 7), (8, 7), 4, cor=[0.25, 0, 0])
    phi, alpha, beta, xyz = random_poses(4, 5)
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz, w32=w32)
    A = op.csr(np.float32 if w32 else np.float64)
    assert A.shape == (4 * og.n_det, og.n_vox)
    rng = np.random.default_rng(6)
    x, y = rng.random(og.n_vox), rng.random(4 * og.n_det)
    tol = 2e-6 if w32 else 1e-12
    np.testing.assert_allclose(A.astype(np.float64) @ x, op.forward(x).ravel(), rtol=tol, atol=tol)
    np.testing.assert_allclose(A.astype(np.float64).T @ y, op.adjoint(y.reshape(4, -1)), rtol=tol, atol=tol)


def test_csr_nnz_per_voxel_matches_survey():
    """Sanity bound on the CSR size: at most 8 corners per sample and about one sample per voxel per view
    (SURVEY.md section 6 quotes 7.3-7.9 per voxel before duplicates are merged at larger sizes)."""
    _, og = make_geoms((16, 16, 16), (16, 16), 3)
    phi, alpha, beta, xyz = random_poses(3, 8, shift=0.5, phis=[0.3, 1.0, 2.2])
    A = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr()
    assert 3.0 < A.nnz / (3 * og.n_vox) < 8.5


def test_shard_views_is_array_split():
    for n, w in [(90, 8), (7, 3), (4, 8), (1500, 8)]:
        parts = [O.shard_views(n, w, r) for r in range(w)]
        assert np.array_equal(np.concatenate(parts), np.arange(n))
        sizes = [len(p) for p in parts]
        assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)


@pytest.mark.parametrize("shape,dshape,kw", [((14, 12, 13), (14, 13), dict()), ((12, 12, 12), (16, 10), dict(tilt=0.25, shift=4.0)),
                                             ((10, 16, 9), (10, 9), dict(step=0.5, cor=[0.4, 0, 0])),
                                             ((10, 10, 10), (10, 10), dict(step=1.7)), ((9, 9, 9), (9, 9), dict(tilt=0.0, shift=0.0))])
def test_subset_variants_match_the_full_loops(shape, dshape, kw):
    """The ray / voxel subset entry points used by the 256^3 - 1024^3 GPU parity tests are the same loops as the full-view
    oracle: rows of A x, rows of the gradient image, and entries of A^T y (the latter collected voxel by voxel from the
    candidate samples around it -- must equal the scatter entry for entry, including border voxels and rays that miss)."""
    n_proj = 4
    pk = {k: kw[k] for k in ("tilt", "shift") if k in kw}
    g, og = make_geoms(shape, dshape, n_proj, cor=kw.get("cor"), step=kw.get("step", 1.0))
    phi, alpha, beta, xyz = random_poses(n_proj, 4, **pk)
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    rng = np.random.default_rng(8)
    vol = rng.random(shape).astype(np.float32)
    y = rng.random((n_proj, og.n_det)).astype(np.float32)
    rays = rng.choice(og.n_det, size=min(60, og.n_det), replace=False)
    full = op.forward(vol.astype(np.float64))
    for i in range(n_proj):
        sub = O.forward_rays(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol, rays)
        np.testing.assert_allclose(sub, full[i][rays], rtol=0, atol=1e-12)
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        ps, gs = O.forward_proj_grad_rays(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol, rays)
        np.testing.assert_allclose(ps, p[rays], rtol=0, atol=1e-12)
        np.testing.assert_allclose(gs, gr[:, rays], rtol=0, atol=1e-10)
    ref = op.adjoint(y.astype(np.float64))
    np.testing.assert_allclose(O.adjoint_views_parallel(og, alpha, beta, phi, xyz, y), ref, rtol=0, atol=1e-11)
    voxels = np.arange(og.n_vox)                       # every voxel: borders, corners, voxels no ray touches
    np.testing.assert_allclose(O.adjoint_voxels(og, alpha, beta, phi, xyz, y, voxels), ref, rtol=0, atol=1e-11)


def test_voxel_driven_coo_matrix_reproduces_the_splat():
    """bilinear_sparse (src/vox_wt_grad.f90:58-112) emits the matrix S whose product with the volume is the det_img that
    bilinear_vox_interp (:1-55) accumulates directly: the two restatements must agree, tap for tap."""
    from scipy import sparse
    g, og = make_geoms((9, 8, 10), (11, 9), 3, cor=[0.3, 0.0, -0.2])
    phi, alpha, beta, xyz = random_poses(3, 14, tilt=0.1, shift=2.5)
    rec = np.random.default_rng(1).random(og.n_vox)
    for i in range(3):
        dat, det, wts = O.voxel_forward_sparse(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i])
        assert wts.dtype == np.float32 and dat.min() >= 0 and det.max() < og.n_det and len(dat) <= 4 * og.n_vox
        S = sparse.coo_matrix((wts.astype(np.float64), (det, dat)), shape=(og.n_det, og.n_vox)).tocsr()
        d_ref, _ = O.voxel_forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], rec)
        np.testing.assert_allclose(S @ rec, d_ref, rtol=0, atol=1e-6)


def test_orphan_forward_project_semantics():
    """forward_project (src/forward_projection.f90), float32: NINT sample count and no cor_shift.  Against the live path
    (float64) it must agree to float32 accuracy when cor_shift = 0 and the sample counts coincide, and it must NOT change when
    the geometry carries a centre-of-rotation shift (the argument is never read, forward_projection.f90:1,10)."""
    g, og = make_geoms((12, 12, 12), (12, 12), 3)
    phi, alpha, beta, xyz = random_poses(3, 2)
    rec = np.random.default_rng(3).random((12, 12, 12)).astype(np.float32)
    orphan = O.forward_project_orphan(og, rec, alpha, beta, phi, xyz)
    live = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).forward(rec)
    assert orphan.shape == live.shape and orphan.dtype == np.float32
    # the live path marches int(r_length/step) = 24 or 23 samples, the orphan NINT(.) = 24: the 24th sample lies at y = +sy - 1,
    # outside the 12-voxel-deep volume, so both see the same terms here
    assert np.linalg.norm(orphan - live) / np.linalg.norm(live) < 5e-6
    _, og2 = make_geoms((12, 12, 12), (12, 12), 3, cor=[0.8, 0.0, 0.0])
    assert np.array_equal(O.forward_project_orphan(og2, rec, alpha, beta, phi, xyz), orphan)


def test_voxel_back_projector_is_the_transposed_splat_matrix():
    """Two independently restated Fortran routines must agree: back_project / voxel_back_bilinear
    (src/external_back_projection.f90:30-68, a per-voxel gather of det_image(x', z')) with origin = vox_origin - cor_shift is
    the transpose of the matrix bilinear_sparse emits (src/vox_wt_grad.f90:58-112, det index fx + ndim_x * fz) -- same transform
    Ry(Rx Rz x + t), same four taps, same per-tap bounds checks -- once the detector image is transposed."""
    from scipy import sparse
    g, og = make_geoms((9, 8, 10), (11, 9), 3, cor=[0.3, 0.0, -0.2])
    phi, alpha, beta, xyz = random_poses(3, 14, tilt=0.1, shift=2.5)
    rng = np.random.default_rng(2)
    y = rng.random((3, 11, 9))                                   # [view][x'][z'] as back_project reads it
    for i in range(3):
        orig = og.vox_origin - og.cor_shift[i]
        back = O.voxel_back_project(og, y[i:i + 1], alpha[i:i + 1], beta[i:i + 1], phi[i:i + 1], xyz[i:i + 1], origin=orig)
        dat, det, wts = O.voxel_forward_sparse(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i])
        S = sparse.coo_matrix((wts.astype(np.float64), (det, dat)), shape=(og.n_det, og.n_vox)).tocsr()
        np.testing.assert_allclose(back, S.T @ y[i].T.ravel(), rtol=0, atol=2e-6)      # weights: float64 products vs float32 products
