"""Host logic of the device-resident solvers (recon.py) and of the alignment layer (alignment.py), run on the CPU
with the emulated kernel cores (tests/emu) standing in for the GPU."""
import numpy as np
import pytest
from scipy import sparse

from oracle import oracle as O
from tomography_alignment_b200 import ProjectionMatrix, pose_table
from tomography_alignment_b200 import alignment as AL
from tomography_alignment_b200.recon import CGLS, SIRT
from helpers import EmuBackend, OracleBackend, make_geoms, random_poses, rel_l2


def _problem(n=10, n_proj=8, seed=3):
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = random_poses(n_proj, seed, shift=0.8)
    R = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr(np.float64)
    c = np.arange(n) - (n - 1) / 2
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    truth = (np.exp(-(X ** 2 + Y ** 2 + Z ** 2) / (0.08 * n * n)) * (np.abs(X) < 0.35 * n)).astype(np.float32)
    b = (R @ truth.ravel()).reshape(n_proj, -1).astype(np.float32)
    angles = np.array([phi, alpha, beta]).T
    return g, og, R, truth, b, angles, xyz


def _sirt_reference(R, b, n_proj, n_vox, niter, positivity, truth):
    """recon/sirt.py:26-78 written against a scipy CSR matrix (float64)."""
    W = R @ np.ones(n_vox)
    V = R.T @ np.ones(R.shape[0])
    V[V == 0.] = np.inf
    W[W == 0.] = np.inf
    V, W = 1. / V, 1. / W
    rec = np.zeros(n_vox)
    err = np.zeros(niter)
    k, stop = 0, 0
    nf = np.linalg.norm(truth)
    while k < niter and not stop:
        res = b - (R @ rec).reshape(n_proj, -1)
        rec += V * (R.T @ (W * res.ravel()))
        if positivity:
            rec[rec < 0.] = 0.
        err[k] = np.linalg.norm(truth - rec) / nf
        if k > 0 and err[k] > err[k - 1]:
            stop = 1
        k += 1
    return rec, err[:k]


def test_device_sirt_matches_reference_iteration():
    g, og, R, truth, b, angles, xyz = _problem()
    s = SIRT(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
    rec, err = s.run_main_iteration(niter=6, positivity=True)
    ref, ref_err = _sirt_reference(R, b.astype(np.float64), 8, g.n_vox, 6, True, truth.ravel().astype(np.float64))
    assert rec.shape == tuple(g.vox_shape) and rec.dtype == np.float32
    assert len(err) == len(ref_err)
    assert rel_l2(rec, ref) < 2e-5
    np.testing.assert_allclose(err, ref_err, rtol=1e-4)
    assert err[-1] < err[0]
    # warm start through options['rec'] (recon/sirt.py:17) continues the iteration
    s2 = SIRT(g, b, angles, xyz, options={"ground_truth": truth, "rec": rec.ravel()}, backend=EmuBackend(g))
    _, err2 = s2.run_main_iteration(niter=2)
    assert err2[0] < err[-1]


def test_device_cgls_converges_and_matches_normal_equations():
    g, og, R, truth, b, angles, xyz = _problem()
    c = CGLS(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
    rec, err = c.run_main_iteration(niter=12)
    assert err[-1] < 0.5 * err[0]
    # CGLS iterates minimise ||b - A x|| over Krylov spaces: the residual norm decreases monotonically
    r = [np.linalg.norm(b.ravel() - R @ rec.ravel())]
    c2 = CGLS(g, b, angles, xyz, backend=EmuBackend(g))
    rec2, err2 = c2.run_main_iteration(niter=3)
    assert np.linalg.norm(b.ravel() - R @ rec2.ravel()) > r[0]
    assert np.all(np.diff(err2) < 0)


@pytest.mark.parametrize("mode", AL.MODES)
def test_alignment_closures_follow_reference_conventions(mode):
    """gradient_<mode>(p) = (-dproj[rows]) . (b - proj) and cost_<mode>(p) = 0.5 ||b - proj||^2 at the pose obtained
    by ADDING p to (xyz_in, angles_in) (alignment_functions.py:113-485)."""
    g, og = make_geoms((9, 9, 9), (9, 9), 1)
    rec = np.random.default_rng(1).random((9, 9, 9))
    meas = np.random.default_rng(2).random(81)
    pm = ProjectionMatrix(g, precision=np.float64, backend=OracleBackend(g))
    geo1 = make_geoms((9, 9, 9), (9, 9), 1)[0]
    geo1.cor_shift = g.cor_shift[0]
    ao = AL.AlignmentUtilities(meas, pm, geo1)
    angles_in, xyz_in = np.array([0.6, 0.004, -0.006]), np.array([0.2, 0.1, -0.3])
    p = np.linspace(0.01, 0.03, len(mode)) * np.array([10.0 if c in "xz" else 1.0 for c in mode])
    cost = getattr(AL, "cost_" + mode)(p, ao, rec, angles_in, xyz_in)
    grad = getattr(AL, "gradient_" + mode)(p, ao, rec, angles_in, xyz_in)
    ang, tr = AL.apply_parameters(mode, p, angles_in, xyz_in)
    pr, gr = O.forward_proj_grad(og, ang[1], ang[2], ang[0], tr, np.zeros(3), rec)
    res = meas - pr
    assert abs(cost - 0.5 * res @ res) < 1e-5 * (0.5 * res @ res)
    rows = [AL._ROW[c] for c in mode]
    assert rel_l2(grad, (-gr[rows]) @ res) < 1e-5
    assert np.array_equal(np.where(AL.vary_mask(mode))[0], sorted(rows))
    # return_vector variants
    assert getattr(AL, "cost_" + mode)(p, ao, rec, angles_in, xyz_in, return_vector=True).shape == (81,)
    assert getattr(AL, "gradient_" + mode)(p, ao, rec, angles_in, xyz_in, return_vector=True).shape == (81, len(mode))


def test_batched_alignment_equals_per_view_closures_and_recovers_shifts():
    n, n_proj = 14, 5
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    c = np.arange(n) - (n - 1) / 2
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    rec = np.exp(-((X - 1) ** 2 + 1.4 * Y ** 2 + 0.7 * (Z + 1) ** 2) / 9.0)
    phi = np.array([0.2, 0.9, 1.5, 2.1, 2.8])
    rng = np.random.default_rng(4)
    true_shift = np.zeros((n_proj, 3))
    true_shift[:, 0] = rng.uniform(-0.6, 0.6, n_proj)
    true_shift[:, 2] = rng.uniform(-0.6, 0.6, n_proj)
    angles = np.array([phi, np.zeros(n_proj), np.zeros(n_proj)]).T
    meas = O.OracleOperator(og, phi=phi, xyz_shift=true_shift).forward(rec).astype(np.float32)
    ba = AL.BatchedAlignment(g, meas, angles, np.zeros((n_proj, 3)), mode="xz", backend=EmuBackend(g))
    x0 = rng.uniform(-0.2, 0.2, (n_proj, 2))
    f, gr = ba.cost_and_gradient(rec, x0)
    pm = ProjectionMatrix(g, precision=np.float64, backend=OracleBackend(g))
    for i in range(n_proj):
        gi = make_geoms((n, n, n), (n, n), 1)[0]
        gi.cor_shift = g.cor_shift[i]
        ao = AL.AlignmentUtilities(meas[i], pm, gi)
        assert abs(f[i] - AL.cost_xz(x0[i], ao, rec, angles[i], np.zeros(3))) <= 1e-4 * max(f[i], 1e-3)
        assert rel_l2(gr[i], AL.gradient_xz(x0[i], ao, rec, angles[i], np.zeros(3))) < 1e-3
    x, fin, it = ba.minimize(rec, bounds=((-3., 3.), (-3., 3.)), maxiter=25)
    assert np.abs(x - true_shift[:, [0, 2]]).max() < 0.05, (x, true_shift[:, [0, 2]])
    assert (fin < 1e-2 * f).all() or fin.max() < 1e-4


# ---------------- recon/regularized.py: Tikhonov and Lasso loops ----------------
def _soft(x, lam):
    return np.sign(x) * np.maximum(np.abs(x) - lam, 0.0)


def _tikhonov_reference(R, b, n_vox, niter, lam, positivity, truth):
    """recon/regularized.py:156-237 against a scipy CSR matrix, with scipy's own Armijo search (flat residual)."""
    from scipy.optimize._linesearch import line_search_armijo
    f = lambda x: 0.5 * np.linalg.norm(R @ x - b) ** 2 + 0.5 * lam * np.linalg.norm(x) ** 2
    rec, err, k, stop = np.zeros(n_vox), np.zeros(niter), 0, 0
    nf = np.linalg.norm(truth)
    while k < niter and not stop:
        res = b - R @ rec
        grad = -(R.T @ res) + lam * rec
        cost = 0.5 * (np.linalg.norm(res) ** 2 + lam * np.linalg.norm(rec) ** 2)
        alpha, _, _ = line_search_armijo(f, rec, -grad, grad, cost, alpha0=1.0)
        assert alpha is not None
        rec = rec - alpha * grad
        if positivity:
            rec[rec < 0.] = 0.
        err[k] = np.linalg.norm(truth - rec) / nf
        if k > 1 and err[k] > err[k - 1]:
            stop = 1
        k += 1
    return rec, err[:k]


def _lasso_reference(R, b, n_vox, niter, lam, alpha0, beta, truth, accelerated):
    """recon/regularized.py:239-413 against a scipy CSR matrix."""
    rec, err, k, stop = np.zeros(n_vox), np.zeros(niter), 0, 0
    x_0, x_1 = np.zeros(n_vox), np.zeros(n_vox)
    nf = np.linalg.norm(truth)
    steps = []
    while k < niter and not stop:
        res = R @ rec - b
        grad = R.T @ res
        t, g0 = alpha0, 0.5 * np.linalg.norm(res) ** 2
        while t > 1e-16:
            xp = _soft(rec - t * grad, t * lam)
            Gt = rec - xp
            if 0.5 * np.linalg.norm(R @ xp - b) ** 2 <= g0 - grad @ Gt + 0.5 / t * np.linalg.norm(Gt) ** 2:
                break
            t *= beta
        steps.append(t)
        if accelerated:
            v = x_1 + (k - 2) / (k + 1) * (x_1 - x_0)
            rec = _soft(v - t * grad, t * lam)
            x_0, x_1 = x_1, rec.copy()
        else:
            rec = _soft(rec - t * grad, t * lam)
        err[k] = np.linalg.norm(truth - rec) / nf
        if k > 1 and err[k] > err[k - 1]:
            stop = 1
        k += 1
    return rec, err[:k], np.array(steps)


def test_scalar_search_armijo_equals_scipy():
    from scipy.optimize._linesearch import scalar_search_armijo as ref_search
    from tomography_alignment_b200.recon import scalar_search_armijo
    for coeffs in ([1.0, -2.0, 30.0], [0.3, -1.0, 400.0], [2.0, -0.5, 0.1], [1.0, -1.0, 1e4, -3e3]):
        phi = lambda a, c=coeffs: sum(ck * a ** i for i, ck in enumerate(c))
        got = scalar_search_armijo(phi, coeffs[0], coeffs[1])
        want = ref_search(phi, coeffs[0], coeffs[1])
        assert got[0] == pytest.approx(want[0], rel=1e-14) and got[1] == pytest.approx(want[1], rel=1e-14)


def test_device_tikhonov_matches_reference_iteration():
    from tomography_alignment_b200.recon import RegularizedRecon
    g, og, R, truth, b, angles, xyz = _problem()
    s = RegularizedRecon(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
    rec, err = s.run_tikhonov_gd(niter=6, reg_param=0.5, positivity=True)
    ref, ref_err = _tikhonov_reference(R, b.ravel().astype(np.float64), g.n_vox, 6, 0.5, True,
                                       truth.ravel().astype(np.float64))
    assert rec.shape == (g.n_vox,) and rec.dtype == np.float32          # the reference returns the flat volume here
    assert len(err) == len(ref_err)
    assert rel_l2(rec, ref) < 5e-5
    np.testing.assert_allclose(err, ref_err, rtol=2e-4)
    assert err[-1] < err[0]


@pytest.mark.parametrize("accelerated", [False, True])
def test_device_lasso_matches_reference_iteration(accelerated):
    from tomography_alignment_b200.recon import RegularizedRecon, soft_thresholding
    import torch
    x = torch.tensor([-2.0, -0.5, 0.0, 0.4, 3.0])
    np.testing.assert_allclose(soft_thresholding(x, 0.5).numpy(), [-1.5, 0.0, 0.0, 0.0, 2.5])
    g, og, R, truth, b, angles, xyz = _problem()
    s = RegularizedRecon(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
    run = s.run_lasso_accelerated if accelerated else s.run_lasso_ista
    rec, err = run(niter=5, reg_param=0.05, alpha0=1.0, beta=0.5)
    ref, ref_err, steps = _lasso_reference(R, b.ravel().astype(np.float64), g.n_vox, 5, 0.05, 1.0, 0.5,
                                           truth.ravel().astype(np.float64), accelerated)
    assert rec.shape == ((g.n_vox,) if accelerated else tuple(g.vox_shape))
    assert len(err) == len(ref_err)
    assert rel_l2(rec.ravel(), ref) < 5e-5
    np.testing.assert_allclose(err, ref_err, rtol=2e-4)
    if not accelerated:
        np.testing.assert_allclose(s.step_size[:len(steps)], steps)
