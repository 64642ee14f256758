"""The reference's OWN callers, imported unchanged from /root/reference, running on the drop-in operators
(BASELINE.json north_star: "recon/sirt.py, cgls.py, regularized.py ... work unchanged").

    utilities.projection_operators  -> tomography_alignment_b200.projection_operators   (aliased in sys.modules)
    utilities.geometry              -> tomography_alignment_b200.geometry
    scipy.optimize.linesearch       -> scipy.optimize._linesearch   (module renamed by scipy >= 1.8)
    utilities.linear_operators      -> empty stub (recon/cgls.py:3 imports a module the reference does not ship)

recon/sirt.py, recon/cgls.py, recon/regularized.py and utilities/alignment_functions.py are then imported from the
reference tree as they are and run against (a) the drop-in operator and (b) the reference's real scipy CSR matrix
rebuilt by the oracle; both must give the same iterates.  CPU tier: the reference's constructor call has no backend
argument, so the shim subclass hands ProjectionMatrix the oracle-backed test backend in place of the GPU (what is tested is
that the reference code runs unmodified on the operator API).  Skipped where /root/reference does not exist (GPU box)."""
import importlib
import os
import sys
import types

import numpy as np
import pytest

from oracle import oracle as O
from helpers import OracleBackend, make_geoms, random_poses, rel_l2

REF = "/root/reference"
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "recon")), reason="reference tree not present")


@pytest.fixture()
def ref(monkeypatch):
    """sys.modules aliases + the reference's caller modules imported from its own files."""
    import scipy.optimize
    import scipy.optimize._linesearch as ls
    import tomography_alignment_b200.geometry as our_geometry
    import tomography_alignment_b200.projection_operators as our_po

    shim = types.ModuleType("utilities.projection_operators")
    shim.__dict__.update({k: v for k, v in our_po.__dict__.items() if not k.startswith("__")})

    class ProjectionMatrix(our_po.ProjectionMatrix):
        def __init__(self, geometry, precision=np.float32):       # the reference's call signature
            super().__init__(geometry, precision=precision, backend=OracleBackend(geometry))
    shim.ProjectionMatrix = ProjectionMatrix

    utilities = types.ModuleType("utilities")
    utilities.__path__ = [os.path.join(REF, "utilities")]          # alignment_functions, tv_denoise: the reference's files
    utilities.projection_operators = shim
    utilities.geometry = our_geometry
    recon = types.ModuleType("recon")
    recon.__path__ = [os.path.join(REF, "recon")]
    stub = types.ModuleType("utilities.linear_operators")
    utilities.linear_operators = stub
    for name, mod in (("utilities", utilities), ("utilities.projection_operators", shim), ("utilities.geometry", our_geometry),
                      ("utilities.linear_operators", stub), ("recon", recon), ("scipy.optimize.linesearch", ls)):
        monkeypatch.setitem(sys.modules, name, mod)
    monkeypatch.setattr(scipy.optimize, "linesearch", ls, raising=False)
    mods = types.SimpleNamespace(ProjectionMatrix=ProjectionMatrix, Geometry=our_geometry.Geometry)
    try:
        mods.sirt = importlib.import_module("recon.sirt")
        mods.cgls = importlib.import_module("recon.cgls")
        mods.regularized = importlib.import_module("recon.regularized")
        mods.alignment_functions = importlib.import_module("utilities.alignment_functions")
        for m in (mods.sirt, mods.cgls, mods.regularized, mods.alignment_functions):
            assert m.__file__.startswith(REF), m.__file__
        yield mods
    finally:
        for name in list(sys.modules):
            if name.split(".")[0] in ("recon", "utilities"):
                sys.modules.pop(name, None)


def _problem(n=16, n_proj=8, seed=0):
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = random_poses(n_proj, seed)
    angles = np.array([phi, alpha, beta]).T
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    rng = np.random.default_rng(seed + 1)
    x = np.zeros((n, n, n), np.float32)
    x[3:n - 3, 4:n - 4, 2:n - 2] = rng.random((n - 6, n - 8, n - 4)).astype(np.float32) + 0.5
    b = op.forward(x).astype(np.float32)                      # (n_proj, n_det)
    return g, og, angles, xyz, op, x, b


def _with_real_csr(cls, solver, csr):
    """A second instance of the reference's solver class whose proj_mat is the reference's real CSR matrix."""
    twin = cls.__new__(cls)
    twin.__dict__.update({k: (v.copy() if isinstance(v, np.ndarray) else v) for k, v in solver.__dict__.items()})
    twin.f_proj_obj, twin.proj_mat = object(), csr
    twin._initialize()
    return twin


def test_reference_sirt_runs_unchanged(ref):
    g, og, angles, xyz, op, x, b = _problem()
    s = ref.sirt.SIRT(g, b.copy(), angles, xyz, options={"precision": np.float32})          # recon/sirt.py:9-40
    assert type(s.f_proj_obj) is ref.ProjectionMatrix and s.proj_mat.shape == (8 * g.n_det, g.n_vox)
    t = _with_real_csr(ref.sirt.SIRT, s, op.csr(np.float32))
    assert rel_l2(s.W, t.W) < 1e-6 and rel_l2(s.V, t.V) < 1e-6
    rec_s, err_s = s.run_main_iteration(niter=5, positivity=True)                           # recon/sirt.py:42-107
    rec_t, err_t = t.run_main_iteration(niter=5, positivity=True)
    assert rec_s.shape == (16, 16, 16) and len(err_s) == len(err_t) == 5
    assert rel_l2(rec_s, rec_t) < 1e-5 and np.allclose(err_s, err_t, rtol=1e-5)
    assert err_s[-1] < err_s[0] and rel_l2(rec_s, x) < 0.5


def test_reference_cgls_runs_unchanged(ref):
    g, og, angles, xyz, op, x, b = _problem(seed=3)
    s = ref.cgls.CGLS(g, b.copy(), angles, xyz)                                             # recon/cgls.py:9-36
    s.method = "matrix"      # recon/cgls.py:51 reads self.method, which upstream never sets (an upstream bug, not the operator's)
    t = _with_real_csr(ref.cgls.CGLS, s, op.csr(np.float32))
    rec_s, err_s = s.run_main_iteration(niter=5)[:2]
    rec_t, err_t = t.run_main_iteration(niter=5)[:2]
    assert rel_l2(rec_s, rec_t) < 2e-5 and np.allclose(err_s, err_t, rtol=1e-4)


def test_reference_regularized_runs_unchanged(ref):
    g, og, angles, xyz, op, x, b = _problem(seed=5)
    csr = op.csr(np.float32)
    for run in (lambda r: r.run_fista(niter=3, hyper=1.e3, beta_tv=0.1, niter_tv=5),        # recon/regularized.py:57-154
                lambda r: r.run_lasso_ista(niter=3, reg_param=0.01),                        # :239-315
                lambda r: r.run_lasso_accelerated(niter=3, reg_param=0.01)):                # :334-413
        # (run_tikhonov_gd, :156-237, cannot run upstream: :177 needs 2-D projections, :190 -> my_tikh_f :418 needs 1-D ones)
        s = ref.regularized.RegularizedRecon(g, b.copy(), angles, xyz)
        s.my_rank = 0                            # :387 reads an attribute only the _mpi twin sets (upstream slip)
        t = _with_real_csr(ref.regularized.RegularizedRecon, s, csr)
        def go(r):
            try:
                return run(r)
            except UnboundLocalError as e:       # run_lasso_ista ends with an unconditional plt.figure() (:312); the iterations are done
                assert "plt" in str(e)
                return None
        out_s, out_t = go(s), go(t)
        rec_s = out_s[0] if isinstance(out_s, tuple) else s.rec
        rec_t = out_t[0] if isinstance(out_t, tuple) else t.rec
        assert np.linalg.norm(rec_s) > 0 and rel_l2(rec_s, rec_t) < 1e-4


def test_reference_alignment_functions_run_unchanged(ref):
    """utilities/alignment_functions.py:7-37,151-188 (AlignmentUtilities, cost_xzab / gradient_xzab) on the drop-in
    ProjectionMatrix.projection_gradient, against the oracle's projection + gradient image."""
    af = ref.alignment_functions
    g, og, angles, xyz, op, x, b = _problem(n=12, n_proj=3, seed=7)
    pm = ref.ProjectionMatrix(g, precision=np.float32)
    i = 1
    g1 = ref.Geometry(1, g.vox_shape, g.vox_pix, g.det_shape, g.det_pix, cor_shift=np.zeros(3))
    g1.cor_shift = g1.cor_shift[0]                       # examples/align_rigid.py hands a single (3,) row
    au = af.AlignmentUtilities(b[i], pm, g1)
    rec = (0.9 * x).astype(np.float32)
    params = np.array([0.3, -0.2, 0.004, -0.003])        # dx, dz, dalpha, dbeta
    c = af.cost_xzab(params, au, rec, angles[i], xyz[i])
    gr = af.gradient_xzab(params, au, rec, angles[i], xyz[i])
    t = np.array([xyz[i][0] + params[0], xyz[i][1], xyz[i][2] + params[1]])
    p_ref, d_ref = O.projection_gradient(og, rec, angles[i][1] + params[2], angles[i][2] + params[3], angles[i][0], t, np.zeros(3))
    res = b[i].ravel() - p_ref
    assert abs(c - 0.5 * res @ res) <= 1e-5 * c
    assert rel_l2(gr, (-d_ref[[0, 2, 4, 5]]) @ res) < 1e-4
    # the reference's own gradient_descent driver (alignment_functions.py:40-110) runs on top of them
    sol = af.gradient_descent(np.zeros(4), af.cost_xzab, af.gradient_xzab, args=(au, rec, angles[i], xyz[i], np.ones(4)),
                              options={"maxiter": 2})
    assert np.all(np.isfinite(np.asarray(sol[0] if isinstance(sol, tuple) else sol, dtype=float)))


def test_reference_align_cc_consumes_the_operator_output(ref, monkeypatch):
    """align/align_cc.py (BASELINE.json north_star lists it among the callers that must work unchanged) never touches the operators:
    it cross-correlates consecutive projections.  Imported from the reference tree as it is (skimage, which the image lacks, is
    stubbed: only cor_flipping / cross_correlation_skimage need it) and fed the (n_proj, nx, nz) projections the drop-in operator
    produces: cross_correlation_numpy must recover integer detector shifts applied through xyz_shift."""
    sk = types.ModuleType("skimage")
    skr = types.ModuleType("skimage.registration")
    skr.phase_cross_correlation = lambda *a, **k: (_ for _ in ()).throw(RuntimeError("skimage is not installed"))
    sk.registration = skr
    monkeypatch.setitem(sys.modules, "skimage", sk)
    monkeypatch.setitem(sys.modules, "skimage.registration", skr)
    align = types.ModuleType("align")
    align.__path__ = [os.path.join(REF, "align")]
    monkeypatch.setitem(sys.modules, "align", align)
    try:
        cc = importlib.import_module("align.align_cc")
        assert cc.__file__.startswith(REF)
        n, n_proj = 16, 4
        g, og = make_geoms((n, n, n), (n, n), n_proj)
        x = np.zeros((n, n, n), np.float32)
        x[5:11, 6:10, 4:12] = np.random.default_rng(0).random((6, 4, 8)).astype(np.float32) + 0.5
        shifts = np.array([[0, 0, 0], [2, 0, 0], [2, 0, 0], [-1, 0, 0]], dtype=np.float64)        # whole detector pixels along x
        pm = ref.ProjectionMatrix(g, precision=np.float32)
        A = pm.projection_matrix(phi=np.zeros(n_proj), xyz_shift=shifts)                            # same angle, shifted object
        from scipy import sparse
        proj = sparse.csr_matrix.dot(A, x.ravel()).reshape(n_proj, n, n)
        offsets, aligned = cc.cross_correlation_numpy(proj)
        assert aligned.shape == proj.shape
        # every projection is rolled back onto the first one: the step-to-step offsets are 2, 0, -3 pixels along x
        assert offsets[:, 0].tolist() == [0.0, 2.0, 2.0, -1.0] and np.all(offsets[:, 1] == 0.0)
        for i in range(1, n_proj):
            assert np.abs(aligned[i] - aligned[0]).max() < 1e-3 * np.abs(proj[0]).max()
    finally:
        sys.modules.pop("align.align_cc", None)
