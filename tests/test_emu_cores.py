"""The __host__ __device__ cores of the CUDA kernels (csrc/ray_core.h, back_core.h), executed on the CPU by
tests/emu, against the oracle.  Tolerances are BASELINE.json's: relative L2 <= 1e-5 for projections and
backprojections, <= 1e-4 for the alignment gradients."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tomography_alignment_b200 import pose_table
from helpers import EmuBackend, make_geoms, random_poses, rel_l2

TOL_PROJ, TOL_GRAD = 1e-5, 1e-4

CASES = [
    # shape, dshape, n_proj, kwargs
    ((16, 16, 16), (16, 16), 6, dict()),
    ((24, 20, 18), (24, 18), 5, dict(cor=[0.7, 0, 0])),
    ((16, 16, 16), (20, 12), 5, dict(tilt=0.2, shift=5.0)),            # detector != volume, big jitter
    ((16, 16, 16), (16, 16), 4, dict(step=0.5)),
    ((16, 16, 16), (16, 16), 4, dict(step=1.7)),                       # |D| > 1: integer part in the address step
    ((12, 12, 12), (12, 12), 3, dict(shift=14.0, phis=[0.2, 1.1, 2.0])),   # rays mostly / entirely miss the volume
    ((5, 40, 3), (5, 3), 3, dict()),                                   # ragged: tiny x/z, long y
    ((33, 9, 35), (33, 35), 2, dict(phis=[0.4, 2.9])),                 # not a multiple of the 32x8 tile
    ((20, 56, 60), (20, 60), 4, dict(phis=[0.3, 1.9, 3.5, 5.2])),      # not a cube, but laid out with the pitches of 64^3 (tomo_pad_pitch)
]


def setup(shape, dshape, n_proj, cor=None, step=1.0, tilt=0.02, shift=2.0, phis=None, seed=0):
    g, og = make_geoms(shape, dshape, n_proj, cor=cor, step=step)
    phi, alpha, beta, xyz = random_poses(n_proj, seed, tilt=tilt, shift=shift, phis=phis)
    be = EmuBackend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    return g, og, be, op, (phi, alpha, beta, xyz)


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES)
def test_forward_and_adjoint(shape, dshape, n_proj, kw):
    g, og, be, op, _ = setup(shape, dshape, n_proj, **kw)
    rng = np.random.default_rng(1)
    vol = rng.random(shape).astype(np.float32)
    ref = op.forward(vol)
    got = be.forward(vol).numpy().reshape(n_proj, -1)
    if np.linalg.norm(ref) > 0:
        assert rel_l2(got, ref) <= TOL_PROJ
    else:
        assert np.abs(got).max() == 0.0
    y = rng.random((n_proj, og.n_det)).astype(np.float32)
    refb = op.adjoint(y)
    gotb = be.adjoint(y).numpy().ravel()
    if np.linalg.norm(refb) > 0:
        assert rel_l2(gotb, refb) <= TOL_PROJ
    else:
        assert np.abs(gotb).max() == 0.0


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES)
def test_projection_gradient(shape, dshape, n_proj, kw):
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    rng = np.random.default_rng(2)
    vol = rng.random(shape).astype(np.float32)
    meas = (op.forward(vol) * 1.02 + 0.05).astype(np.float32)
    out = be.proj_grad(vol, meas=meas)
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][i].numpy(), gr) <= TOL_GRAD
        res = meas[i].astype(np.float64) - p
        assert rel_l2(out["grad6"][i].numpy(), -gr @ res) <= TOL_GRAD
        assert abs(out["cost"][i].item() - 0.5 * res @ res) <= 1e-5 * (0.5 * res @ res)


def test_compile_time_stride_variants_all_sign_octants():
    """64^3 is one of the cubes whose padded strides are template constants in ray_core.h (ray_march_fixed): eight variants, one
    per sign octant of the step D.  phi picks the signs of D_x, D_y, the tilt that of D_z; both marches against the oracle."""
    n = 64
    phis = np.array([0.3, 0.3, 1.9, 1.9, 3.5, 3.5, 5.2, 5.2])
    alpha = np.array([0.03, -0.03] * 4)
    beta = np.array([-0.02, 0.025] * 4)
    xyz = np.random.default_rng(5).uniform(-2, 2, (8, 3))
    g, og = make_geoms((n, n, n), (n, n), 8)
    be = EmuBackend(g)
    be.set_poses(pose_table(np.array([phis, alpha, beta]).T, xyz, g.cor_shift))
    vol = np.random.default_rng(6).random((n, n, n)).astype(np.float32)
    out = be.proj_grad(vol)
    octants = set()
    for i in range(8):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phis[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][i].numpy(), gr) <= TOL_GRAD
        vs = O.ViewSetup(og, alpha[i], beta[i], phis[i], xyz[i], og.cor_shift[i])
        d = vs.r_hat[:, 0]
        octants.add(tuple(np.sign(d).astype(int)))
    assert len(octants) == 8, octants
    fwd = be.forward(vol).numpy().reshape(8, -1)
    assert rel_l2(fwd, out["proj"].numpy().reshape(8, -1)) <= 2e-6          # float32 march vs fixed-point march


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES[:3])
def test_voxel_driven_bilinear_backprojector(shape, dshape, n_proj, kw):
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    y = np.random.default_rng(3).random((n_proj,) + tuple(dshape)).astype(np.float32)
    ref = O.voxel_back_project(og, y, alpha, beta, phi, xyz)
    assert rel_l2(be.voxel_back(y).numpy(), ref) <= TOL_PROJ


@pytest.mark.parametrize("shape,dshape,kw", [((16, 16, 16), (16, 16), dict()), ((14, 20, 37), (14, 37), dict(cor=[0.4, 0, 0])),
                                              ((12, 12, 12), (18, 9), dict(shift=6.0)), ((16, 16, 16), (16, 16), dict(step=0.5)),
                                              ((20, 56, 60), (20, 60), dict())])       # pitches of 64^3, 64 planes to visit
def test_separable_forward_for_untilted_views(shape, dshape, kw):
    """alpha = beta = 0 (the API's default poses): the separable cores of sep_core.h reproduce the oracle."""
    n_proj = 7
    g, og, be, op, _ = setup(shape, dshape, n_proj, tilt=0.0, phis=[0.0, 0.4, np.pi / 4, np.pi / 2, 2.0, 2.9, np.pi], **kw)
    assert np.all(be.views[:, 146] == 1.0)                 # V_SEP set by tomo_views_compute_host
    vol = np.random.default_rng(8).random(shape).astype(np.float32)
    assert rel_l2(be.forward(vol).numpy().reshape(n_proj, -1), op.forward(vol)) <= TOL_PROJ
    y = np.random.default_rng(9).random((n_proj, og.n_det)).astype(np.float32)
    assert rel_l2(be.adjoint(y).numpy(), op.adjoint(y)) <= TOL_PROJ             # separable adjoint cores
    phi, alpha, beta, xyz = _[0], _[1], _[2], _[3]
    out = be.proj_grad(vol, meas=y)                                              # separable gradient cores
    for i in (1, 2, 4, 5):          # phi = 0, pi/2, pi sit exactly on lattice planes (one-sided derivative: DESIGN.md section 5)
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][i].numpy(), gr) <= TOL_GRAD, i
        res = y[i].astype(np.float64) - p
        assert rel_l2(out["grad6"][i].numpy(), -gr @ res) <= TOL_GRAD
    # a tilted table is not flagged
    g2, og2, be2, op2, _ = setup(shape, dshape, 3, tilt=0.01)
    assert np.all(be2.views[:, 146] == 0.0)


@pytest.mark.parametrize("det_pix,tilt", [((1.5, 0.75), 0.02), ((0.8, 1.5), 0.02), ((1.5, 0.75), 0.0), ((0.8, 1.5), 0.0)])
def test_detector_pitch_differs_from_voxel_size(det_pix, tilt):
    """Detector pixels that are not voxel sized: U and W are scaled, adjacent detector rows can share a z cell
    (W_z = 0.75) or skip one (W_z = 1.5); tilted (generic cores) and untilted (separable cores)."""
    shape, dshape, n_proj = (14, 16, 18), (12, 20), 5
    g, og = make_geoms(shape, dshape, n_proj, det_pix=det_pix)
    phi, alpha, beta, xyz = random_poses(n_proj, 13, tilt=tilt, shift=1.5)
    be = EmuBackend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    rng = np.random.default_rng(14)
    vol = rng.random(shape).astype(np.float32)
    y = rng.random((n_proj, og.n_det)).astype(np.float32)
    assert rel_l2(be.forward(vol).numpy().reshape(n_proj, -1), op.forward(vol)) <= TOL_PROJ
    assert rel_l2(be.adjoint(y).numpy(), op.adjoint(y)) <= TOL_PROJ
    out = be.proj_grad(vol)
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["dproj"][i].numpy(), gr) <= TOL_GRAD


def test_exact_lattice_pose_phi0():
    """phi = alpha = beta = 0, t = 0: every sample sits on a lattice plane in x and z.  Projection = sum over y,
    and the one-sided derivatives agree with the reference because the floor of an exact integer is exact."""
    g, og, be, op, _ = setup((12, 10, 9), (12, 9), 1, tilt=0.0, shift=0.0, phis=[0.0])
    vol = np.random.default_rng(4).random((12, 10, 9)).astype(np.float32)
    out = be.proj_grad(vol)
    np.testing.assert_allclose(out["proj"][0].numpy(), vol.astype(np.float64).sum(axis=1), rtol=2e-6)
    p, gr = O.forward_proj_grad(og, 0.0, 0.0, 0.0, np.zeros(3), np.zeros(3), vol)
    for k in (0, 2, 3, 4, 5):      # row 1 (ty) is identically ~0
        assert rel_l2(out["dproj"][0, k].numpy(), gr[k]) <= TOL_GRAD


def test_against_reference_numpy_golden():
    """Directly against the reference-authored fixtures (zero-shell volumes, see make_golden.py)."""
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_numpy_cases.npz"))
    for name in gold["case_names"]:
        q = lambda k: gold["%s/%s" % (name, k)]
        alpha, beta, phi = (float(v) for v in q("pose"))
        g, _ = make_geoms(tuple(q("vox_shape")), tuple(q("det_shape")), 1, cor=q("cor"), step=float(q("step")))
        be = EmuBackend(g)
        be.set_poses(pose_table(np.array([[phi, alpha, beta]]), q("xyz")[None, :], q("cor")[None, :]))
        out = be.proj_grad(q("rec").astype(np.float32))
        assert rel_l2(out["proj"][0].numpy(), q("proj")) <= TOL_PROJ, name
        assert rel_l2(out["dproj"][0].numpy(), q("grad")) <= TOL_GRAD, name
        shape = tuple(q("vox_shape"))
        at = be.adjoint(q("y")[None, :].astype(np.float32)).numpy().reshape(shape)[1:-1, 1:-1, 1:-1]
        assert rel_l2(at, q("At_dot_y").reshape(shape)[1:-1, 1:-1, 1:-1]) <= TOL_PROJ, name


def test_adjoint_pair_at_64cubed():
    """<A x, y> = <x, A^T y> for the emulated kernel pair at the reference's example size (64^3, 4 views)."""
    g, og, be, op, _ = setup((64, 64, 64), (64, 64), 4, seed=9, phis=[0.1, 0.9, 1.7, 2.8])
    rng = np.random.default_rng(10)
    x = rng.random((64, 64, 64)).astype(np.float32)
    y = rng.random((4, 64 * 64)).astype(np.float32)
    lhs = np.vdot(be.forward(x).numpy().astype(np.float64).ravel(), y.astype(np.float64).ravel())
    rhs = np.vdot(x.astype(np.float64).ravel(), be.adjoint(y).numpy().astype(np.float64).ravel())
    assert abs(lhs - rhs) <= 1e-6 * abs(rhs)


def test_orphan_forward_project_variant():
    """ProjectionMatrix.forward_project = the orphan forward_project (src/forward_projection.f90): NINT sample count, cor_shift
    ignored -- against the float32 restatement of the Fortran (oracle.forward_project_orphan), on a volume WIDER than deep so
    that the extra trailing sample the NINT rule adds does touch voxels (the live path would drop it for some views)."""
    from tomography_alignment_b200 import ProjectionMatrix
    shape, dshape, n_proj = (30, 8, 12), (30, 12), 5
    g, og = make_geoms(shape, dshape, n_proj, cor=[0.6, 0.0, 0.0])           # the orphan ignores this shift
    phi, alpha, beta, xyz = random_poses(n_proj, 12, tilt=0.03, shift=1.0, phis=[0.1, 0.9, 1.5, 2.2, 3.0])
    rec = np.random.default_rng(7).random(shape).astype(np.float32)
    pm = ProjectionMatrix(g, backend=EmuBackend(g))
    ax = pm.forward_project(rec, alpha, beta, phi, xyz, cor_shift=g.cor_shift)
    ref = O.forward_project_orphan(og, rec, alpha, beta, phi, xyz)
    assert ax.shape == (n_proj, g.n_det) and ax.dtype == np.float32
    assert rel_l2(ax, ref) <= TOL_PROJ
    # and it is NOT the live operator here: the live path applies cor_shift (and may march one sample fewer)
    live = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).forward(rec)
    assert rel_l2(ax, live) > 1e-3


ZQ_CASES = [((24, 24, 24), (24, 24), 6, dict()), ((24, 20, 26), (24, 27), 5, dict(tilt=0.05, shift=3.0)),
            ((16, 16, 16), (20, 13), 5, dict(tilt=0.04, shift=5.0)), ((33, 19, 35), (33, 35), 3, dict(tilt=0.03, cor=[0.4, 0, 0])),
            ((16, 16, 16), (16, 16), 4, dict(step=0.5)), ((16, 16, 16), (16, 16), 4, dict(step=1.7)),
            ((12, 12, 12), (12, 12), 3, dict(shift=14.0)), ((5, 40, 3), (5, 3), 3, dict())]


@pytest.mark.parametrize("shape,dshape,n_proj,kw", ZQ_CASES)
def test_zquad_core_forward_and_gradient(shape, dshape, n_proj, kw):
    """csrc/zq_core.h (one thread = four z-adjacent rays, aligned 8-plane windows, irregular samples deferred to the exact
    float64 evaluation) against the oracle: ragged detector heights (ndz % 4 != 0), rays that leave the volume in z inside a
    group, steps != 1, big shifts."""
    g, og = make_geoms(shape, dshape, n_proj, cor=kw.get("cor"), step=kw.get("step", 1.0))
    phi, alpha, beta, xyz = random_poses(n_proj, 3, tilt=kw.get("tilt", 0.02), shift=kw.get("shift", 2.0))
    be = EmuBackend(g, zquad=True)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    assert np.all(be.views[:, 150] == 1.0)                   # V_ZQ: every view qualifies
    vol = np.random.default_rng(1).random(shape).astype(np.float32)
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    ref = op.forward(vol)
    if np.linalg.norm(ref) > 0:
        assert rel_l2(be.forward(vol).numpy().reshape(n_proj, -1), ref) <= TOL_PROJ
    out = be.proj_grad(vol)
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].numpy(), p) <= TOL_PROJ and rel_l2(out["dproj"][i].numpy(), gr) <= TOL_GRAD
    # views that do not qualify (large tilt) keep the per-ray core
    be2 = EmuBackend(g, zquad=True)
    be2.set_poses(pose_table(np.array([phi, alpha + 0.2, beta]).T, xyz, g.cor_shift))
    assert np.all(be2.views[:, 150] == 0.0)
