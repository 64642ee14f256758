"""Parity tests proper: the CUDA path, called through the C ABI (ctypes -> libtomo_b200.so), against the
CPU oracle on the same seeded inputs.  Tolerances are BASELINE.json's: relative L2 <= 1e-5 for projections
and backprojections, <= 1e-4 for the alignment gradients."""
import os

import numpy as np
import pytest
import torch
from scipy import sparse

from oracle import oracle as O
from tomography_alignment_b200 import Geometry, ProjectionMatrix, pose_table
from tomography_alignment_b200.phantom import benchmark_poses, shepp3d
from helpers import EmuBackend, make_geoms, random_poses, rel_l2

pytestmark = pytest.mark.gpu

TOL_PROJ, TOL_GRAD = 1e-5, 1e-4

CASES = [
    ((16, 16, 16), (16, 16), 6, dict()),
    ((24, 20, 18), (24, 18), 5, dict(cor=[0.7, 0, 0])),
    ((16, 16, 16), (20, 12), 5, dict(tilt=0.2, shift=5.0)),
    ((16, 16, 16), (16, 16), 4, dict(step=0.5)),
    ((16, 16, 16), (16, 16), 4, dict(step=1.7)),
    ((12, 12, 12), (12, 12), 3, dict(shift=14.0, phis=[0.2, 1.1, 2.0])),
    ((5, 40, 3), (5, 3), 3, dict()),
    ((33, 9, 35), (33, 35), 2, dict(phis=[0.4, 2.9])),
    ((40, 40, 40), (40, 40), 1, dict(phis=[0.77])),                       # n_proj == 1
    ((20, 56, 60), (20, 60), 4, dict(phis=[0.3, 1.9, 3.5, 5.2])),         # not a cube, laid out with the pitches of 64^3 (tomo_pad_pitch)
]


def cuda_backend(g):
    from tomography_alignment_b200.cuda_backend import CudaBackend
    return CudaBackend(g, "cuda:0")


def setup(shape, dshape, n_proj, cor=None, step=1.0, tilt=0.02, shift=2.0, phis=None, seed=0):
    g, og = make_geoms(shape, dshape, n_proj, cor=cor, step=step)
    phi, alpha, beta, xyz = random_poses(n_proj, seed, tilt=tilt, shift=shift, phis=phis)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    return g, og, be, op, (phi, alpha, beta, xyz)


def test_native_library_is_loaded():
    import ctypes
    from tomography_alignment_b200 import _lib
    L = _lib.load()
    assert isinstance(L, ctypes.CDLL) and L.tomo_version() >= 100
    maps = open("/proc/self/maps").read()
    assert "libtomo_b200.so" in maps


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES)
def test_forward_and_adjoint_vs_oracle(shape, dshape, n_proj, kw):
    g, og, be, op, _ = setup(shape, dshape, n_proj, **kw)
    rng = np.random.default_rng(1)
    vol = rng.random(shape).astype(np.float32)
    ref = op.forward(vol)
    got = be.forward(torch.as_tensor(vol)).cpu().numpy().reshape(n_proj, -1)
    if np.linalg.norm(ref) > 0:
        assert rel_l2(got, ref) <= TOL_PROJ
    else:
        assert np.abs(got).max() == 0.0
    y = rng.random((n_proj, og.n_det)).astype(np.float32)
    refb = op.adjoint(y)
    for gather in (False, True):            # tile-scatter kernel and the independent per-voxel gather kernel
        gotb = be.adjoint(torch.as_tensor(y), gather=gather).cpu().numpy().ravel()
        if np.linalg.norm(refb) > 0:
            assert rel_l2(gotb, refb) <= TOL_PROJ, gather
        else:
            assert np.abs(gotb).max() == 0.0
    # accumulate flag: vol += A^T y
    base = torch.full(tuple(shape), 0.5, dtype=torch.float32, device="cuda")
    be.adjoint(torch.as_tensor(y), out=base, accumulate=True)
    assert rel_l2(base.cpu().numpy().ravel() - 0.5, refb) <= 2e-5 or np.linalg.norm(refb) == 0


@pytest.mark.parametrize("tilt,phis", [(0.0, [0.0, np.pi / 2, np.pi, 0.3]), (0.02, None), (0.12, None), (0.45, None)])
def test_tile_scatter_adjoint_many_tiles(tilt, phis):
    """80 x 72 x 70 volume = 5 x 5 x 3 tiles of 16 x 16 x 30 (ragged last tiles), 12 views: exercises tile
    ownership, ghost cells, colour classes (C grows with the tilt; 0.45 rad leaves the scatter envelope for
    some views and falls back to the in-kernel gather) and the same-z-cell lane deferral."""
    shape, dshape, n_proj = (80, 72, 70), (84, 76), 12 if phis is None else 4
    g, og, be, op, _ = setup(shape, dshape, n_proj, tilt=tilt, shift=3.0 if tilt else 0.0, phis=phis, seed=11)
    y = np.random.default_rng(12).random((n_proj, og.n_det)).astype(np.float32)
    ref = op.adjoint(y)
    got = be.adjoint(torch.as_tensor(y))
    assert rel_l2(got.cpu().numpy(), ref) <= TOL_PROJ
    assert rel_l2(be.adjoint(torch.as_tensor(y), gather=True).cpu().numpy(), ref) <= TOL_PROJ
    for _ in range(2):                      # bitwise reproducible
        assert torch.equal(be.adjoint(torch.as_tensor(y)), got)


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES)      # all shapes, incl. the wide-thin volume where the trailing sample counts
def test_projection_gradient_vs_oracle(shape, dshape, n_proj, kw):
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    rng = np.random.default_rng(2)
    vol = rng.random(shape).astype(np.float32)
    meas = (op.forward(vol) * 1.02 + 0.05).astype(np.float32)
    out = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(meas))
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].cpu().numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD
        res = meas[i].astype(np.float64) - p
        assert rel_l2(out["grad6"][i].cpu().numpy(), -gr @ res) <= TOL_GRAD
        assert abs(out["cost"][i].item() - 0.5 * res @ res) <= 1e-5 * (0.5 * res @ res)


@pytest.mark.parametrize("shape,dshape,n_proj,kw", CASES[:3])
def test_voxel_driven_bilinear_backprojector(shape, dshape, n_proj, kw):
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    y = np.random.default_rng(3).random((n_proj,) + tuple(dshape)).astype(np.float32)
    ref = O.voxel_back_project(og, y, alpha, beta, phi, xyz)
    assert rel_l2(be.voxel_back(torch.as_tensor(y)).cpu().numpy(), ref) <= TOL_PROJ


VOXBACK_TMA_CASES = [
    # (volume, detector, n_proj, make_geoms kwargs, pose kwargs, origin offset)
    ((40, 36, 52), (48, 56), 7, dict(), dict(tilt=0.02, shift=2.0), None),                 # partial bricks, all views staged by TMA
    ((32, 32, 64), (64, 72), 6, dict(), dict(tilt=0.0, shift=0.0, phis=[0.0, np.pi / 2, np.pi, 0.3, 1.0, 2.2]), None),  # exact lattice hits
    ((48, 40, 40), (40, 44), 6, dict(), dict(tilt=0.3, shift=9.0), None),                  # big tilts: some views exceed the box -> plain kernel adds them
    ((36, 36, 36), (56, 60), 5, dict(vox_pix=[1.3, 0.8, 1.1]), dict(tilt=0.05, shift=3.0), None),   # anisotropic voxels
    ((40, 40, 40), (40, 48), 5, dict(), dict(tilt=0.02, shift=25.0), [7.5, 0.0, -11.25]),  # footprints leave the detector: TMA zero fill = bounds checks
    ((24, 24, 40), (44, 42), 4, dict(), dict(tilt=0.02, shift=1.0), None),                 # ndz % 4 != 0: the plain kernel alone
]


@pytest.mark.parametrize("shape,dshape,n_proj,gkw,pkw,dorg", VOXBACK_TMA_CASES)
def test_voxel_driven_backprojector_tma_staged(shape, dshape, n_proj, gkw, pkw, dorg):
    """tomo_back_voxel_bilinear at sizes where the TMA-staged kernel runs (detector >= 32 x 44, ndz % 4 == 0): brick footprints,
    zero fill outside the detector, views too tilted for the staged box, accumulate flag."""
    g, og = make_geoms(shape, dshape, n_proj, **gkw)
    phi, alpha, beta, xyz = random_poses(n_proj, 17, **pkw)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    y = np.random.default_rng(3).random((n_proj,) + tuple(dshape)).astype(np.float32)
    origin = None if dorg is None else np.asarray(og.det_orig, dtype=np.float64) + np.asarray(dorg)
    ref = O.voxel_back_project(og, y, alpha, beta, phi, xyz, origin=origin)
    got = be.voxel_back(torch.as_tensor(y), origin=origin)
    assert rel_l2(got.cpu().numpy(), ref) <= TOL_PROJ
    # accumulate into an existing volume, and bitwise reproducibility
    base = torch.full(tuple(shape), 0.25, dtype=torch.float32, device="cuda")
    acc = be.voxel_back(torch.as_tensor(y), origin=origin, out=base.clone(), accumulate=True)
    assert rel_l2(acc.cpu().numpy(), ref.reshape(shape) + 0.25) <= TOL_PROJ
    assert torch.equal(got, be.voxel_back(torch.as_tensor(y), origin=origin))


def test_voxel_backprojector_tma_equals_plain_kernel_at_256():
    """Size-independent check at 256^3 x 12 jittered views: the TMA-staged kernel (float32 box-local positions) against the
    plain gather kernel (float64 positions per voxel), which the same call takes when the projections are not 16-byte aligned."""
    n, n_proj = 256, 12
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    gen = torch.Generator(device="cuda").manual_seed(5)
    buf = torch.rand(n_proj * n * n + 1, device="cuda", generator=gen)
    y_unaligned = buf[1:].view(n_proj, n, n)                  # base address = 4 mod 16: plain kernel
    y_aligned = y_unaligned.clone()                           # fresh allocation: TMA kernel
    assert y_unaligned.data_ptr() % 16 != 0 and y_aligned.data_ptr() % 16 == 0
    v_tma = be.voxel_back(y_aligned)
    v_plain = be.voxel_back(y_unaligned)
    err = torch.linalg.vector_norm((v_tma - v_plain).double()) / torch.linalg.vector_norm(v_plain.double())
    assert float(err) <= TOL_PROJ
    assert float(v_plain.abs().max()) > 0


@pytest.mark.parametrize("tilt", [0.02, 0.0])
def test_banded_launch_order_with_a_ragged_last_band(tilt):
    """200 detector columns = 25 x-tiles = two bands of 13 (the last one ragged): forward, gradient and both adjoints of the
    generic (tilt != 0) and separable (tilt == 0) kernels against the oracle."""
    shape, dshape, n_proj = (200, 120, 36), (200, 36), 3        # ny >= nx / 1.7: see DESIGN "known non-parity 2"
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, tilt=tilt, shift=1.5, phis=[0.2, 1.3, 2.6], seed=9)
    rng = np.random.default_rng(2)
    vol = rng.random(shape).astype(np.float32)
    y = rng.random((n_proj,) + dshape).astype(np.float32)
    assert rel_l2(be.forward(torch.as_tensor(vol)).cpu().numpy(), op.forward(vol)) <= TOL_PROJ
    assert rel_l2(be.adjoint(torch.as_tensor(y)).cpu().numpy(), op.adjoint(y)) <= TOL_PROJ
    out = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(y), want_proj=True, want_dproj=True)
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].cpu().numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD
        res = y[i].ravel().astype(np.float64) - p
        assert rel_l2(out["grad6"][i].cpu().numpy(), -gr @ res) <= TOL_GRAD


def test_reference_numpy_golden_fixtures():
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", "ref_numpy_cases.npz"))
    for name in gold["case_names"]:
        q = lambda k: gold["%s/%s" % (name, k)]
        alpha, beta, phi = (float(v) for v in q("pose"))
        g, _ = make_geoms(tuple(q("vox_shape")), tuple(q("det_shape")), 1, cor=q("cor"), step=float(q("step")))
        be = cuda_backend(g)
        be.set_poses(pose_table(np.array([[phi, alpha, beta]]), q("xyz")[None, :], q("cor")[None, :]))
        out = be.proj_grad(torch.as_tensor(q("rec").astype(np.float32)))
        assert rel_l2(out["proj"][0].cpu().numpy(), q("proj")) <= TOL_PROJ, name
        assert rel_l2(out["dproj"][0].cpu().numpy(), q("grad")) <= TOL_GRAD, name
        shape = tuple(q("vox_shape"))
        at = be.adjoint(torch.as_tensor(q("y")[None, :].astype(np.float32))).cpu().numpy().reshape(shape)
        assert rel_l2(at[1:-1, 1:-1, 1:-1], q("At_dot_y").reshape(shape)[1:-1, 1:-1, 1:-1]) <= TOL_PROJ, name
        # per-corner boundary semantics (volume non-zero on its shell): the reference's numpy twin on the zero-padded problem
        full = be.forward(torch.as_tensor(q("rec_full").astype(np.float32)))[0].cpu().numpy()
        assert rel_l2(full, q("A_dot_rec_full_padded")) <= TOL_PROJ, name
        for gather in (False, True):
            atf = be.adjoint(torch.as_tensor(q("y_full")[None, :].astype(np.float32)), gather=gather).cpu().numpy()
            assert rel_l2(atf, q("At_dot_y_full_padded")) <= TOL_PROJ, (name, gather)
        outf = be.proj_grad(torch.as_tensor(q("rec_full").astype(np.float32)))
        assert rel_l2(outf["proj"][0].cpu().numpy(), q("proj_full_padded")) <= TOL_PROJ, name
        assert rel_l2(outf["dproj"][0].cpu().numpy(), q("grad_full_padded")) <= TOL_GRAD, name


def test_gpu_equals_cpu_emulation_of_the_same_cores():
    """The kernels and tests/emu run the same __host__ __device__ code; results agree to rounding."""
    shape, dshape, n_proj = (20, 24, 28), (20, 28), 4
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, seed=5)
    emu = EmuBackend(g)
    emu.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    vol = np.random.default_rng(6).random(shape).astype(np.float32)
    a = be.proj_grad(torch.as_tensor(vol))
    b = emu.proj_grad(vol)
    assert rel_l2(a["proj"].cpu().numpy(), b["proj"].numpy()) < 1e-6
    assert rel_l2(a["dproj"].cpu().numpy(), b["dproj"].numpy()) < 1e-5


def test_config0_64cubed_90_views_shepp_logan():
    """BASELINE.json configs[0]: 64^3 phantom, 90 views with the jitter of examples/generate_data.py."""
    n, n_proj = 64, 90
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    vol = shepp3d(n)
    ref = op.forward(vol)
    got = be.forward(torch.as_tensor(vol)).cpu().numpy().reshape(n_proj, -1)
    assert rel_l2(got, ref) <= TOL_PROJ
    refb = op.adjoint(ref)
    gotb = be.adjoint(torch.as_tensor(ref.astype(np.float32))).cpu().numpy().ravel()
    assert rel_l2(gotb, refb) <= TOL_PROJ
    # gradients of a subset of views at perturbed poses (what one alignment step evaluates)
    sel = [0, 17, 44, 45, 89]
    rec = (vol * 0.9).astype(np.float32)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T[sel] * np.array([1.0, 0.5, 0.5]), xyz[sel] * 0.5,
                            g.cor_shift[sel]))
    meas = torch.as_tensor(ref[sel].astype(np.float32))
    out = be.proj_grad(torch.as_tensor(rec), meas=meas)
    for k, i in enumerate(sel):
        p, gr = O.forward_proj_grad(og, alpha[i] * 0.5, beta[i] * 0.5, phi[i], xyz[i] * 0.5, og.cor_shift[i], rec)
        assert rel_l2(out["proj"][k].cpu().numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][k].cpu().numpy(), gr) <= TOL_GRAD, i
        res = ref[i] - p
        assert rel_l2(out["grad6"][k].cpu().numpy(), -gr @ res) <= TOL_GRAD, i


def test_results_are_bitwise_deterministic():
    g, og, be, op, _ = setup((48, 48, 48), (48, 48), 12, seed=3)
    rng = np.random.default_rng(4)
    vol = torch.as_tensor(rng.random((48, 48, 48)).astype(np.float32)).cuda()
    y = torch.as_tensor(rng.random((12, 48 * 48)).astype(np.float32)).cuda()
    f1, b1 = be.forward(vol).clone(), be.adjoint(y).clone()
    g1 = be.proj_grad(vol, meas=y)
    g1 = {k: v.clone() for k, v in g1.items()}
    for _ in range(3):
        assert torch.equal(be.forward(vol), f1) and torch.equal(be.adjoint(y), b1)
        g2 = be.proj_grad(vol, meas=y)
        assert torch.equal(g2["grad6"], g1["grad6"]) and torch.equal(g2["cost"], g1["cost"])
        assert torch.equal(g2["dproj"], g1["dproj"])


def test_adjointness_at_256cubed():
    """Size-independent property at a BASELINE.json size: <A x, y> = <x, A^T y> (float64 dot products of the
    float32 results), 256^3 with a handful of jittered views."""
    n, n_proj = 256, 6
    g, _ = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(360)
    sel = [0, 50, 123, 180, 271, 359]
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T[sel], xyz[sel], g.cor_shift))
    gen = torch.Generator(device="cuda").manual_seed(7)
    x = torch.rand((n, n, n), device="cuda", generator=gen)
    y = torch.rand((n_proj, n, n), device="cuda", generator=gen)
    lhs = torch.dot(be.forward(x).double().ravel(), y.double().ravel()).item()
    rhs = torch.dot(x.double().ravel(), be.adjoint(y).double().ravel()).item()
    assert abs(lhs - rhs) <= 2e-6 * abs(rhs)
    # linearity of the forward projector
    x2 = torch.rand((n, n, n), device="cuda", generator=gen)
    lin = be.forward(x + 2 * x2) - (be.forward(x) + 2 * be.forward(x2))
    assert lin.abs().max().item() <= 1e-4 * be.forward(x).abs().max().item()


def test_phi0_projection_is_sum_over_y_at_256():
    n = 256
    g, _ = make_geoms((n, n, n), (n, n), 1)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([[0.0, 0.0, 0.0]]), np.zeros((1, 3)), np.zeros((1, 3))))
    x = torch.rand((n, n, n), device="cuda")
    p = be.forward(x)[0]
    assert torch.allclose(p, x.double().sum(dim=1).float(), rtol=1e-5, atol=1e-4)


def test_drop_in_operator_on_gpu_matches_reference_csr():
    """ProjectionMatrix through the scipy idioms of recon/sirt.py, on the GPU, against the reference's CSR."""
    g, og = make_geoms((12, 12, 12), (12, 12), 5)
    phi, alpha, beta, xyz = random_poses(5, 9)
    pm = ProjectionMatrix(g, precision=np.float32, device="cuda:0")
    A = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    R = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz).csr(np.float64)
    rng = np.random.default_rng(0)
    x, y = rng.random(g.n_vox).astype(np.float32), rng.random(5 * g.n_det).astype(np.float32)
    ax = sparse.csr_matrix.dot(A, x)
    aty = sparse.csc_matrix.dot(sparse.csr_matrix.transpose(A), y)
    assert isinstance(ax, np.ndarray) and isinstance(aty, np.ndarray)
    assert rel_l2(ax, R @ x) <= TOL_PROJ and rel_l2(aty, R.T @ y) <= TOL_PROJ
    xt = torch.as_tensor(x).cuda()
    assert (A @ xt).is_cuda and rel_l2((A @ xt).cpu().numpy(), R @ x) <= TOL_PROJ
    proj, grad = pm.projection_gradient(x.reshape(12, 12, 12), alpha[1], beta[1], phi[1], xyz[1], g.cor_shift[1])
    p, gr = O.projection_gradient(og, x, alpha[1], beta[1], phi[1], xyz[1], og.cor_shift[1])
    assert proj.dtype == np.float32 and grad.shape == (6, g.n_det)
    assert rel_l2(proj, p) <= TOL_PROJ and rel_l2(grad, gr) <= TOL_GRAD


def test_host_buffer_entry_points_match_device_path():
    """forward_host / adjoint_host / proj_grad_host (copies overlapped with the kernels in view chunks, sub-tables
    of the view table) against the device-resident calls; also through the numpy path of the operator."""
    g, og, be, op, (phi, alpha, beta, xyz) = setup((40, 36, 44), (40, 44), 11, seed=21)
    rng = np.random.default_rng(22)
    vol = rng.random((40, 36, 44)).astype(np.float32)
    y = rng.random((11, 40, 44)).astype(np.float32)
    f_dev = be.forward(torch.as_tensor(vol)).cpu()
    b_dev = be.adjoint(torch.as_tensor(y)).cpu()
    g_dev = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(y), want_dproj=False)
    # queued volume download: valid after sync_host(); the uploaded volume stays usable by the next operator (vol_dev)
    be.forward_host(torch.as_tensor(vol).pin_memory())
    out_q = torch.zeros((40, 36, 44), dtype=torch.float32).pin_memory()
    be.adjoint_host(torch.as_tensor(y).pin_memory(), out_host=out_q, wait=False)
    g6_q, c_q = be.proj_grad_host(None, torch.as_tensor(y).pin_memory(), vol_dev=be._buf("vol", be.vol_shape))
    be.sync_host()
    assert torch.equal(out_q, b_dev) and torch.equal(g6_q, g_dev["grad6"].cpu()) and torch.equal(c_q, g_dev["cost"].cpu())
    for chunk in (None, 3, 11, 50):
        f_h = be.forward_host(torch.as_tensor(vol).pin_memory(), chunk_views=chunk)
        assert f_h.device.type == "cpu" and torch.equal(f_h.reshape(f_dev.shape), f_dev)
        b_h = be.adjoint_host(y, chunk_views=chunk)
        assert rel_l2(b_h.numpy(), b_dev.numpy()) < 1e-6
        g6, c = be.proj_grad_host(vol, y, chunk_views=chunk)
        assert torch.equal(g6, g_dev["grad6"].cpu()) and torch.equal(c, g_dev["cost"].cpu())
    assert rel_l2(b_h.numpy(), op.adjoint(y.reshape(11, -1))) <= TOL_PROJ
    pm = ProjectionMatrix(g, device="cuda:0", backend=be)
    A = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    ax = A @ vol.ravel()
    assert isinstance(ax, np.ndarray) and np.array_equal(ax, f_dev.numpy().ravel())
    assert rel_l2(A.T @ y.ravel(), b_dev.numpy()) < 1e-6


@pytest.mark.parametrize("shape,dshape,kw", [((12, 10, 14), (12, 14), dict()), ((16, 16, 16), (20, 12), dict(tilt=0.1, shift=3.0)),
                                              ((10, 10, 10), (10, 10), dict(cor=[0.4, 0.0, -0.3]))])
def test_voxel_driven_splat_and_gradient_image(shape, dshape, kw):
    """tomo_voxel_splat against the restated vox_wt_grad.bilinear_vox_interp (src/vox_wt_grad.f90:1-55)."""
    n_proj = 3
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    rec = np.random.default_rng(5).random(shape).astype(np.float32)
    det, grad = be.voxel_splat(torch.as_tensor(rec))
    for i in range(n_proj):
        d_ref, g_ref = O.voxel_forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], rec)
        assert rel_l2(det[i].cpu().numpy(), d_ref) <= TOL_PROJ
        assert rel_l2(grad[i].cpu().numpy(), g_ref) <= TOL_GRAD
    pm = ProjectionMatrix(g, device="cuda:0")
    d1, g1 = pm.voxel_projection_gradient(rec, alpha[0], beta[0], phi[0], xyz[0], g.cor_shift[0])
    d_ref, g_ref = O.voxel_forward_proj_grad(og, alpha[0], beta[0], phi[0], xyz[0], og.cor_shift[0], rec)
    assert d1.shape == (g.n_det,) and g1.shape == (6, g.n_det) and rel_l2(d1, d_ref) <= TOL_PROJ and rel_l2(g1, g_ref) <= TOL_GRAD


def test_device_resident_sirt_and_cgls_on_gpu():
    """recon.SIRT / recon.CGLS (device-resident loops) on a 32^3 phantom: same iterates as the CPU emulation of the
    same kernels, error decreasing, reference return conventions."""
    from tomography_alignment_b200.recon import CGLS, SIRT
    n, n_proj = 32, 24
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    angles = np.array([phi, alpha, beta]).T
    truth = shepp3d(n)
    be = cuda_backend(g)
    be.set_poses(pose_table(angles, xyz, g.cor_shift))
    b = be.forward(torch.as_tensor(truth)).cpu().numpy().reshape(n_proj, -1)
    s = SIRT(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
    rec, err = s.run_main_iteration(niter=15, positivity=True)
    assert rec.shape == (n, n, n) and len(err) == 15 and np.all(np.diff(err) < 0) and err[-1] < 0.75 * err[0]
    s_emu = SIRT(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
    rec_e, err_e = s_emu.run_main_iteration(niter=3, positivity=True)
    np.testing.assert_allclose(err[:3], err_e, rtol=1e-4)
    c = CGLS(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
    rec_c, err_c = c.run_main_iteration(niter=10)
    assert err_c[-1] < err[9]            # CGLS converges faster than SIRT per iteration


def test_device_resident_tikhonov_and_lasso_on_gpu():
    """recon.RegularizedRecon.run_tikhonov_gd / run_lasso_ista / run_lasso_accelerated on the GPU: the same iterates as the
    CPU emulation of the same kernels (the host loops are checked against recon/regularized.py in the CPU tier)."""
    from tomography_alignment_b200.recon import RegularizedRecon
    n, n_proj = 24, 12
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    angles = np.array([phi, alpha, beta]).T
    truth = shepp3d(n)
    be = cuda_backend(g)
    be.set_poses(pose_table(angles, xyz, g.cor_shift))
    b = be.forward(torch.as_tensor(truth)).cpu().numpy().reshape(n_proj, -1)
    for name, kw in (("run_tikhonov_gd", dict(niter=4, reg_param=0.2)), ("run_lasso_ista", dict(niter=3, reg_param=0.01)),
                     ("run_lasso_accelerated", dict(niter=3, reg_param=0.01))):
        gpu = RegularizedRecon(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
        emu = RegularizedRecon(g, b, angles, xyz, options={"ground_truth": truth}, backend=EmuBackend(g))
        rec, err = getattr(gpu, name)(**kw)
        rec_e, err_e = getattr(emu, name)(**kw)
        assert len(err) == len(err_e) and err[-1] < 1.0, name
        assert rel_l2(rec.ravel(), rec_e.ravel()) < 5e-5, name
        np.testing.assert_allclose(err, err_e, rtol=1e-4, err_msg=name)


def test_batched_alignment_recovers_jitter_on_gpu():
    """BatchedAlignment (all views per launch) recovers the xz shifts and the alpha/beta tilts of
    examples/generate_data.py-style jitter on a 48^3 phantom, like examples/align_rigid.py:40-52 does per view."""
    from tomography_alignment_b200.alignment import BatchedAlignment
    n, n_proj = 48, 10
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    rng = np.random.default_rng(31)
    phi = np.linspace(0.1, 3.0, n_proj)
    alpha, beta = rng.uniform(-0.012, 0.012, n_proj), rng.uniform(-0.012, 0.012, n_proj)
    xyz = np.zeros((n_proj, 3))
    xyz[:, 0], xyz[:, 2] = rng.uniform(-1.5, 1.5, n_proj), rng.uniform(-1.5, 1.5, n_proj)
    c = np.arange(n) - (n - 1) / 2
    X, Y, Z = np.meshgrid(c, c, c, indexing="ij")
    rec = (shepp3d(n) + 0.5 * np.exp(-((X - 6) ** 2 + (Y + 4) ** 2 + (Z - 3) ** 2) / 30.0)).astype(np.float32)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    meas = be.forward(torch.as_tensor(rec)).cpu().numpy().reshape(n_proj, -1)
    ba = BatchedAlignment(g, meas, np.array([phi, 0 * phi, 0 * phi]).T, np.zeros((n_proj, 3)), mode="xzab", device="cuda:0")
    f0, _ = ba.cost_and_gradient(torch.as_tensor(rec).cuda(), np.zeros((n_proj, 4)))
    x, f, it = ba.minimize(torch.as_tensor(rec).cuda(), bounds=((-3., 3.), (-3., 3.), (-0.02, 0.02), (-0.02, 0.02)),
                           maxiter=60)
    assert (f < 0.02 * f0).all(), (f / f0)
    assert np.abs(x[:, 0] - xyz[:, 0]).max() < 0.1 and np.abs(x[:, 1] - xyz[:, 2]).max() < 0.1
    assert np.abs(x[:, 2] - alpha).max() < 4e-3 and np.abs(x[:, 3] - beta).max() < 4e-3


@pytest.mark.parametrize("shape,dshape,kw", [((16, 16, 16), (16, 16), dict()), ((14, 20, 37), (14, 37), dict(cor=[0.4, 0, 0])),
                                              ((12, 12, 12), (18, 9), dict(shift=6.0)), ((40, 40, 300), (40, 300), dict()),
                                              ((16, 16, 16), (16, 16), dict(step=0.5)),
                                              ((20, 56, 60), (20, 60), dict())])       # pitches of 64^3, 64 planes to visit
def test_separable_forward_for_untilted_views(shape, dshape, kw):
    """alpha = beta = 0 (the default poses of projection_matrix): sep_forward_kernel, incl. volumes taller than one
    z chunk (300 planes = 3 chunks) and a table that mixes tilted and untilted views."""
    n_proj = 7
    phis = [0.0, 0.4, np.pi / 4, np.pi / 2, 2.0, 2.9, np.pi]
    g, og, be, op, _ = setup(shape, dshape, n_proj, tilt=0.0, phis=phis, **kw)
    assert bool((be.views[:, 146] == 1.0).all())
    vol = np.random.default_rng(8).random(shape).astype(np.float32)
    assert rel_l2(be.forward(torch.as_tensor(vol)).cpu().numpy().reshape(n_proj, -1), op.forward(vol)) <= TOL_PROJ
    y = np.random.default_rng(9).random((n_proj, og.n_det)).astype(np.float32)
    refb = op.adjoint(y)
    gotb = be.adjoint(torch.as_tensor(y))
    assert rel_l2(gotb.cpu().numpy(), refb) <= TOL_PROJ                         # sep_zgather + sep_adjoint kernels
    assert rel_l2(be.adjoint(torch.as_tensor(y), gather=True).cpu().numpy(), refb) <= TOL_PROJ
    assert torch.equal(be.adjoint(torch.as_tensor(y)), gotb)                     # bitwise reproducible
    base = torch.full(tuple(shape), 0.25, dtype=torch.float32, device="cuda")
    be.adjoint(torch.as_tensor(y), out=base, accumulate=True)
    assert rel_l2(base.cpu().numpy().ravel() - 0.25, refb) <= 2e-5
    # separable projection + gradient (views 0, 3, 6 sit exactly on lattice planes: one-sided derivative)
    _phi, _alpha, _beta, _xyz = _
    outg = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(y))
    for i in (1, 2, 4, 5):
        p, gr = O.forward_proj_grad(og, _alpha[i], _beta[i], _phi[i], _xyz[i], og.cor_shift[i], vol)
        assert rel_l2(outg["proj"][i].cpu().numpy(), p) <= TOL_PROJ
        assert rel_l2(outg["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD, i
        res = y[i].astype(np.float64) - p
        assert rel_l2(outg["grad6"][i].cpu().numpy(), -gr @ res) <= TOL_GRAD
        assert abs(outg["cost"][i].item() - 0.5 * res @ res) <= 1e-5 * (0.5 * res @ res)
    out2 = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(y))
    assert torch.equal(out2["grad6"], outg["grad6"]) and torch.equal(out2["dproj"], outg["dproj"])
    # mixed table: views 1, 4 tilted
    phi, alpha, beta, xyz = random_poses(n_proj, 3, tilt=0.0, phis=phis)
    alpha[[1, 4]] = [0.01, -0.02]
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    assert be.views[:, 146].cpu().tolist() == [1.0, 0.0, 1.0, 1.0, 0.0, 1.0, 1.0]
    mixed = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    assert rel_l2(be.forward(torch.as_tensor(vol)).cpu().numpy().reshape(n_proj, -1), mixed.forward(vol)) <= TOL_PROJ
    assert rel_l2(be.adjoint(torch.as_tensor(y)).cpu().numpy(), mixed.adjoint(y)) <= TOL_PROJ
    outm = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(y))
    for i in (1, 2, 4, 5):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(outm["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD, i
        assert rel_l2(outm["grad6"][i].cpu().numpy(), -gr @ (y[i].astype(np.float64) - p)) <= TOL_GRAD, i


@pytest.mark.parametrize("det_pix,tilt", [((1.5, 0.75), 0.02), ((0.8, 1.5), 0.02), ((1.5, 0.75), 0.0), ((0.8, 1.5), 0.0),
                                          ((1.0, 0.62), 0.03)])
def test_detector_pitch_differs_from_voxel_size(det_pix, tilt):
    """W_z = 0.75 / 0.62: adjacent lanes of the tile kernel share z cells all the time (second-pass deferral);
    W_z = 1.5: rows skip cells; untilted variants run the separable kernels (3-tap z gather)."""
    shape, dshape, n_proj = (44, 40, 70), (40, 90), 6
    g, og = make_geoms(shape, dshape, n_proj, det_pix=det_pix)
    phi, alpha, beta, xyz = random_poses(n_proj, 13, tilt=tilt, shift=1.5)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    rng = np.random.default_rng(14)
    vol = rng.random(shape).astype(np.float32)
    y = rng.random((n_proj, og.n_det)).astype(np.float32)
    assert rel_l2(be.forward(torch.as_tensor(vol)).cpu().numpy().reshape(n_proj, -1), op.forward(vol)) <= TOL_PROJ
    refb = op.adjoint(y)
    got = be.adjoint(torch.as_tensor(y))
    assert rel_l2(got.cpu().numpy(), refb) <= TOL_PROJ
    assert rel_l2(be.adjoint(torch.as_tensor(y), gather=True).cpu().numpy(), refb) <= TOL_PROJ
    assert torch.equal(be.adjoint(torch.as_tensor(y)), got)
    out = be.proj_grad(torch.as_tensor(vol))
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD


def test_error_codes_surface_as_exceptions():
    from tomography_alignment_b200 import _lib
    g, _ = make_geoms((8, 8, 8), (8, 8), 2)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.zeros((2, 3)), np.zeros((2, 3)), np.zeros((2, 3))))
    with pytest.raises(ValueError):
        be.forward(torch.zeros(7))
    with pytest.raises(ValueError):
        be.adjoint(torch.zeros(5))
    with pytest.raises(ValueError):
        be.proj_grad(torch.zeros((8, 8, 8)), want_grad6=True)
    with pytest.raises(_lib.TomoError):
        _lib.check(be.lib.tomo_forward(be._g(), None, 2, None, None, None), "tomo_forward")


@pytest.mark.parametrize("step,tilt", [(0.3, 0.05), (0.25, 0.12), (0.45, 0.02)])
def test_tile_scatter_with_small_steps_and_tilt(step, tilt):
    """step_size < 0.5 packs 1/step as many samples (and as much z drift per voxel) into a tile: the colour-class bound
    (csrc/views.cpp) must still keep same-colour rays apart -- tile kernel against the oracle, against the independent gather
    kernel, and bitwise repeatable."""
    shape, dshape, n_proj = (64, 56, 66), (70, 70), 8
    g, og, be, op, _ = setup(shape, dshape, n_proj, step=step, tilt=tilt, shift=2.0, seed=31)
    y = np.random.default_rng(32).random((n_proj, og.n_det)).astype(np.float32)
    ref = op.adjoint(y)
    got = be.adjoint(torch.as_tensor(y))
    assert rel_l2(got.cpu().numpy(), ref) <= TOL_PROJ
    assert rel_l2(be.adjoint(torch.as_tensor(y), gather=True).cpu().numpy(), ref) <= TOL_PROJ
    for _ in range(3):
        assert torch.equal(be.adjoint(torch.as_tensor(y)), got)


def test_out_buffers_are_validated():
    g, og, be, op, _ = setup((16, 16, 16), (16, 16), 3)
    vol = torch.rand((16, 16, 16), device="cuda")
    y = torch.rand((3, 16, 16), device="cuda")
    for bad in (torch.empty((3, 16, 15), device="cuda"), torch.empty((3, 16, 16)), torch.empty((3, 16, 16), device="cuda", dtype=torch.float64),
                torch.empty((3, 16, 32), device="cuda")[:, :, ::2]):
        with pytest.raises(ValueError):
            be.forward(vol, out=bad)
    for bad in (torch.empty((16, 16, 17), device="cuda"), torch.empty((16, 16, 16)), torch.empty((16, 32, 16), device="cuda")[:, ::2]):
        with pytest.raises(ValueError):
            be.adjoint(y, out=bad)
        with pytest.raises(ValueError):
            be.voxel_back(y, out=bad)
    out_np = np.zeros((3, 16, 16), np.float32)            # numpy out_host is filled in place
    r = be.forward_host(vol.cpu().numpy(), out_host=out_np)
    assert r is out_np and np.array_equal(out_np, be.forward(vol).cpu().numpy())


def test_operators_are_independent_on_the_gpu():
    """Two operators from one ProjectionMatrix (true vs estimated poses) keep their own view tables."""
    g, og = make_geoms((20, 20, 20), (20, 20), 4)
    phi, alpha, beta, xyz = random_poses(4, 9)
    pm = ProjectionMatrix(g, precision=np.float32, device="cuda:0")
    A1 = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    x = torch.rand(g.n_vox, device="cuda")
    y1 = (A1 @ x).clone()
    A2 = pm.projection_matrix(alpha=alpha, beta=beta, phi=phi + 0.3, xyz_shift=xyz)
    A3 = pm.projection_matrix(phi=phi[:2])
    r2 = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi + 0.3, xyz_shift=xyz)
    for _ in range(2):
        assert torch.equal(A1 @ x, y1) and (A1 @ x).numel() == 4 * g.n_det
        assert rel_l2((A2 @ x).cpu().numpy(), r2.forward(x.cpu().numpy()).ravel()) <= TOL_PROJ
        assert (A3 @ x).numel() == 2 * g.n_det
        yy = torch.rand(4 * g.n_det, device="cuda")
        assert rel_l2((A2.T @ yy).cpu().numpy(), r2.adjoint(yy.cpu().numpy().reshape(4, -1))) <= TOL_PROJ
        pm.projection_gradient(x.reshape(20, 20, 20), alpha[0], beta[0], phi[0], xyz[0], g.cor_shift[0])
    assert torch.equal(A1 @ x, y1)


@pytest.mark.parametrize("tilt", [0.02, 0.0])
def test_adjoint_in_x_slabs_equals_one_launch(tilt):
    """tomo_back_adjoint_slab: the volume backprojected slab by slab (what the overlapped multi-GPU all-reduce does) is bitwise the
    single launch -- tilted views (tile kernel), untilted views (separable adjoint: Yz filled by the first slab) and a mixed table."""
    shape, dshape, n_proj = (90, 40, 64), (96, 64), 6
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, tilt=tilt, shift=2.0, seed=41)
    if tilt:                                   # mixed table: two untilted views among the tilted ones
        alpha[1] = beta[1] = alpha[4] = beta[4] = 0.0
        be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
        op = O.OracleOperator(og, alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    y = torch.rand((n_proj,) + dshape, device="cuda")
    whole = be.adjoint(y)
    assert rel_l2(whole.cpu().numpy(), op.adjoint(y.cpu().numpy())) <= TOL_PROJ
    gx = be.slab_granularity()
    assert gx == 20 and be.slabs(3) == [(0, 40), (40, 60), (60, 90)] and be.slabs(50) == [(0, 20), (20, 40), (40, 60), (60, 80), (80, 90)]
    for n_slabs in (2, 3, 50):
        out = torch.full(shape, 7.0, device="cuda")
        for x0, x1 in be.slabs(n_slabs):
            be.adjoint(y, out=out, x_range=(x0, x1))
        assert torch.equal(out, whole)
    with pytest.raises(Exception):
        be.adjoint(y, out=torch.empty(shape, device="cuda"), x_range=(10, 40))        # not a tile-row boundary


def test_view_kinds_skip_launches_without_changing_results():
    """The TOMO_KINDS_* mask lets the operators skip kernels that would find no view; results equal the kinds = 0 path."""
    from tomography_alignment_b200 import _lib
    shape, dshape = (40, 36, 44), (40, 44)
    for tilt, expect in ((0.0, _lib.load().tomo_version() and (1 | 4)), (0.03, 1 | 2 | 8)):
        g, og, be, op, _ = setup(shape, dshape, 5, tilt=tilt, shift=1.0 if tilt else 0.0, seed=3)
        assert be.kinds == expect
        vol = torch.rand(shape, device="cuda")
        y = torch.rand((5,) + dshape, device="cuda")
        n0 = be.launches
        f, b, gr = be.forward(vol), be.adjoint(y), be.proj_grad(vol, meas=y)
        launched = be.launches - n0
        kinds = be.kinds
        be.kinds = 0                                  # unknown: every kernel is launched
        n0 = be.launches
        f0, b0, gr0 = be.forward(vol), be.adjoint(y), be.proj_grad(vol, meas=y)
        assert be.launches - n0 > launched
        be.kinds = kinds
        assert torch.equal(f, f0) and torch.equal(b, b0) and torch.equal(gr["grad6"], gr0["grad6"]) and torch.equal(gr["dproj"], gr0["dproj"])


@pytest.mark.parametrize("shape,dshape,kw", [((12, 10, 14), (12, 14), dict()), ((16, 16, 16), (20, 12), dict(tilt=0.1, shift=3.0)),
                                              ((10, 10, 10), (10, 10), dict(cor=[0.4, 0.0, -0.3]))])
def test_voxel_splat_deterministic_and_transpose(shape, dshape, kw):
    """tomo_voxel_splat_deterministic (64-bit fixed-point accumulation) against the restated bilinear_vox_interp, bitwise
    repeatable; tomo_voxel_splat_adjoint against the transposed COO matrix of bilinear_sparse (src/vox_wt_grad.f90:58-112);
    <S x, y> = <x, S^T y>."""
    n_proj = 3
    g, og, be, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    rec = np.random.default_rng(5).random(shape).astype(np.float32)
    det, grad = be.voxel_splat(torch.as_tensor(rec), deterministic=True)
    y = np.random.default_rng(6).random((n_proj, dshape[1], dshape[0])).astype(np.float32)       # [view][z'][x']
    st_ref = np.zeros(og.n_vox)
    for i in range(n_proj):
        d_ref, g_ref = O.voxel_forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], rec)
        assert rel_l2(det[i].cpu().numpy(), d_ref) <= 1e-6          # float32 output of an exact sum of float32 terms
        assert rel_l2(grad[i].cpu().numpy(), g_ref) <= TOL_GRAD
        dat, di, w = O.voxel_forward_sparse(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i])
        S = sparse.coo_matrix((w.astype(np.float64), (di, dat)), shape=(og.n_det, og.n_vox)).tocsr()
        st_ref += S.T @ y[i].ravel().astype(np.float64)
    for _ in range(2):
        d2, g2 = be.voxel_splat(torch.as_tensor(rec), deterministic=True)
        assert torch.equal(d2, det) and torch.equal(g2, grad)
    d_only, none = be.voxel_splat(torch.as_tensor(rec), want_grad=False, deterministic=True)
    assert none is None and torch.equal(d_only, det)
    st = be.voxel_splat_adjoint(torch.as_tensor(y))
    assert rel_l2(st.cpu().numpy(), st_ref) <= TOL_PROJ
    # the orphan backprojector (src/external_back_projection.f90) with origin = vox_origin - cor_shift is the same operator on the
    # transposed detector image: two kernels (TMA-staged / plain gather vs splat transpose) and two Fortran routines agree
    y_xz = torch.as_tensor(np.ascontiguousarray(y.transpose(0, 2, 1)))
    vb = be.voxel_back(y_xz, origin=np.asarray(g.vox_origin) - np.asarray(g.cor_shift[0]))
    assert rel_l2(vb.cpu().numpy(), st.cpu().numpy()) <= TOL_PROJ
    lhs = float((det.double().cpu() * torch.as_tensor(y).double()).sum())
    rhs = float((st.double().cpu().ravel() * torch.as_tensor(rec).double().ravel()).sum())
    assert abs(lhs - rhs) <= 1e-5 * abs(rhs)


def test_voxel_splat_at_256_cubed_many_views():
    """The splat's view index lives in grid.x (r1 capped ny/4 * n_proj at 65535: TOMO_E_RANGE at 512^3 x 720): 256^3 x 1100
    views = 70400 (y block, view) pairs; one seeded view of the result against the oracle, the deterministic variant bitwise
    repeatable and within float32 rounding of the atomic one."""
    n, n_proj = 256, 1100
    g, og = make_geoms((n, n, n), (n, n), n_proj)
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    be = cuda_backend(g)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    rec = torch.rand((n, n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    det, _ = be.voxel_splat(rec, want_grad=False)
    det_fx, _ = be.voxel_splat(rec, want_grad=False, deterministic=True)
    i = 777
    d_ref, _ = O.voxel_forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], rec.cpu().numpy())
    assert rel_l2(det[i].cpu().numpy(), d_ref) <= TOL_PROJ and rel_l2(det_fx[i].cpu().numpy(), d_ref) <= 1e-6
    assert rel_l2(det.cpu().numpy(), det_fx.cpu().numpy()) <= 1e-6
    d2, _ = be.voxel_splat(rec, want_grad=False, deterministic=True)
    assert torch.equal(d2, det_fx)


def test_orphan_forward_project_on_gpu():
    """ProjectionMatrix.forward_project (orphan semantics: NINT sample count, cor_shift ignored) against the float32
    restatement of src/forward_projection.f90."""
    shape, dshape, n_proj = (30, 8, 12), (30, 12), 5
    g, og = make_geoms(shape, dshape, n_proj, cor=[0.6, 0.0, 0.0])
    phi, alpha, beta, xyz = random_poses(n_proj, 12, tilt=0.03, shift=1.0, phis=[0.1, 0.9, 1.5, 2.2, 3.0])
    rec = np.random.default_rng(7).random(shape).astype(np.float32)
    pm = ProjectionMatrix(g, device="cuda:0")
    ax = pm.forward_project(rec, alpha, beta, phi, xyz, cor_shift=g.cor_shift)
    assert ax.shape == (n_proj, g.n_det) and rel_l2(ax, O.forward_project_orphan(og, rec, alpha, beta, phi, xyz)) <= TOL_PROJ


@pytest.mark.parametrize("shape,dshape,n_proj,kw", [((40, 36, 44), (40, 44), 5, dict()), ((33, 19, 35), (33, 35), 3, dict(tilt=0.03, cor=[0.4, 0, 0])),
                                                      ((64, 64, 70), (70, 66), 4, dict(tilt=0.05, shift=4.0)), ((16, 16, 16), (16, 16), 4, dict(step=0.5))])
def test_zquad_kernels_vs_oracle(shape, dshape, n_proj, kw):
    """CudaBackend(zquad=True): zq_kernel_forward / zq_kernel_gradient (four z-adjacent rays per thread, 128-bit window loads,
    csrc/zq_core.h) against the oracle and against the default per-ray kernels; bitwise repeatable."""
    from tomography_alignment_b200.cuda_backend import CudaBackend
    g, og, be0, op, (phi, alpha, beta, xyz) = setup(shape, dshape, n_proj, **kw)
    be = CudaBackend(g, "cuda:0", zquad=True)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    assert be.kinds == (1 | 8 | 32) and be0.kinds == (1 | 2 | 8)
    rng = np.random.default_rng(2)
    vol = rng.random(shape).astype(np.float32)
    meas = (op.forward(vol) * 1.02 + 0.05).astype(np.float32)
    f = be.forward(torch.as_tensor(vol))
    assert rel_l2(f.cpu().numpy().reshape(n_proj, -1), op.forward(vol)) <= TOL_PROJ
    out = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(meas))
    ref = be0.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(meas))
    for i in range(n_proj):
        p, gr = O.forward_proj_grad(og, alpha[i], beta[i], phi[i], xyz[i], og.cor_shift[i], vol)
        assert rel_l2(out["proj"][i].cpu().numpy(), p) <= TOL_PROJ and rel_l2(out["dproj"][i].cpu().numpy(), gr) <= TOL_GRAD
        res = meas[i].astype(np.float64) - p
        assert rel_l2(out["grad6"][i].cpu().numpy(), -gr @ res) <= TOL_GRAD
    assert rel_l2(out["grad6"].cpu().numpy(), ref["grad6"].cpu().numpy()) <= 1e-5
    for _ in range(2):
        assert torch.equal(be.forward(torch.as_tensor(vol)), f)
        o2 = be.proj_grad(torch.as_tensor(vol), meas=torch.as_tensor(meas))
        assert torch.equal(o2["grad6"], out["grad6"]) and torch.equal(o2["dproj"], out["dproj"])


def test_zquad_kernels_at_256_cubed():
    """The z-quad kernels at a BASELINE size: 256^3, 4 views of benchmark_poses(360), full outputs against the oracle."""
    from tomography_alignment_b200.cuda_backend import CudaBackend
    n, sel = 256, [3, 100, 181, 300]
    g, og = make_geoms((n, n, n), (n, n), len(sel))
    phi, alpha, beta, xyz = (a[sel] for a in benchmark_poses(360))
    be = CudaBackend(g, "cuda:0", zquad=True)
    be.set_poses(pose_table(np.array([phi, alpha, beta]).T, xyz, g.cor_shift))
    assert be.kinds & 32
    vol_d = torch.rand((n, n, n), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    vol = vol_d.cpu().numpy()
    fwd = be.forward(vol_d)
    out = be.proj_grad(vol_d)
    rays = np.arange(og.n_det)
    for k in range(len(sel)):
        p, gr = O.forward_proj_grad_rays(og, alpha[k], beta[k], phi[k], xyz[k], og.cor_shift[k], vol, rays)
        assert rel_l2(fwd[k].cpu().numpy(), p) <= TOL_PROJ and rel_l2(out["proj"][k].cpu().numpy(), p) <= TOL_PROJ
        assert rel_l2(out["dproj"][k].cpu().numpy(), gr) <= TOL_GRAD
