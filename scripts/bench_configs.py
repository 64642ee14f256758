"""BASELINE.json configs 1-3 at the solver level (one B200), device-resident:
  config 1  SIRT, 256^3 x 360 views, known geometry (forward + backprojection only): ms per iteration
  config 2  rigid alignment, 256^3 x 360 views: one batched cost + 6-DOF gradient evaluation of all views (what the reference
            does with 2 x 360 single-view calls per optimiser step) and one full alternation = BatchedAlignment.minimize
            (xzab, 3 iterations) + 5 SIRT iterations at the current poses
  config 3  CGLS and FISTA-TV, 512^3 x 720 views (single GPU here; bench.py covers the sharded operators): ms per iteration
Prints one JSON line per config."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tomography_alignment_b200 import Geometry, pose_table
from tomography_alignment_b200.alignment import BatchedAlignment
from tomography_alignment_b200.cuda_backend import CudaBackend
from tomography_alignment_b200.phantom import benchmark_poses, shepp3d
from tomography_alignment_b200.recon import CGLS, SIRT, RegularizedRecon


def timed(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3, r


def data(n, n_proj, tilted):
    g = Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    if not tilted:
        alpha, beta, xyz = alpha * 0, beta * 0, xyz * 0
    angles = np.array([phi, alpha, beta]).T
    truth = shepp3d(n, device="cuda")
    be = CudaBackend(g, "cuda:0")
    be.set_poses(pose_table(angles, xyz, g.cor_shift))
    b = be.forward(truth).cpu().numpy().reshape(n_proj, -1)
    return g, angles, xyz, truth.cpu().numpy(), b


which = sys.argv[1:] or ["1", "2", "3"]
if "1" in which:
    n, n_proj = 256, 360
    g, angles, xyz, truth, b = data(n, n_proj, tilted=False)
    s = SIRT(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
    s.run_main_iteration(niter=2)
    ms, (rec, err) = timed(lambda: s.run_main_iteration(niter=20, positivity=True))
    print(json.dumps({"config": "SIRT 256^3 x 360 views, known geometry", "ms_per_iteration": ms / len(err), "iterations": len(err),
                      "updates_per_s": 2.0 * n ** 3 * n_proj / (ms / len(err) * 1e-3), "rms_error_first_last": [float(err[0]), float(err[-1])]}))
if "2" in which:
    n, n_proj = 256, 360
    g, angles, xyz, truth, b = data(n, n_proj, tilted=True)
    est_angles, est_xyz = angles.copy(), xyz.copy()
    est_angles[:, 1:] = 0.0; est_xyz[:] = 0.0                      # start from the nominal geometry (examples/align_rigid.py)
    rec = torch.as_tensor(truth).cuda()
    ba = BatchedAlignment(g, b, est_angles, est_xyz, mode="xzab", device="cuda:0")
    ba.cost_and_gradient(rec, np.zeros((n_proj, 4)))
    ms_eval, _ = timed(lambda: ba.cost_and_gradient(rec, np.zeros((n_proj, 4))))
    def alternation():
        x, f, it = ba.minimize(rec, bounds=[(-3, 3), (-3, 3), (-0.02, 0.02), (-0.02, 0.02)], maxiter=3)
        a2, t2 = est_angles.copy(), est_xyz.copy()
        t2[:, 0] += x[:, 0]; t2[:, 2] += x[:, 1]; a2[:, 1] += x[:, 2]; a2[:, 2] += x[:, 3]
        s = SIRT(g, b, a2, t2, options={"ground_truth": truth}, device="cuda:0")
        s.run_main_iteration(niter=5, positivity=True)
        return x, ba.evaluations
    ms_alt, (x, evals) = timed(alternation)
    err_x = float(np.abs(x[:, 0] - xyz[:, 0]).mean()); err_a = float(np.abs(x[:, 2] - angles[:, 1]).mean())
    print(json.dumps({"config": "rigid alignment 256^3 x 360 views (xzab) alternating with SIRT", "ms_per_batched_cost_gradient": ms_eval,
                      "ms_per_alternation(3 align iters + 5 SIRT iters)": ms_alt, "proj_grad_evaluations": evals,
                      "mean_abs_error_tx_px_after": err_x, "mean_abs_error_alpha_rad_after": err_a,
                      "updates_per_s_of_one_evaluation": 1.0 * n ** 3 * n_proj / (ms_eval * 1e-3)}))
if "3" in which:
    n, n_proj = 512, 720
    g, angles, xyz, truth, b = data(n, n_proj, tilted=False)
    c = CGLS(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
    c.run_main_iteration(niter=1)
    ms_c, out = timed(lambda: c.run_main_iteration(niter=5))
    r = RegularizedRecon(g, b, angles, xyz, options={"ground_truth": truth}, device="cuda:0")
    # step 1/hyper must stay below 1/||A||^2 ~ 1/(n_proj * n): hyper = 1e6 here (the reference's default 1e4 is for its 64^3 example)
    ms_f, (rec_f, err_f) = timed(lambda: r.run_fista(niter=4, hyper=1.e6, beta_tv=0.1, niter_tv=10))
    print(json.dumps({"config": "CGLS / FISTA-TV 512^3 x 720 views, known geometry, 1 GPU", "cgls_ms_per_iteration": ms_c / 5,
                      "fista_tv_ms_per_iteration(10 TV prox iterations each)": ms_f / len(err_f), "fista_iterations": len(err_f),
                      "fista_rms_error": [float(e) for e in err_f]}))
