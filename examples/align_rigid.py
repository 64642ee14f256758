"""examples/align_rigid.py of the reference, on the B200 operators: alternate SIRT reconstruction with per-view
rigid alignment of (tx, tz, alpha, beta).

Two alignment drivers:
  --driver scipy     the reference's loop verbatim: scipy L-BFGS-B per view on cost_xzab / gradient_xzab
                     (examples/align_rigid.py:40-52) -- 2 single-view operator calls per optimiser step;
  --driver batched   (default) all views per launch with alignment.BatchedAlignment.

    python examples/generate_data.py --size 64 --views 90 --out data.npz
    python examples/align_rigid.py data.npz [--outer 4] [--sirt-iters 100] [--driver batched]
"""
import argparse
import copy
import os
import sys

import numpy as np
from scipy import optimize

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomography_alignment_b200 import alignment as alignment_functions        # noqa: E402
from tomography_alignment_b200 import geometry, projection_operators           # noqa: E402
from tomography_alignment_b200.alignment import cost_xzab, gradient_xzab, BatchedAlignment   # noqa: E402
from tomography_alignment_b200.recon import SIRT                               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("data")
    ap.add_argument("--outer", type=int, default=4)
    ap.add_argument("--sirt-iters", type=int, default=100)
    ap.add_argument("--driver", default="batched", choices=["batched", "scipy"])
    a = ap.parse_args()
    f = np.load(a.data)
    proj, alpha, beta, xyz, phi = (f["data/" + k] for k in ("projections", "alpha", "beta", "xyz", "phi"))
    ground_truth = f["data/phantom"]
    nx, ny, nz = ground_truth.shape
    n_proj = proj.shape[0]
    geom = geometry.Geometry(n_proj, np.array([nx, ny, nz]), np.ones(3, ), np.array([nx, nz]), np.ones(2, ))
    proj_obj = projection_operators.ProjectionMatrix(geom, precision=np.float32)

    alpha_rec, beta_rec = np.zeros(n_proj), np.zeros(n_proj)
    xyz_rec = np.zeros((n_proj, 3))
    rec = np.zeros_like(ground_truth)
    bounds = ((-3., 3.), (-3.0, 3.0), (-0.02, 0.02), (-0.02, 0.02))               # align_rigid.py:48
    for it in range(1, a.outer + 1):
        sirt_obj = SIRT(geom, proj.reshape(n_proj, -1), np.array([phi, alpha_rec, beta_rec]).T, xyz_rec,
                        options={'ground_truth': ground_truth, 'rec': rec.ravel()})
        rec, err = sirt_obj.run_main_iteration(niter=a.sirt_iters, positivity=True)
        if a.driver == "scipy":
            for i in range(n_proj):
                this_geo = copy.deepcopy(geom)
                this_geo.cor_shift = geom.cor_shift[i]
                align_obj = alignment_functions.AlignmentUtilities(proj[i], proj_obj, this_geo)
                params = np.array([xyz_rec[i, 0], xyz_rec[i, 2], alpha_rec[i], beta_rec[i]])
                res = optimize.minimize(cost_xzab, params, method='L-BFGS-B', jac=gradient_xzab,
                                        args=(align_obj, rec, np.array([phi[i], 0.0, 0.0]), np.zeros(3, )),
                                        bounds=bounds, options={'disp': False})
                xyz_rec[i, 0], xyz_rec[i, 2] = res.x[:2]
                alpha_rec[i], beta_rec[i] = res.x[2:]
        else:
            ba = BatchedAlignment(geom, proj.reshape(n_proj, -1), np.array([phi, 0 * phi, 0 * phi]).T,
                                  np.zeros((n_proj, 3)), mode="xzab")
            x0 = np.array([xyz_rec[:, 0], xyz_rec[:, 2], alpha_rec, beta_rec]).T
            x, cost, n_it = ba.minimize(rec, x0=x0, bounds=bounds, maxiter=40)
            xyz_rec[:, 0], xyz_rec[:, 2], alpha_rec, beta_rec = x[:, 0], x[:, 1], x[:, 2], x[:, 3]
        print("outer %d: SIRT RMSE %.4f | max |dx| %.3f px, |dz| %.3f px, |dalpha| %.4f deg, |dbeta| %.4f deg"
              % (it, err[-1], np.abs(xyz_rec[:, 0] - xyz[:, 0]).max(), np.abs(xyz_rec[:, 2] - xyz[:, 2]).max(),
                 np.rad2deg(np.abs(alpha_rec - alpha).max()), np.rad2deg(np.abs(beta_rec - beta).max())))
    np.savez_compressed(os.path.splitext(a.data)[0] + "_aligned.npz", rec=rec, alpha=alpha_rec, beta=beta_rec, xyz=xyz_rec)


if __name__ == "__main__":
    main()
