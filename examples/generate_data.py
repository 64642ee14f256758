"""examples/generate_data.py of the reference, on the B200 operators.

Builds the Shepp-Logan phantom, draws the per-view jitter of the reference script (examples/generate_data.py:16-23:
alpha, beta within +-1 degree, tx, tz within +-2 px) and simulates the misaligned projections.  The reference
never writes the HDF5 file its align_rigid.py reads (and h5py is not in this image): the data set is written as
an .npz with the same keys (data/projections, data/alpha, data/beta, data/xyz, data/phi, data/phantom).

    python examples/generate_data.py [--size 64] [--views 90] [--out data_64.npz] [--seed 20240229]
"""
import argparse
import os
import sys

import numpy as np
from scipy import sparse

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tomography_alignment_b200 import geometry, projection_operators          # noqa: E402
from tomography_alignment_b200.phantom import benchmark_poses, shepp3d         # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=64)
    ap.add_argument("--views", type=int, default=90)
    ap.add_argument("--out", default=None)
    ap.add_argument("--seed", type=int, default=20240229)
    a = ap.parse_args()
    nx = ny = nz = a.size
    n_proj = a.views

    shepp = shepp3d(nx)                                                         # generate_data.py:10
    geom = geometry.Geometry(n_proj, np.array([nx, ny, nz]), np.ones(3, ), np.array([nx, nz]), np.ones(2, ))
    phi, alpha, beta, xyz = benchmark_poses(n_proj, a.seed)                     # :16-23, seeded

    proj_obj = projection_operators.ProjectionMatrix(geom, precision=np.float32)
    pmat = proj_obj.projection_matrix(alpha=alpha, beta=beta, phi=phi, xyz_shift=xyz)
    proj = sparse.csr_matrix.dot(pmat, shepp.ravel()).reshape(n_proj, nx, nz)   # :29, unchanged idiom

    out = a.out or "data_%d_%d.npz" % (a.size, a.views)
    np.savez_compressed(out, **{"data/projections": proj, "data/alpha": alpha, "data/beta": beta, "data/xyz": xyz,
                                "data/phi": phi, "data/phantom": shepp})
    print("wrote %s: projections %s, |proj| = %.4f" % (out, proj.shape, np.linalg.norm(proj)))


if __name__ == "__main__":
    main()
