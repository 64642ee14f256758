"""Generate golden fixtures from the REFERENCE ITSELF (run in the build container only).

The reference's native kernels are Fortran and cannot be built here (no gfortran), but
the reference also ships pure-numpy twins of them and all of its setup math in numpy:

    utilities/geometry.py                  Geometry
    utilities/rotations.py                 rot_*, der_rot_*
    utilities/ray_voxel_utilities.py       transform_points, derivative_ray_points,
                                           ray_tracing_trilinear, ray_weights_der   (:6-50, :173-345)
    utilities/voxel_utilities.py           rigid_transformation, derivative_rigid   (:6-48)

They import once ``src.ray_wt_grad`` / ``src.vox_wt_grad`` (the f2py modules) are stubbed.  This script
imports them from /root/reference, runs them on small seeded cases and stores inputs + outputs in
tests/golden/ref_numpy_cases.npz.  /root/reference does not exist on the GPU box, so only the .npz
travels; tests/test_oracle_golden.py checks the oracle against it.

Caveat: the numpy twins keep a sample only when all 8 corners are inside the
volume, the live Fortran keeps every in-bounds corner.  The two agree exactly when the volume is zero
on its outermost one-voxel shell (every in-bounds corner of a boundary sample lies on that shell), so
the fixture volumes have a zero shell.  The Fortran's per-corner bounds checks themselves are pinned by a second
evaluation of the SAME reference-authored twin on the problem padded by one voxel of zeros on every side (volume
shape N + 2, ray points shifted by +1): there every sample that touches a real voxel has all 8 corners inside the
padded array, so the twin keeps it with all its in-bounds corners -- exactly the live Fortran's result for a volume
that is NOT zero on its shell ("*_full_padded" entries).

Usage:  python tests/golden/make_golden.py
"""
import copy
import os
import sys
import types

import numpy as np

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ref_numpy_cases.npz")


def import_reference():
    sys.path.insert(0, REF)
    import src  # the reference's (empty) package
    for name in ("ray_wt_grad", "vox_wt_grad"):
        stub = types.ModuleType("src." + name)
        sys.modules["src." + name] = stub
        setattr(src, name, stub)
    from utilities import geometry, rotations, ray_voxel_utilities, voxel_utilities
    return geometry, rotations, ray_voxel_utilities, voxel_utilities


CASES = [
    # name, vox_shape, det_shape, alpha, beta, phi, xyz_shift, cor_shift, step_size
    ("cube8_generic", (8, 8, 8), (8, 8), 0.011, -0.017, 0.4, (0.31, 0.0, -0.27), (0.0, 0.0, 0.0), 1.0),
    ("cube10_cor", (10, 10, 10), (10, 10), -0.02, 0.015, 1.9, (-1.3, 0.2, 0.8), (0.4, 0.0, 0.0), 1.0),
    ("box_10_12_9", (10, 12, 9), (10, 9), 0.005, 0.02, 2.6, (0.6, -0.4, 0.45), (-0.3, 0.0, 0.0), 1.0),
    ("cube8_halfstep", (8, 8, 8), (8, 8), 0.013, 0.009, 0.9, (0.2, 0.0, 0.1), (0.0, 0.0, 0.0), 0.5),
    ("cube9_bigtilt", (9, 9, 9), (9, 9), 0.15, -0.11, 0.75, (0.12, 0.3, -0.2), (0.1, 0.0, 0.0), 1.0),
    ("cube8_phi0_tilt", (8, 8, 8), (8, 8), 0.0123, 0.0077, 0.0, (0.37, 0.0, 0.21), (0.0, 0.0, 0.0), 1.0),
]


def main():
    geometry, rotations, rvu, vu = import_reference()
    rng = np.random.default_rng(20240229)
    out = {"case_names": np.array([c[0] for c in CASES])}
    for name, vshape, dshape, alpha, beta, phi, xyz, cor, step in CASES:
        vshape = np.array(vshape)
        dshape = np.array(dshape)
        xyz = np.array(xyz, dtype=np.float64)
        cor = np.array(cor, dtype=np.float64)
        geo = geometry.Geometry(1, vshape, np.ones(3), dshape, np.ones(2), cor_shift=cor, step_size=step)
        geo.cor_shift = cor                      # what projection_operators.py:102 / :115 hands down
        rec = rng.random(tuple(vshape))
        rec[0, :, :] = rec[-1, :, :] = 0.0
        rec[:, 0, :] = rec[:, -1, :] = 0.0
        rec[:, :, 0] = rec[:, :, -1] = 0.0
        # the setup forward_proj_grad does in numpy (ray_voxel_utilities.py:124-131)
        geo.source_centers[0, :] += geo.cor_shift[0]
        geo.det_centers[0, :] += geo.cor_shift[0]
        p0 = rvu.transform_points(geo.source_centers, alpha, beta, phi, xyz) - geo.vox_origin[:, np.newaxis]
        p1 = rvu.transform_points(geo.det_centers, alpha, beta, phi, xyz) - geo.vox_origin[:, np.newaxis]
        der = rvu.derivative_ray_points(geo.source_centers, (geo.det_centers - geo.source_centers)[:, 0],
                                        alpha, beta, phi, xyz)
        proj, grad = rvu.ray_weights_der(p0, p1, geo, (phi, alpha, beta), xyz, rec)
        wts, det_inds, inds, _ = rvu.ray_tracing_trilinear(vshape, p0, p1, geo.vox_ds, step, precision=np.float64)
        # dense interior matrix (duplicates summed) -> A x and A^T y of the interior samples
        A = np.zeros((geo.n_det, geo.n_vox))
        np.add.at(A, (det_inds, inds), wts)
        y = rng.random(geo.n_det)
        # per-corner boundary semantics: the twin on the zero-padded problem (see the module docstring); own generator so that
        # the entries above keep their values
        rng2 = np.random.default_rng(sum(ord(c) for c in name))
        rec_full = rng2.random(tuple(vshape))
        y_full = rng2.random(geo.n_det)
        pshape = vshape + 2
        rec_pad = np.zeros(tuple(pshape))
        rec_pad[1:-1, 1:-1, 1:-1] = rec_full
        wts_p, det_p, inds_p, _ = rvu.ray_tracing_trilinear(pshape, p0 + 1.0, p1 + 1.0, geo.vox_ds, step, precision=np.float64)
        Ap = np.zeros((geo.n_det, int(np.prod(pshape))))
        np.add.at(Ap, (det_p, inds_p), wts_p)
        out[name + "/rec_full"] = rec_full
        out[name + "/y_full"] = y_full
        out[name + "/A_dot_rec_full_padded"] = Ap.dot(rec_pad.ravel())
        out[name + "/At_dot_y_full_padded"] = Ap.T.dot(y_full).reshape(tuple(pshape))[1:-1, 1:-1, 1:-1].ravel()
        # the gradient twin on the same padded problem: it reads only vox_shape off the geometry for the bounds/indices, the pose
        # derivative tables do not depend on the volume's extent
        geo_p = copy.copy(geo)
        geo_p.vox_shape = pshape
        proj_p, grad_p = rvu.ray_weights_der(p0 + 1.0, p1 + 1.0, geo_p, (phi, alpha, beta), xyz, rec_pad)
        out[name + "/proj_full_padded"] = proj_p
        out[name + "/grad_full_padded"] = grad_p
        vox_rot = vu.rigid_transformation(geo.vox_centers, alpha, beta, phi, xyz)
        vox_der = vu.derivative_rigid(geo.vox_centers, alpha, beta, phi, xyz)
        pre = name + "/"
        out[pre + "vox_shape"] = vshape
        out[pre + "det_shape"] = dshape
        out[pre + "pose"] = np.array([alpha, beta, phi], dtype=np.float64)
        out[pre + "xyz"] = xyz
        out[pre + "cor"] = cor
        out[pre + "step"] = np.array(step)
        out[pre + "rec"] = rec
        out[pre + "y"] = y
        out[pre + "vox_origin"] = geo.vox_origin
        out[pre + "p0"] = p0
        out[pre + "p1"] = p1
        out[pre + "der"] = der
        out[pre + "proj"] = proj
        out[pre + "grad"] = grad
        out[pre + "A_dot_rec"] = A.dot(rec.ravel())
        out[pre + "At_dot_y"] = A.T.dot(y)
        out[pre + "vox_rot"] = vox_rot
        out[pre + "vox_der"] = vox_der
    # rotation matrices at a few angles (utilities/rotations.py)
    angs = np.array([0.0, 0.3, -1.1, np.pi / 2, np.pi])
    out["rot/angles"] = angs
    for fn in ("rot_x", "rot_y", "rot_z", "der_rot_x", "der_rot_y", "der_rot_z"):
        out["rot/" + fn] = np.array([getattr(rotations, fn)(a) for a in angs])
    # the reference's Shepp-Logan generator; np.lib.index_tricks was made private in numpy 2
    # (SURVEY.md F8), so alias it for the import -- the reference's file is untouched
    if not hasattr(np.lib, "index_tricks"):
        np.lib.index_tricks = np.lib._index_tricks_impl
    from utilities import generate_phantom
    for n in (16, 24):
        out["shepp3d/%d" % n] = generate_phantom.shepp3d(n)
    # TV proximal step: the reference's tv_denoise.py is pure numpy and runs as is
    from utilities import tv_denoise
    tv_im = rng.random((12, 10, 14)).astype(np.float32) + np.linspace(0, 2, 14, dtype=np.float32)[None, None, :]
    out["tv/im"] = tv_im
    out["tv/div_in"] = rng.random((3, 12, 10, 14)).astype(np.float32)
    out["tv/div"] = tv_denoise.div(out["tv/div_in"])
    out["tv/gradient"] = tv_denoise.gradient(tv_im)
    for w, nit in ((0.05, 7), (0.5, 20), (2.0, 60)):
        out["tv/denoise_w%g_n%d" % (w, nit)] = tv_denoise.denoise_fista(tv_im, weight=w, niter=nit)
    out["tv/tv_norm_3d"] = np.array(tv_denoise.tv_norm_3d(tv_im))
    np.savez_compressed(OUT, **out)
    print("wrote", OUT, os.path.getsize(OUT), "bytes")


if __name__ == "__main__":
    main()
