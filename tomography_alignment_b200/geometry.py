"""Parallel-beam geometry with the attribute names of the reference's ``Geometry``
(utilities/geometry.py:9-105), so code written against the reference keeps working.

Differences that do not change any value:
  * ``vox_centers`` (3, n_vox) float64 costs 24*n_vox bytes (3.2 GB at 512^3) and is only needed by
    the voxel-driven path; it is built lazily on first access instead of in ``__init__``;
  * ``to_c()`` exports the scalars the grids derive from as the ``TomoGeom`` struct of
    include/tomo_b200.h -- the CUDA kernels never read the grids.
"""
import ctypes

import numpy as np


class TomoGeom(ctypes.Structure):
    """ctypes mirror of ``struct TomoGeom`` (include/tomo_b200.h)."""
    _fields_ = [("nx", ctypes.c_int32), ("ny", ctypes.c_int32), ("nz", ctypes.c_int32),
                ("ndx", ctypes.c_int32), ("ndz", ctypes.c_int32),
                ("vox_origin", ctypes.c_double * 3), ("vox_pix", ctypes.c_double * 3),
                ("det_x0", ctypes.c_double), ("det_z0", ctypes.c_double),
                ("det_dx", ctypes.c_double), ("det_dz", ctypes.c_double),
                ("src_y", ctypes.c_double), ("det_y", ctypes.c_double),
                ("step_size", ctypes.c_double)]


class Geometry(object):
    """Detector and object setup for parallel beam geometry (utilities/geometry.py:14-48).

    :param n_proj: int, number of projections
    :param voxel_shape: int, (3,)
    :param voxel_pixsize: float (3,)
    :param detector_shape: int (2,)
    :param detector_pixsize: float (2,)
    :param cor_shift: None, (3,) or (n_proj, 3) centre-of-rotation shift
    :param step_size: ray-marching step
    """

    def __init__(self, n_proj, voxel_shape, voxel_pixsize, detector_shape, detector_pixsize,
                 cor_shift=None, step_size=1.0):
        self.n_proj = n_proj
        self.vox_shape = np.asarray(voxel_shape)
        self.vox_pix = np.asarray(voxel_pixsize)
        self.vox_size = self.vox_shape * self.vox_pix
        self.n_vox = np.prod(self.vox_shape)
        self.det_shape = np.asarray(detector_shape)
        self.det_pix = np.asarray(detector_pixsize)
        self.det_size = self.det_shape * self.det_pix
        self.n_det = np.prod(self.det_shape)
        self.vox_ds = np.array([1, 1, 1])
        if cor_shift is None:
            self.cor_shift = np.zeros((n_proj, 3))
        else:
            cor_shift = np.asarray(cor_shift)
            if len(cor_shift.shape) == 2:
                assert (cor_shift.shape[0] == n_proj)
                assert (cor_shift.shape[1] == 3)
                self.cor_shift = cor_shift
            elif len(cor_shift.shape) == 1:
                assert (np.size(cor_shift) == 3)
                self.cor_shift = np.tile(cor_shift, n_proj).reshape(n_proj, 3)
            else:
                # same message and same (lack of) consequence as utilities/geometry.py:44
                print('shape or size of cor_shift not valid')
        self.step_size = step_size
        self._vox_centers = None
        self._voxel_detector_grid()

    def _voxel_detector_grid(self):
        # voxel centres and origin, utilities/geometry.py:80-87
        nx, ny, nz = self.vox_shape
        sx, sy, sz = self.vox_size
        self._vx = np.linspace(-sx / 2, sx / 2, nx, endpoint=False) + 0.5
        self._vy = np.linspace(-sy / 2, sy / 2, ny, endpoint=False) + 0.5
        self._vz = np.linspace(-sz / 2, sz / 2, nz, endpoint=False) + 0.5
        self.vox_origin = np.array([self._vx.min(), self._vy.min(), self._vz.min()])
        # detector grid, utilities/geometry.py:89-100
        ndx, ndz = self.det_shape
        dsx, dsz = self.det_size
        self._dx = np.linspace(-dsx / 2, dsx / 2, ndx, endpoint=False) + 0.5
        self._dz = np.linspace(-dsz / 2, dsz / 2, ndz, endpoint=False) + 0.5
        xd, zd = np.meshgrid(self._dx, self._dz, indexing='ij')
        self._y_source = -sy
        self._y_det = sy
        self.source_centers = np.array([xd.ravel(), self._y_source * np.ones((self.n_det,)), zd.ravel()])
        self.det_centers = np.array([xd.ravel(), self._y_det * np.ones((self.n_det,)), zd.ravel()])
        # voxel-based method, utilities/geometry.py:102-105
        self.det_orig = np.array([self._dx.min(), self._vy.min(), self._dz.min()])
        fx, fz = float(self.vox_shape[0] / self.det_shape[0]), float(self.vox_shape[2] / self.det_shape[1])
        self.factor = np.array([fx, 1., fz])

    @property
    def vox_centers(self):
        if self._vox_centers is None:
            x, y, z = np.meshgrid(self._vx, self._vy, self._vz, indexing='ij')
            self._vox_centers = np.array([x.ravel(), y.ravel(), z.ravel()])
        return self._vox_centers

    def to_c(self):
        """The ``TomoGeom`` the C ABI takes.  Grid origins and pitches are read back from the grids
        built above, so any value the reference would compute is reproduced bit for bit."""
        g = TomoGeom()
        g.nx, g.ny, g.nz = (int(v) for v in self.vox_shape)
        g.ndx, g.ndz = (int(v) for v in self.det_shape)
        for a in range(3):
            g.vox_origin[a] = float(self.vox_origin[a])
            g.vox_pix[a] = float(self.vox_pix[a])
        g.det_x0, g.det_z0 = float(self._dx[0]), float(self._dz[0])
        g.det_dx = float(self.det_pix[0])
        g.det_dz = float(self.det_pix[1])
        g.src_y, g.det_y = float(self._y_source), float(self._y_det)
        g.step_size = float(self.step_size)
        return g
