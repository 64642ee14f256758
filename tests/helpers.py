"""Shared test helpers: seeded cases, and two CPU stand-ins for CudaBackend used ONLY by tests:

  OracleBackend  applies the oracle (oracle/oracle.py) -- exercises the host logic of
                 ProjectionMatrix / ProjectionOperator / ShardedProjector without a GPU;
  EmuBackend     runs the __host__ __device__ cores of the CUDA kernels on the CPU
                 (tests/emu/libtomo_emu.so) -- checks the kernel arithmetic against the oracle.
"""
import ctypes
import os

import numpy as np
import torch

from oracle import oracle as O
from tomography_alignment_b200 import Geometry, pose_table
from tomography_alignment_b200.projection_operators import full_pose_table
from tomography_alignment_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64).ravel()
    b = np.asarray(b, dtype=np.float64).ravel()
    return np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300)


def make_geoms(shape, dshape, n_proj, cor=None, step=1.0, vox_pix=None, det_pix=None):
    vox_pix = np.ones(3) if vox_pix is None else np.asarray(vox_pix, dtype=float)
    det_pix = np.ones(2) if det_pix is None else np.asarray(det_pix, dtype=float)
    cor = None if cor is None else np.asarray(cor, dtype=float)
    g = Geometry(n_proj, np.array(shape), vox_pix, np.array(dshape), det_pix, cor_shift=cor, step_size=step)
    og = O.OracleGeometry(n_proj, np.array(shape), vox_pix, np.array(dshape), det_pix, cor_shift=cor, step_size=step)
    return g, og


def random_poses(n_proj, seed, tilt=0.02, shift=2.0, phis=None):
    """Poses in the style of examples/generate_data.py:16-23 (seeded; ty also jittered)."""
    rng = np.random.default_rng(seed)
    phi = np.linspace(0.0, np.pi, n_proj) if phis is None else np.asarray(phis, dtype=float)
    alpha = rng.uniform(-tilt, tilt, n_proj)
    beta = rng.uniform(-tilt, tilt, n_proj)
    xyz = rng.uniform(-shift, shift, (n_proj, 3))
    return phi, alpha, beta, xyz


def _P(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class _HostBackendBase(object):
    def __init__(self, geometry):
        self.geometry = geometry
        self.vol_shape = tuple(int(v) for v in geometry.vox_shape)
        self.det_shape = tuple(int(v) for v in geometry.det_shape)
        self.n_det = self.det_shape[0] * self.det_shape[1]
        self.n_proj = 0
        self.poses = None

    def _as_vol(self, vol):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(vol, dtype=np.float32))).contiguous()

    def _as_proj(self, y):
        return torch.as_tensor(np.ascontiguousarray(np.asarray(y, dtype=np.float32))).contiguous()

    def set_poses(self, poses):
        poses = np.asarray(poses, dtype=np.float64)
        poses = poses.reshape(-1, poses.shape[-1])
        self.poses12 = np.ascontiguousarray(poses) if poses.shape[1] == 12 else None     # caller-chosen sample counts
        self.poses = np.ascontiguousarray(poses[:, :9])
        self.n_proj = self.poses.shape[0]
        self._bound_state = None


class OracleBackend(_HostBackendBase):
    def __init__(self, geometry):
        super().__init__(geometry)
        self.og = O.OracleGeometry(geometry.n_proj, geometry.vox_shape, geometry.vox_pix, geometry.det_shape,
                                   geometry.det_pix, step_size=geometry.step_size)

    def _op(self):
        p = self.poses
        self.og.cor_shift = p[:, 6:9]
        self.og.n_proj = self.n_proj
        return O.OracleOperator(self.og, alpha=p[:, 1], beta=p[:, 2], phi=p[:, 0], xyz_shift=p[:, 3:6])

    def forward(self, vol, out=None):
        y = self._op().forward(np.asarray(vol, dtype=np.float64).ravel())
        return torch.as_tensor(y.astype(np.float32)).reshape((self.n_proj,) + self.det_shape)

    def adjoint(self, y, out=None, accumulate=False):
        v = self._op().adjoint(np.asarray(y, dtype=np.float64).reshape(self.n_proj, -1))
        return torch.as_tensor(v.astype(np.float32)).reshape(self.vol_shape)

    def proj_grad(self, vol, meas=None, want_proj=True, want_dproj=True, want_grad6=None, repad=True):
        p = self.poses
        proj = np.zeros((self.n_proj, self.n_det))
        dproj = np.zeros((self.n_proj, 6, self.n_det))
        for i in range(self.n_proj):
            proj[i], dproj[i] = O.forward_proj_grad(self.og, p[i, 1], p[i, 2], p[i, 0], p[i, 3:6], p[i, 6:9],
                                                    np.asarray(vol, dtype=np.float64))
        out = {"proj": torch.as_tensor(proj.astype(np.float32)).reshape((self.n_proj,) + self.det_shape),
               "dproj": torch.as_tensor(dproj.astype(np.float32)), "grad6": None, "cost": None}
        if meas is not None:
            res = np.asarray(meas, dtype=np.float64).reshape(self.n_proj, -1) - proj
            out["grad6"] = torch.as_tensor(np.einsum("vkr,vr->vk", -dproj, res))
            out["cost"] = torch.as_tensor(0.5 * (res ** 2).sum(axis=1))
        return out


class EmuBackend(_HostBackendBase):
    """Same arithmetic as the CUDA kernels, executed by tests/emu/libtomo_emu.so on the CPU."""

    def __init__(self, geometry, zquad=False):
        super().__init__(geometry)
        self.zquad = zquad          # let qualifying views take the z-quad core (csrc/zq_core.h), as CudaBackend(zquad=True) does
        self.L = _lib.load()
        self.E = ctypes.CDLL(os.path.join(ROOT, "tests", "emu", "libtomo_emu.so"))
        self.cg = geometry.to_c()
        self.views = None

    def set_poses(self, poses):
        super().set_poses(poses)
        self.views = np.zeros((self.n_proj, _lib.VIEW_STRIDE))
        full = self.poses12 if self.poses12 is not None else full_pose_table(self.geometry, self.poses,
                                                                             flags=1.0 if self.zquad else 0.0)
        rc = self.L.tomo_views_compute_host(ctypes.byref(self.cg), _P(full), self.n_proj, _P(self.views))
        _lib.check(rc, "tomo_views_compute_host")

    def _pad(self, vol):
        vol = np.ascontiguousarray(np.asarray(vol, dtype=np.float32).reshape(self.vol_shape))
        pad = np.zeros(self.L.tomo_padded_volume_bytes(ctypes.byref(self.cg)) // 4, np.float32)
        self.E.emu_pad(ctypes.byref(self.cg), _P(vol), _P(pad))
        return pad

    def forward(self, vol, out=None):
        pad = self._pad(vol)
        proj = np.zeros((self.n_proj, self.n_det), np.float32)
        self.E.emu_proj_grad(ctypes.byref(self.cg), _P(self.views), self.n_proj, _P(pad), None, _P(proj), None, None,
                             None, 0)
        return torch.as_tensor(proj).reshape((self.n_proj,) + self.det_shape)

    def adjoint(self, y, out=None, accumulate=False, voxel_bilinear=False, origin=None):
        y = np.ascontiguousarray(np.asarray(y, dtype=np.float32).reshape(self.n_proj, -1))
        vol = np.zeros(self.vol_shape, np.float32)
        org = np.asarray(self.geometry.det_orig if origin is None else origin, dtype=np.float64)
        self.E.emu_back(ctypes.byref(self.cg), _P(self.views), self.n_proj, _P(y), _P(vol), 0,
                        int(voxel_bilinear), _P(org))
        return torch.as_tensor(vol)

    def voxel_back(self, y, origin=None, out=None, accumulate=False):
        return self.adjoint(y, voxel_bilinear=True, origin=origin)

    def proj_grad(self, vol, meas=None, want_proj=True, want_dproj=True, want_grad6=None, repad=True):
        pad = self._pad(vol)
        proj = np.zeros((self.n_proj, self.n_det), np.float32)
        dproj = np.zeros((self.n_proj, 6, self.n_det), np.float32)
        grad6 = np.zeros((self.n_proj, 6))
        cost = np.zeros(self.n_proj)
        m = None if meas is None else np.ascontiguousarray(np.asarray(meas, dtype=np.float32).reshape(self.n_proj, -1))
        self.E.emu_proj_grad(ctypes.byref(self.cg), _P(self.views), self.n_proj, _P(pad),
                             None if m is None else _P(m), _P(proj), _P(dproj), _P(grad6), _P(cost), 1)
        return {"proj": torch.as_tensor(proj).reshape((self.n_proj,) + self.det_shape),
                "dproj": torch.as_tensor(dproj),
                "grad6": torch.as_tensor(grad6) if m is not None else None,
                "cost": torch.as_tensor(cost) if m is not None else None}
