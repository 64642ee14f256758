"""Device side of the operators: torch tensors own the HBM buffers, libtomo_b200.so runs the
sm_100a kernels on torch's current stream.  torch is plumbing here (memory, streams,
torch.distributed); all arithmetic of the hot path happens inside the C ABI.

No CPU fallback: constructing a CudaBackend without a CUDA device raises.
"""
import ctypes

import numpy as np
import torch

from . import _lib


def _ptr(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


class CudaBackend(object):
    """Owns the per-view table, the zero-bordered volume copy and the reduction workspace for one
    geometry on one GPU.  Volumes are float32 [nx, ny, nz] (z fastest), projections float32
    [n_proj, ndx, ndz] (iz fastest): the reference's layouts (include/tomo_b200.h)."""

    def __init__(self, geometry, device=None):
        if not torch.cuda.is_available():
            raise _lib.TomoError("tomography_alignment_b200 needs a CUDA device (B200, sm_100a); "
                                 "there is no CPU fallback for the projection operators")
        self.lib = _lib.load()
        self.geometry = geometry
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cgeom = geometry.to_c()
        self.vol_shape = tuple(int(v) for v in geometry.vox_shape)
        self.det_shape = tuple(int(v) for v in geometry.det_shape)
        self.n_det = self.det_shape[0] * self.det_shape[1]
        self.n_proj = 0
        self.views = None
        self._volpad = None
        self._ws = None
        self.launches = 0        # kernels of ours launched so far (bench.py reports the count)

    # -- helpers -------------------------------------------------------------------------------
    def _stream(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def _g(self):
        return ctypes.byref(self.cgeom)

    def _as_vol(self, vol):
        vol = torch.as_tensor(vol)
        if vol.device != self.device or vol.dtype != torch.float32:
            vol = vol.to(device=self.device, dtype=torch.float32, non_blocking=True)
        vol = vol.contiguous()
        if vol.numel() != int(np.prod(self.vol_shape)):
            raise ValueError("volume has %d elements, geometry expects %d" % (vol.numel(), int(np.prod(self.vol_shape))))
        return vol

    def _as_proj(self, y):
        y = torch.as_tensor(y)
        if y.device != self.device or y.dtype != torch.float32:
            y = y.to(device=self.device, dtype=torch.float32, non_blocking=True)
        y = y.contiguous()
        if y.numel() != self.n_proj * self.n_det:
            raise ValueError("projections have %d elements, operator expects %d" % (y.numel(), self.n_proj * self.n_det))
        return y

    # -- poses ---------------------------------------------------------------------------------
    def set_poses(self, poses):
        """poses: float64 (n_proj, 9) = phi, alpha, beta, tx, ty, tz, cor_x, cor_y, cor_z."""
        poses = np.ascontiguousarray(np.asarray(poses, dtype=np.float64).reshape(-1, _lib.POSE_STRIDE))
        n = poses.shape[0]
        if self.views is None or self.views.shape[0] != n:
            self.views = torch.empty((n, _lib.VIEW_STRIDE), dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_views_upload(self._g(), poses.ctypes.data_as(ctypes.c_void_p), n, _ptr(self.views),
                                            self._stream())
        _lib.check(rc, "tomo_views_upload")
        self.n_proj = n

    # -- operators -----------------------------------------------------------------------------
    def pad(self, vol):
        """Zero-bordered copy of the volume the ray-driven kernels read (tomo_pad_volume)."""
        vol = self._as_vol(vol)
        nbytes = self.lib.tomo_padded_volume_bytes(self._g())
        if self._volpad is None or self._volpad.numel() * 4 != nbytes:
            self._volpad = torch.empty(nbytes // 4, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_pad_volume(self._g(), _ptr(vol), _ptr(self._volpad), self._stream())
        _lib.check(rc, "tomo_pad_volume")
        self.launches += 1
        return self._volpad

    def forward(self, vol, out=None):
        """proj = A vol: float32 (n_proj, ndx, ndz) on the device."""
        volpad = self.pad(vol)
        if out is None:
            out = torch.empty((self.n_proj,) + self.det_shape, dtype=torch.float32, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_forward(self._g(), _ptr(self.views), self.n_proj, _ptr(volpad), _ptr(out), self._stream())
        _lib.check(rc, "tomo_forward")
        self.launches += 1
        return out

    def adjoint(self, y, out=None, accumulate=False, gather=False):
        """vol (+)= A^T y: float32 (nx, ny, nz) on the device.  ``gather=True`` uses the per-voxel gather
        kernel (tomo_back_adjoint_gather), the independent formulation kept as a cross-check."""
        y = self._as_proj(y)
        if out is None:
            out = torch.empty(self.vol_shape, dtype=torch.float32, device=self.device)
            accumulate = False
        fn = self.lib.tomo_back_adjoint_gather if gather else self.lib.tomo_back_adjoint
        with torch.cuda.device(self.device):
            rc = fn(self._g(), _ptr(self.views), self.n_proj, _ptr(y), _ptr(out), int(bool(accumulate)), self._stream())
        _lib.check(rc, "tomo_back_adjoint")
        self.launches += 1
        return out

    def voxel_back(self, y, origin=None, out=None, accumulate=False):
        """Orphan voxel-driven bilinear backprojector (src/back_projection.f90)."""
        y = self._as_proj(y)
        origin = np.asarray(self.geometry.det_orig if origin is None else origin, dtype=np.float64)
        org = (ctypes.c_double * 3)(*[float(v) for v in origin])
        if out is None:
            out = torch.empty(self.vol_shape, dtype=torch.float32, device=self.device)
            accumulate = False
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_back_voxel_bilinear(self._g(), _ptr(self.views), self.n_proj, org, _ptr(y), _ptr(out),
                                                   int(bool(accumulate)), self._stream())
        _lib.check(rc, "tomo_back_voxel_bilinear")
        self.launches += 1
        return out

    def proj_grad(self, vol, meas=None, want_proj=True, want_dproj=True, want_grad6=None, repad=True):
        """Projection + 6-DOF gradient for all current views (tomo_proj_grad).

        Returns dict(proj (n_proj, ndx, ndz) f32, dproj (n_proj, 6, n_det) f32,
                     grad6 (n_proj, 6) f64, cost (n_proj,) f64); entries not requested are None."""
        volpad = self.pad(vol) if repad or self._volpad is None else self._volpad
        if want_grad6 is None:
            want_grad6 = meas is not None
        n = self.n_proj
        proj = torch.empty((n,) + self.det_shape, dtype=torch.float32, device=self.device) if want_proj else None
        dproj = torch.empty((n, 6, self.n_det), dtype=torch.float32, device=self.device) if want_dproj else None
        grad6 = cost = None
        ws_bytes = 0
        if meas is not None:
            meas = self._as_proj(meas)
        if want_grad6:
            if meas is None:
                raise ValueError("grad6 needs the measured projections")
            grad6 = torch.empty((n, 6), dtype=torch.float64, device=self.device)
            cost = torch.empty((n,), dtype=torch.float64, device=self.device)
            ws_bytes = self.lib.tomo_proj_grad_workspace_bytes(self._g(), n)
            if self._ws is None or self._ws.numel() * 8 < ws_bytes:
                self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_proj_grad(self._g(), _ptr(self.views), n, _ptr(volpad), _ptr(meas), _ptr(proj), _ptr(dproj),
                                         _ptr(grad6), _ptr(cost), _ptr(self._ws) if want_grad6 else None,
                                         ws_bytes, self._stream())
        _lib.check(rc, "tomo_proj_grad")
        self.launches += 2 if want_grad6 else 1
        return {"proj": proj, "dproj": dproj, "grad6": grad6, "cost": cost}
