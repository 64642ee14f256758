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

    def __init__(self, geometry, device=None, zquad=False):
        if not torch.cuda.is_available():
            raise _lib.TomoError("tomography_alignment_b200 needs a CUDA device (B200, sm_100a); "
                                 "there is no CPU fallback for the projection operators")
        self.lib = _lib.load()
        self.geometry = geometry
        # zquad=True: nearly untilted views take the z-quad forward / gradient kernels (csrc/zq_core.h: four z-adjacent rays
        # per thread, 128-bit loads).  Same results to rounding; measured slower than the per-ray kernels on B200, hence off.
        self.zquad = bool(zquad)
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cgeom = geometry.to_c()
        self.vol_shape = tuple(int(v) for v in geometry.vox_shape)
        self.det_shape = tuple(int(v) for v in geometry.det_shape)
        self.n_det = self.det_shape[0] * self.det_shape[1]
        self.n_proj = 0
        self.kinds = 0           # TOMO_KINDS_* mask of the current view table (0: unknown)
        self.views = None
        self._volpad = None
        self._ws = None
        self.launches = 0        # kernels of ours launched so far (bench.py reports the count)
        self._bound_state = None  # identity of the ProjectionOperator state whose poses are current
        self._copy_out = None    # stream of queued result downloads (adjoint_host(wait=False))
        self._copy = None        # side stream for host<->device copies overlapped with the kernels
        self._dbuf = {}          # cached device staging buffers of the host-buffer entry points
        self.h2d_bytes = 0       # bytes moved by the host-buffer entry points (bench.py reports them)
        self.d2h_bytes = 0

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

    def _check_out(self, out, shape, what):
        """``out=`` buffers are written by raw pointer: refuse anything the kernels cannot address."""
        n = int(np.prod(shape))
        if not isinstance(out, torch.Tensor) or out.device != self.device or out.dtype != torch.float32 \
                or not out.is_contiguous() or out.numel() != n:
            raise ValueError("%s: out= must be a contiguous float32 tensor of %d elements on %s"
                             % (what, n, self.device))
        return out

    def _views_at(self, first):
        """Device pointer of view record ``first`` (sub-tables are valid inputs of every operator)."""
        return ctypes.c_void_p(self.views.data_ptr() + first * _lib.VIEW_STRIDE * 8)

    def _buf(self, name, shape, dtype=torch.float32):
        t = self._dbuf.get(name)
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != dtype:
            t = torch.empty(shape, dtype=dtype, device=self.device)
            self._dbuf[name] = t
        return t

    def _back_ws(self, n_views):
        """Workspace of the separable adjoint (z-transposed projections of the untilted views)."""
        nbytes = self.lib.tomo_back_adjoint_workspace_bytes(self._g(), n_views)
        t = self._dbuf.get("back_ws")
        if t is None or t.numel() * 4 < nbytes:
            t = torch.empty((nbytes + 3) // 4, dtype=torch.float32, device=self.device)
            self._dbuf["back_ws"] = t
        return t, nbytes

    def _copy_stream(self):
        if self._copy is None:
            self._copy = torch.cuda.Stream(device=self.device)
        return self._copy

    @staticmethod
    def _host(x, dtype=torch.float32):
        """numpy array or CPU tensor -> contiguous CPU tensor of ``dtype`` (no copy when already so)."""
        t = torch.as_tensor(x)
        if t.dtype != dtype:
            t = t.to(dtype)
        return t.contiguous()

    def _chunks(self, chunk_views):
        # default: up to 8 chunks, but no chunk below 32 views (every launch pays its wave tail and, for the adjoint, one
        # read-modify-write pass over the volume)
        c = max(1, int(chunk_views) if chunk_views else max(32, (self.n_proj + 7) // 8))
        return [(a, min(self.n_proj, a + c)) for a in range(0, self.n_proj, c)]

    # -- poses ---------------------------------------------------------------------------------
    def set_poses(self, poses):
        """poses: float64 (n_proj, 9) = phi, alpha, beta, tx, ty, tz, cor_x, cor_y, cor_z (``pose_table``); the
        reference's per-view sample count is appended here (``full_pose_table``)."""
        from .projection_operators import full_pose_table
        poses = full_pose_table(self.geometry, poses, flags=1.0 if self.zquad else 0.0)
        n = poses.shape[0]
        # always a fresh table: operators returned by earlier projection_matrix() calls keep (and re-bind) theirs
        host = np.empty((n, _lib.VIEW_STRIDE), dtype=np.float64)
        rc = self.lib.tomo_views_compute_host(self._g(), poses.ctypes.data_as(ctypes.c_void_p), n,
                                              host.ctypes.data_as(ctypes.c_void_p))
        _lib.check(rc, "tomo_views_compute_host")
        # which kernel families this table needs: the operators then skip the launches that would find no view
        self.kinds = int(self.lib.tomo_views_kinds(host.ctypes.data_as(ctypes.c_void_p), n))
        self.views = torch.as_tensor(host).to(self.device)
        self._bound_state = None
        self.n_proj = n

    def bind_views(self, views, n_proj, kinds=0):
        """Make a view table uploaded by an earlier set_poses() the current one again (ProjectionOperator._bind)."""
        if views.device != self.device or views.dtype != torch.float64 or tuple(views.shape) != (n_proj, _lib.VIEW_STRIDE):
            raise ValueError("bind_views: not a view table of this backend")
        self.views = views
        self.n_proj = int(n_proj)
        self.kinds = int(kinds)

    def _count(self, op):
        """Kernels one call of ``op`` launches for the current table (bench.py reports the total)."""
        k = self.kinds
        gen, sep, zq = bool(k & 2) or not k, bool(k & 4) or not k, bool(k & 32) or not k
        if op == "forward":
            return int(gen) + int(sep) + int(zq)
        if op == "adjoint":
            tile = bool(k & 8) or not k
            return int(tile) + int(bool(k & 16) or not k) + 2 * int(sep)
        if op == "grad":
            return int(gen) + int(sep) + int(zq)
        return (int(gen) + int(sep) + int(zq)) * 2   # grad + finalize passes

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
        else:
            self._check_out(out, (self.n_proj,) + self.det_shape, "forward")
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_forward_ex(self._g(), _ptr(self.views), self.n_proj, self.kinds, _ptr(volpad), _ptr(out),
                                          self._stream())
        _lib.check(rc, "tomo_forward")
        self.launches += self._count("forward")
        return out

    def adjoint(self, y, out=None, accumulate=False, gather=False, x_range=None):
        """vol (+)= A^T y: float32 (nx, ny, nz) on the device.  ``gather=True`` uses the per-voxel gather
        kernel (tomo_back_adjoint_gather), the independent formulation kept as a cross-check.
        ``x_range=(x0, x1)`` writes only the x-slab [x0, x1) of ``out`` (still the full-volume tensor; multiples of
        ``slab_granularity()``, slabs issued in ascending order from 0): the caller can all-reduce a finished slab
        while the next one is computed."""
        y = self._as_proj(y)
        if out is None:
            out = torch.empty(self.vol_shape, dtype=torch.float32, device=self.device)
            accumulate = False
        else:
            self._check_out(out, self.vol_shape, "adjoint")
        with torch.cuda.device(self.device):
            if gather:
                rc = self.lib.tomo_back_adjoint_gather(self._g(), _ptr(self.views), self.n_proj, _ptr(y), _ptr(out),
                                                       int(bool(accumulate)), self._stream())
            else:
                ws, nbytes = self._back_ws(self.n_proj)
                x0, x1 = (0, self.vol_shape[0]) if x_range is None else (int(x_range[0]), int(x_range[1]))
                rc = self.lib.tomo_back_adjoint_slab(self._g(), _ptr(self.views), self.n_proj, self.kinds, _ptr(y), _ptr(out),
                                                     int(bool(accumulate)), _ptr(ws), nbytes, x0, x1, self._stream())
        _lib.check(rc, "tomo_back_adjoint")
        self.launches += 1 if gather else self._count("adjoint")
        return out

    def slab_streams(self):
        """Two side streams the slab launches of the overlapped multi-GPU backprojection alternate between."""
        if getattr(self, "_slab_pool", None) is None:
            self._slab_pool = [torch.cuda.Stream(device=self.device), torch.cuda.Stream(device=self.device)]
        return self._slab_pool

    def slab_granularity(self):
        return int(self.lib.tomo_back_adjoint_slab_granularity())

    def slabs(self, n_slabs):
        """``n_slabs`` x-ranges of whole tile rows covering the volume (fewer when the volume has fewer tile rows)."""
        gx, nx = self.slab_granularity(), self.vol_shape[0]
        rows = (nx + gx - 1) // gx
        cuts = sorted(set(int(round(k * rows / float(max(1, n_slabs)))) for k in range(max(1, n_slabs) + 1)))
        return [(a * gx, min(nx, b * gx)) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]

    # -- host-buffer entry points: copies overlapped with the kernels in view chunks ------------------
    def forward_host(self, x_host, out_host=None, chunk_views=None, vol_dev=None):
        """proj = A x with HOST input and output (numpy or CPU tensors; pinned memory makes the copies
        asynchronous).  The volume goes up once; the views are projected in chunks and each chunk's
        device->host copy runs on a side stream under the next chunk's kernel.
        ``vol_dev``: the volume is already on this device (e.g. broadcast from the rank that uploaded it)."""
        if vol_dev is None:
            x_host = self._host(x_host).reshape(-1)
            if x_host.numel() != int(np.prod(self.vol_shape)):
                raise ValueError("volume has %d elements, geometry expects %d" % (x_host.numel(), int(np.prod(self.vol_shape))))
        if out_host is None:
            out_host = torch.empty((self.n_proj,) + self.det_shape, dtype=torch.float32, pin_memory=True)
        ret = out_host
        out_host = torch.as_tensor(out_host)                 # numpy arrays are wrapped, not copied
        if out_host.dtype != torch.float32 or not out_host.is_contiguous() or out_host.numel() != self.n_proj * self.n_det \
                or out_host.device.type != "cpu":
            raise ValueError("forward_host: out_host must be a contiguous float32 host buffer of %d elements"
                             % (self.n_proj * self.n_det))
        out3 = out_host.reshape((self.n_proj,) + self.det_shape)
        cur, cp = torch.cuda.current_stream(self.device), self._copy_stream()
        if vol_dev is None:
            vol_d = self._buf("vol", self.vol_shape)
            vol_d.reshape(-1).copy_(x_host, non_blocking=True)
            self.h2d_bytes += 4 * x_host.numel()
        else:
            vol_d = self._as_vol(vol_dev)
        volpad = self.pad(vol_d)
        proj_d = self._buf("proj", (self.n_proj,) + self.det_shape)
        with torch.cuda.device(self.device):
            for a, b in self._chunks(chunk_views):
                rc = self.lib.tomo_forward_ex(self._g(), self._views_at(a), b - a, self.kinds, _ptr(volpad), _ptr(proj_d[a:b]),
                                              self._stream())
                _lib.check(rc, "tomo_forward")
                self.launches += self._count("forward")
                ev = torch.cuda.Event()
                ev.record(cur)
                with torch.cuda.stream(cp):
                    cp.wait_event(ev)
                    out3[a:b].copy_(proj_d[a:b], non_blocking=True)
        cp.synchronize()
        cur.wait_stream(cp)
        self.d2h_bytes += 4 * out3.numel()
        return ret

    def sync_host(self):
        """Wait until every copy queued by the ``*_host`` entry points has landed (for calls made with ``wait=False``)."""
        self._copy_stream().synchronize()
        if self._copy_out is not None:
            self._copy_out.synchronize()
        torch.cuda.current_stream(self.device).synchronize()

    def adjoint_host(self, y_host, out_host=None, chunk_views=None, to_host=True, wait=True):
        """vol = A^T y with HOST input and output: projection chunks go up on a side stream while the
        previous chunk is backprojected (accumulating launches), the volume comes down once.
        ``to_host=False`` returns the device volume instead (callers that all-reduce before the download): it is
        this backend's reusable staging buffer, valid until the next ``*_host`` call on the backend -- clone it to
        keep it.  Either way the call returns only after the uploads of ``y_host`` have completed, so the host
        buffer may be reused at once.  ``wait=False`` queues the download of the volume on the copy stream and returns
        without waiting for it (``out_host`` is valid after ``sync_host()``): the next operator's kernels overlap it."""
        y_host = self._host(y_host)
        if y_host.numel() != self.n_proj * self.n_det:
            raise ValueError("projections have %d elements, operator expects %d" % (y_host.numel(), self.n_proj * self.n_det))
        y_host = y_host.reshape(self.n_proj, -1)
        if out_host is None and to_host:
            out_host = torch.empty(self.vol_shape, dtype=torch.float32, pin_memory=True)
        cur, cp = torch.cuda.current_stream(self.device), self._copy_stream()
        y_d = self._buf("proj", (self.n_proj,) + self.det_shape).reshape(self.n_proj, -1)
        vol_d = self._buf("bp", self.vol_shape)          # not "vol": the volume forward_host uploaded stays valid
        chunks = self._chunks(chunk_views)
        ws, ws_bytes = self._back_ws(max(b - a for a, b in chunks))
        cp.wait_stream(cur)                      # y_d / vol_d may still be in use by earlier work on `cur`
        evs = []
        with torch.cuda.stream(cp):
            for a, b in chunks:
                y_d[a:b].copy_(y_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                evs.append(ev)
        with torch.cuda.device(self.device):
            for k, (a, b) in enumerate(chunks):
                cur.wait_event(evs[k])
                rc = self.lib.tomo_back_adjoint_slab(self._g(), self._views_at(a), b - a, self.kinds, _ptr(y_d[a:b]), _ptr(vol_d),
                                                     int(k > 0), _ptr(ws), ws_bytes, 0, self.vol_shape[0], self._stream())
                _lib.check(rc, "tomo_back_adjoint")
                self.launches += self._count("adjoint")
        self.h2d_bytes += 4 * y_host.numel()
        if not to_host:
            cp.synchronize()                     # the pinned input may be reused by the caller from here on
            return vol_d
        self.d2h_bytes += 4 * out_host.numel()
        if not wait:
            # a stream of its own: on the upload stream the copy would sit in front of the next operator's input chunks
            if self._copy_out is None:
                self._copy_out = torch.cuda.Stream(device=self.device)
            done = torch.cuda.Event()
            done.record(cur)
            with torch.cuda.stream(self._copy_out):
                self._copy_out.wait_event(done)
                out_host.reshape(-1).copy_(vol_d.reshape(-1), non_blocking=True)
            return out_host
        out_host.reshape(-1).copy_(vol_d.reshape(-1), non_blocking=True)
        cur.synchronize()
        return out_host

    def proj_grad_host(self, vol_host, meas_host, chunk_views=None, vol_dev=None, to_host=True):
        """Fused residual gradients with HOST inputs: returns (grad6 (n_proj, 6), cost (n_proj,)) float64 CPU
        tensors (device tensors with ``to_host=False``).  The measured projections go up in chunks under the
        previous chunk's kernel.  ``vol_dev``: the volume is already on this device.  With ``to_host=False`` the
        returned tensors are this backend's reusable buffers (valid until the next ``proj_grad_host`` call)."""
        meas_host = self._host(meas_host)
        if vol_dev is None:
            vol_host = self._host(vol_host).reshape(-1)
        if (vol_dev is None and vol_host.numel() != int(np.prod(self.vol_shape))) or meas_host.numel() != self.n_proj * self.n_det:
            raise ValueError("volume / measured projections do not match the geometry and the current poses")
        meas_host = meas_host.reshape(self.n_proj, -1)
        cur, cp = torch.cuda.current_stream(self.device), self._copy_stream()
        vol_d = self._buf("vol", self.vol_shape)
        m_d = self._buf("proj", (self.n_proj,) + self.det_shape).reshape(self.n_proj, -1)
        grad6 = self._buf("grad6", (self.n_proj, 6), torch.float64)
        cost = self._buf("cost", (self.n_proj,), torch.float64)
        chunks = self._chunks(chunk_views)
        cp.wait_stream(cur)
        evs = []
        with torch.cuda.stream(cp):
            for a, b in chunks:
                m_d[a:b].copy_(meas_host[a:b], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(cp)
                evs.append(ev)
        if vol_dev is None:
            vol_d.reshape(-1).copy_(vol_host, non_blocking=True)
            self.h2d_bytes += 4 * vol_host.numel()
        else:
            vol_d = self._as_vol(vol_dev)
        volpad = self.pad(vol_d)
        cmax = max(b - a for a, b in chunks)
        ws_bytes = self.lib.tomo_proj_grad_workspace_bytes(self._g(), cmax)
        if self._ws is None or self._ws.numel() * 8 < ws_bytes:
            self._ws = torch.empty((ws_bytes + 7) // 8, dtype=torch.float64, device=self.device)
        with torch.cuda.device(self.device):
            for k, (a, b) in enumerate(chunks):
                cur.wait_event(evs[k])
                rc = self.lib.tomo_proj_grad_ex(self._g(), self._views_at(a), b - a, self.kinds, _ptr(volpad), _ptr(m_d[a:b]),
                                                None, None, _ptr(grad6[a:b]), _ptr(cost[a:b]), _ptr(self._ws), ws_bytes,
                                                self._stream())
                _lib.check(rc, "tomo_proj_grad")
                self.launches += self._count("grad6")
        self.h2d_bytes += 4 * meas_host.numel()
        if not to_host:
            cp.synchronize()                     # uploads of the pinned inputs are complete
            if vol_dev is None:
                cur.synchronize()                # ... including the volume, copied on the compute stream
            return grad6, cost
        g6, c = grad6.cpu(), cost.cpu()          # synchronises `cur`
        self.d2h_bytes += 8 * (g6.numel() + c.numel())
        return g6, c

    def voxel_back(self, y, origin=None, out=None, accumulate=False):
        """Orphan voxel-driven bilinear backprojector (src/back_projection.f90)."""
        y = self._as_proj(y)
        origin = np.asarray(self.geometry.det_orig if origin is None else origin, dtype=np.float64)
        org = (ctypes.c_double * 3)(*[float(v) for v in origin])
        if out is None:
            out = torch.empty(self.vol_shape, dtype=torch.float32, device=self.device)
            accumulate = False
        else:
            self._check_out(out, self.vol_shape, "voxel_back")
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_back_voxel_bilinear(self._g(), _ptr(self.views), self.n_proj, org, _ptr(y), _ptr(out),
                                                   int(bool(accumulate)), self._stream())
        _lib.check(rc, "tomo_back_voxel_bilinear")
        self.launches += 1
        return out

    def voxel_splat(self, vol, want_grad=True, deterministic=False):
        """Orphan voxel-driven forward splat (+ gradient image), src/vox_wt_grad.f90:1-55: returns
        (det (n_proj, ndz, ndx), grad (n_proj, 6, ndz, ndx) or None), x fastest.  ``deterministic=True`` accumulates in
        64-bit fixed point (tomo_voxel_splat_deterministic): bitwise reproducible, at the price of an integer workspace."""
        vol = self._as_vol(vol)
        ndx, ndz = self.det_shape
        det = torch.empty((self.n_proj, ndz, ndx), dtype=torch.float32, device=self.device)
        grad = torch.empty((self.n_proj, 6, ndz, ndx), dtype=torch.float32, device=self.device) if want_grad else None
        with torch.cuda.device(self.device):
            if deterministic:
                nbytes = self.lib.tomo_voxel_splat_workspace_bytes(self._g(), self.n_proj, int(bool(want_grad)))
                ws = torch.empty((nbytes + 7) // 8, dtype=torch.int64, device=self.device)
                rc = self.lib.tomo_voxel_splat_deterministic(self._g(), _ptr(self.views), self.n_proj, _ptr(vol), _ptr(det),
                                                             _ptr(grad), _ptr(ws), nbytes, self._stream())
                self.launches += 2
            else:
                rc = self.lib.tomo_voxel_splat(self._g(), _ptr(self.views), self.n_proj, _ptr(vol), _ptr(det), _ptr(grad),
                                               self._stream())
        _lib.check(rc, "tomo_voxel_splat")
        self.launches += 1
        return det, grad

    def voxel_splat_adjoint(self, det, out=None, accumulate=False):
        """vol (+)= S^T det for the splat matrix S of bilinear_sparse (src/vox_wt_grad.f90:58-112); det (n_proj, ndz, ndx)."""
        det = self._as_proj(det)
        if out is None:
            out = torch.empty(self.vol_shape, dtype=torch.float32, device=self.device)
            accumulate = False
        else:
            self._check_out(out, self.vol_shape, "voxel_splat_adjoint")
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_voxel_splat_adjoint(self._g(), _ptr(self.views), self.n_proj, _ptr(det), _ptr(out),
                                                   int(bool(accumulate)), self._stream())
        _lib.check(rc, "tomo_voxel_splat_adjoint")
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
            rc = self.lib.tomo_proj_grad_ex(self._g(), _ptr(self.views), n, self.kinds, _ptr(volpad), _ptr(meas), _ptr(proj),
                                            _ptr(dproj), _ptr(grad6), _ptr(cost), _ptr(self._ws) if want_grad6 else None,
                                            ws_bytes, self._stream())
        _lib.check(rc, "tomo_proj_grad")
        self.launches += self._count("grad6" if want_grad6 else "grad")
        return {"proj": proj, "dproj": dproj, "grad6": grad6, "cost": cost}
