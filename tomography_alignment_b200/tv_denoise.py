"""Total-variation proximal operator with the interface of the reference's ``utilities/tv_denoise.py``
(SURVEY.md section 8f, row N3).

``denoise_fista(im, weight, niter, eps, check_gap_frequency)`` solves argmin_res 1/2 |im - res|^2 + weight * TV(res)
with the dual FISTA of Beck & Teboulle exactly as the reference iterates it (tv_denoise.py:98-170): same
step 1/(factor * weight) with factor 12 in 3-D, same momentum, same dual-gap stopping test every
``check_gap_frequency`` iterations, and -- like the reference -- the returned image is the one of the last gap
check.  The per-iteration work runs in two fused CUDA kernels (csrc/tv_kernels.cu); the infrequent dual-gap
reductions use torch.  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out.
"""
import ctypes

import numpy as np
import torch

from . import _lib


class CudaTvOps(object):
    """The two dual-iteration kernels behind the C ABI."""

    def __init__(self, device):
        if not torch.cuda.is_available():
            raise _lib.TomoError("tv_denoise needs a CUDA device (there is no CPU fallback)")
        self.lib = _lib.load()
        self.device = torch.device(device)

    def _s(self):
        return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def dual_error(self, weight, p, im, err):
        nx, ny, nz = im.shape
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_tv_dual_error(nx, ny, nz, float(weight), ctypes.c_void_p(p.data_ptr()),
                                             ctypes.c_void_p(im.data_ptr()), ctypes.c_void_p(err.data_ptr()), self._s())
        _lib.check(rc, "tomo_tv_dual_error")

    def dual_update(self, inv_fw, t_factor, err, aux, gim):
        nx, ny, nz = err.shape
        with torch.cuda.device(self.device):
            rc = self.lib.tomo_tv_dual_update(nx, ny, nz, float(inv_fw), float(t_factor), ctypes.c_void_p(err.data_ptr()),
                                              ctypes.c_void_p(aux.data_ptr()), ctypes.c_void_p(gim.data_ptr()), self._s())
        _lib.check(rc, "tomo_tv_dual_update")


def _fwd_diff_norm(new):
    """sum sqrt(gx^2 + gy^2 + gz^2) of the zero-terminated forward differences (dual_gap, tv_denoise.py:84-92)."""
    g2 = torch.zeros_like(new, dtype=torch.float64)
    for d in range(3):
        df = torch.diff(new, dim=d).double() ** 2
        sl = [slice(None)] * 3
        sl[d] = slice(0, -1)
        g2[tuple(sl)] += df
    return torch.sqrt(g2).sum()


def dual_gap(im, new, gap, weight):
    """tv_denoise.py:77-95 (3-D), float64 accumulation."""
    im_norm = (im.double() ** 2).sum()
    tv_new = 2 * weight * _fwd_diff_norm(new)
    dg = (gap.double() ** 2).sum() + tv_new - im_norm + (new.double() ** 2).sum()
    return float(0.5 / im_norm * dg)


def tv_norm_3d(x):
    """tv_denoise.py:62-64: l2 norm of the forward-difference gradient field."""
    x = torch.as_tensor(x)
    return float(torch.sqrt(sum((torch.diff(x, dim=d).double() ** 2).sum() for d in range(3))))


def denoise_fista(im, weight=50, niter=200, eps=1.e-5, check_gap_frequency=3, ops=None):
    """tv_denoise.py:98-170 for 3-D float32 volumes."""
    was_numpy = not isinstance(im, torch.Tensor)
    if ops is None:
        dev = im.device if (not was_numpy and im.is_cuda) else torch.device("cuda", torch.cuda.current_device()) \
            if torch.cuda.is_available() else None
        ops = CudaTvOps(dev)
    dev = ops.device
    imd = torch.as_tensor(np.ascontiguousarray(im, dtype=np.float32) if was_numpy else im).to(dev, torch.float32).contiguous()
    if imd.ndim != 3:
        raise ValueError("denoise_fista: the B200 kernels handle 3-D volumes (got %d-D)" % imd.ndim)
    factor = 12.0
    grad_im = torch.zeros((3,) + tuple(imd.shape), dtype=torch.float32, device=dev)
    grad_aux = torch.zeros_like(grad_im)
    err = torch.empty_like(imd)
    t = 1.
    i = 0
    new = imd.clone()
    while i < niter:
        ops.dual_error(weight, grad_aux, imd, err)                    # error = weight * div(grad_aux) - im
        t_new = 0.5 * (1 + np.sqrt(1 + 4 * t ** 2))
        t_factor = (t - 1) / t_new
        ops.dual_update(1.0 / (factor * weight), t_factor, err, grad_aux, grad_im)
        t = t_new
        if (i % check_gap_frequency) == 0:
            ops.dual_error(weight, grad_im, imd, err)                 # err = gap - im  ->  new = -err, gap = err + im
            new = -err
            gap = err + imd
            if dual_gap(imd, new, gap, weight) < eps:
                break
        i += 1
    return new.cpu().numpy() if was_numpy else new
