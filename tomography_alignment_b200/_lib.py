"""ctypes binding of libtomo_b200.so (C ABI: include/tomo_b200.h).

There is no CPU fallback: if the library is missing, or a compute entry point is called without a
CUDA device, the call raises.  ``build()`` compiles the library in-tree with nvcc for sm_100a.
"""
import ctypes
import os
import subprocess

from .geometry import TomoGeom

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TOMO_B200_LIB", os.path.join(_HERE, "libtomo_b200.so"))   # override: tuning builds only
VIEW_STRIDE = 160
POSE_STRIDE = 12
PAD = 2

_lib = None


class TomoError(RuntimeError):
    pass


def build(force=False):
    """Compile csrc/*.cu into libtomo_b200.so (nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo)."""
    csrc = os.path.join(_HERE, "csrc")
    if force and os.path.exists(LIB_PATH):
        os.remove(LIB_PATH)
    subprocess.check_call(["make", "-s", "-C", csrc])
    return LIB_PATH


def load():
    """Load the library; raises TomoError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise TomoError("libtomo_b200.so not found at %s: run `python -c 'import __graft_entry__ as g; g.build()'` "
                        "(there is no CPU fallback for the projection operators)" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    vp, sz, ci = ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int
    G = ctypes.POINTER(TomoGeom)
    L.tomo_version.restype = ci
    L.tomo_version.argtypes = []
    L.tomo_last_error.restype = ctypes.c_char_p
    L.tomo_last_error.argtypes = []
    L.tomo_views_compute_host.restype = ci
    L.tomo_views_compute_host.argtypes = [G, vp, ci, vp]
    L.tomo_views_bytes.restype = sz
    L.tomo_views_bytes.argtypes = [ci]
    L.tomo_views_upload.restype = ci
    L.tomo_views_upload.argtypes = [G, vp, ci, vp, vp]
    L.tomo_views_kinds.restype = ci
    L.tomo_views_kinds.argtypes = [vp, ci]
    L.tomo_forward_ex.restype = ci
    L.tomo_forward_ex.argtypes = [G, vp, ci, ci, vp, vp, vp]
    L.tomo_back_adjoint_slab_granularity.restype = ci
    L.tomo_back_adjoint_slab_granularity.argtypes = []
    L.tomo_back_adjoint_slab.restype = ci
    L.tomo_back_adjoint_slab.argtypes = [G, vp, ci, ci, vp, vp, ci, vp, sz, ci, ci, vp]
    L.tomo_proj_grad_ex.restype = ci
    L.tomo_proj_grad_ex.argtypes = [G, vp, ci, ci, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    L.tomo_padded_volume_bytes.restype = sz
    L.tomo_padded_volume_bytes.argtypes = [G]
    L.tomo_pad_volume.restype = ci
    L.tomo_pad_volume.argtypes = [G, vp, vp, vp]
    L.tomo_forward.restype = ci
    L.tomo_forward.argtypes = [G, vp, ci, vp, vp, vp]
    L.tomo_back_adjoint.restype = ci
    L.tomo_back_adjoint.argtypes = [G, vp, ci, vp, vp, ci, vp]
    L.tomo_back_adjoint_workspace_bytes.restype = sz
    L.tomo_back_adjoint_workspace_bytes.argtypes = [G, ci]
    L.tomo_back_adjoint_ws.restype = ci
    L.tomo_back_adjoint_ws.argtypes = [G, vp, ci, vp, vp, ci, vp, sz, vp]
    L.tomo_back_adjoint_gather.restype = ci
    L.tomo_back_adjoint_gather.argtypes = [G, vp, ci, vp, vp, ci, vp]
    L.tomo_back_voxel_bilinear.restype = ci
    L.tomo_back_voxel_bilinear.argtypes = [G, vp, ci, vp, vp, vp, ci, vp]
    L.tomo_voxel_splat.restype = ci
    L.tomo_voxel_splat.argtypes = [G, vp, ci, vp, vp, vp, vp]
    L.tomo_voxel_splat_workspace_bytes.restype = sz
    L.tomo_voxel_splat_workspace_bytes.argtypes = [G, ci, ci]
    L.tomo_voxel_splat_deterministic.restype = ci
    L.tomo_voxel_splat_deterministic.argtypes = [G, vp, ci, vp, vp, vp, vp, sz, vp]
    L.tomo_voxel_splat_adjoint.restype = ci
    L.tomo_voxel_splat_adjoint.argtypes = [G, vp, ci, vp, vp, ci, vp]
    L.tomo_tv_dual_error.restype = ci
    L.tomo_tv_dual_error.argtypes = [ci, ci, ci, ctypes.c_float, vp, vp, vp, vp]
    L.tomo_tv_dual_update.restype = ci
    L.tomo_tv_dual_update.argtypes = [ci, ci, ci, ctypes.c_float, ctypes.c_float, vp, vp, vp, vp]
    L.tomo_proj_grad_workspace_bytes.restype = sz
    L.tomo_proj_grad_workspace_bytes.argtypes = [G, ci]
    L.tomo_proj_grad.restype = ci
    L.tomo_proj_grad.argtypes = [G, vp, ci, vp, vp, vp, vp, vp, vp, vp, sz, vp]
    _lib = L
    return L


def check(code, what):
    """Raise TomoError(tomo_last_error()) for a non-zero ABI return code."""
    if code != 0:
        msg = load().tomo_last_error().decode("utf-8", "replace")
        raise TomoError("%s failed (code %d): %s" % (what, code, msg))
