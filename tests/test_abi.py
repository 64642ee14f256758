"""The C-ABI library loads and exports every symbol include/tomo_b200.h declares (no compute calls)."""
import ctypes
import os
import re

from tomography_alignment_b200 import _lib
from tomography_alignment_b200.geometry import TomoGeom

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tomo_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"TOMO_API\s+[\w\s\*]+?\b(tomo_\w+)\s*\(", src)))


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["tomo_version", "tomo_last_error", "tomo_views_compute_host", "tomo_views_bytes", "tomo_views_upload",
              "tomo_padded_volume_bytes", "tomo_pad_volume", "tomo_forward", "tomo_back_adjoint",
              "tomo_back_voxel_bilinear", "tomo_proj_grad_workspace_bytes", "tomo_proj_grad"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    L = ctypes.CDLL(_lib.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(L, s), s


def test_version_and_constants_match_header():
    src = open(HEADER).read()
    L = _lib.load()
    assert L.tomo_version() == int(re.search(r"#define TOMO_B200_VERSION (\d+)", src).group(1))
    assert _lib.VIEW_STRIDE == int(re.search(r"#define TOMO_VIEW_STRIDE\s+(\d+)", src).group(1))
    assert _lib.POSE_STRIDE == int(re.search(r"#define TOMO_POSE_STRIDE\s+(\d+)", src).group(1))
    assert _lib.PAD == int(re.search(r"#define TOMO_PAD\s+(\d+)", src).group(1))
    assert L.tomo_views_bytes(10) == 10 * _lib.VIEW_STRIDE * 8


def test_struct_layout_and_size_queries():
    # 5 int32 (+4 pad) + 13 doubles
    assert ctypes.sizeof(TomoGeom) == 24 + 13 * 8
    L = _lib.load()
    g = TomoGeom()
    g.nx, g.ny, g.nz, g.ndx, g.ndz = 30, 20, 10, 30, 10
    nzp = ((10 + 2 * _lib.PAD + 31) // 32) * 32
    assert L.tomo_padded_volume_bytes(ctypes.byref(g)) == 4 * ((30 + 4) * (20 + 4) * nzp + 64)      # + 32 floats of slack before and after
    # block partials of the generic kernel (8 x 32 ray tiles) + of the separable kernel (8 ix x z chunks)
    assert L.tomo_proj_grad_workspace_bytes(ctypes.byref(g), 3) == 8 * 7 * (((30 + 7) // 8) * 1 + ((30 + 7) // 8) * 1 + ((30 + 31) // 32) * 1) * 3   # generic + separable + z-quad tiles


def test_library_is_sm100a_only():
    """cuobjdump lists exactly one ELF target: sm_100a (no multi-arch dispatch)."""
    import shutil
    import subprocess
    cu = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cu):
        return
    out = subprocess.run([cu, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, out
