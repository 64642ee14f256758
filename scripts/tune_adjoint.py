"""Time the operators of whatever library TOMO_B200_LIB points at (tuning only)."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tomography_alignment_b200 import Geometry, pose_table
from tomography_alignment_b200.cuda_backend import CudaBackend
from tomography_alignment_b200.phantom import benchmark_poses
n, n_proj = int(sys.argv[1]), int(sys.argv[2])
which = sys.argv[3] if len(sys.argv) > 3 else "fbg"
g = Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
phi, alpha, beta, xyz = benchmark_poses(720)
if os.environ.get("UNTILTED"):
    alpha, beta = alpha * 0, beta * 0
sel = np.linspace(0, 719, n_proj).astype(int)
be = CudaBackend(g, "cuda:0", zquad=bool(os.environ.get("ZQUAD")))
be.set_poses(pose_table(np.array([phi, alpha, beta]).T[sel], xyz[sel], g.cor_shift))
torch.manual_seed(0)
vol = torch.rand((n, n, n), device="cuda")
y = torch.rand((n_proj, n, n), device="cuda")
bp = torch.empty((n, n, n), device="cuda")
proj = torch.empty((n_proj, n, n), device="cuda")
def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return s.elapsed_time(e) / reps
out = {"lib": os.path.basename(os.environ.get("TOMO_B200_LIB", "default")), "n": n, "views": n_proj, "zquad": bool(os.environ.get("ZQUAD"))}
if "f" in which: out["fwd_ms"] = t(lambda: be.forward(vol, out=proj))
if "b" in which:
    out["back_ms"] = t(lambda: be.adjoint(y, out=bp))
    out["back_checksum"] = int(bp.view(torch.int32).to(torch.int64).sum().item())      # bitwise fingerprint for A/B builds
if "v" in which: out["voxback_ms"] = t(lambda: be.voxel_back(y, out=bp))
if "g" in which: out["grad_ms"] = t(lambda: be.proj_grad(vol, meas=y, want_proj=False, want_dproj=False, repad=False))
print(json.dumps(out))
