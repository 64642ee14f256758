#!/usr/bin/env python
"""bench.py -- headline benchmark of the projection hot path (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size 512] [--views 720] [--impl b200|reference]

One *step* = one forward projection of all views + one exact-adjoint backprojection of all views + one
projection-and-6-DOF-gradient pass over all views (SURVEY.md section 8d "s per alignment iter") of a
Shepp-Logan phantom with the jittered poses of examples/generate_data.py (seeded).  The unit of work is the
voxel-ray update: n_vox * n_proj per operator, 3 operators per step.

N > 1 (launched by torchrun, one rank per GPU): the views are sharded across ranks (interleaved: view i on rank
i mod N, so that every rank sees the whole angular range; --shard contiguous gives the np.array_split blocks of
recon/sirt_mpi.py:40), the volume is replicated, the backprojected volume is summed with an NCCL all-reduce that is
queued behind the backprojection and waited for after the gradient kernel, and the per-view gradient table is
assembled with a zero-padded all-reduce; the total problem is fixed (strong scaling).  The end-to-end path keeps the
volume, the backprojection and the projections in host buffers all ranks share (POSIX shared memory, page-locked):
each rank moves its 1/N over its own PCIe link.

Output: ONE JSON line on rank 0 (see the keys at the bottom).  `value` times device-resident inputs with CUDA
events; `e2e` times the same three operators through the public ProjectionMatrix API with pinned HOST
buffers, host<->device copies inside the timed region.
--impl reference times the CPU oracle port of the reference's loops on the host cores (the reference's own
Fortran cannot be built here: no gfortran).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "voxel-ray updates/s fwd+back+grad"
UNIT = "voxel-ray updates/s"


_T0 = time.perf_counter()

# stdout carries exactly one JSON line: keep a private handle on the real stdout and point fd 1 at stderr, so that
# nothing a library prints there (NCCL's version banner, for one) can end up next to the result
_RESULT_OUT = os.fdopen(os.dup(1), "w")
os.dup2(2, 1)


def emit(line):
    _RESULT_OUT.write(json.dumps(line) + "\n")
    _RESULT_OUT.flush()


def log(msg):
    """Progress to stderr (stdout carries only the JSON line)."""
    sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--views", type=int, default=720)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-size", type=int, default=256, help="volume size of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--slabs", type=int, default=1,
                    help="x-slabs of the backprojection (all-reduce of slab k under the kernel of slab k+1); 1 = one launch on a side "
                         "stream, its all-reduce hidden behind the gradient kernel (measured best: profiles/README.md)")
    ap.add_argument("--shard", default="interleaved", choices=["interleaved", "contiguous"],
                    help="view sharding for N > 1: interleaved (balanced) or the reference's contiguous np.array_split blocks")
    ap.add_argument("--host-phantom", action="store_true",
                    help="build the phantom with numpy on the host (profiling runs: keeps torch's phantom kernels out of ncu launch lists)")
    return ap.parse_args()


def workload_name(a):
    return "%d^3 x %d views, fwd + adjoint back + 6-DOF grad" % (a.size, a.views)


# ------------------------------------------------------------------------------------------------------
# CPU leg (oracle port): cpu_baseline of the b200 arm and the whole --impl reference arm
# ------------------------------------------------------------------------------------------------------
def cpu_sample(cpu_size, target_s=12.0):
    """Time fwd+back+grad of the oracle port on a bounded sample: cpu_size^3, one view per worker at a time,
    all host cores (capped at 64 workers: each keeps a private float64 volume, 128 MiB at 256^3)."""
    from oracle import oracle as O
    from tomography_alignment_b200.phantom import benchmark_poses
    cores = os.cpu_count() or 1
    workers = max(1, min(cores, 64))
    # calibrate on one view per worker, then size the sample for ~target_s seconds
    dt, upd = O.cpu_fwd_back_grad(cpu_size, workers, workers, benchmark_poses(workers))
    rounds = int(max(1, min(8, target_s / max(dt, 1e-3))))
    if rounds > 1:
        n_views = workers * rounds
        dt, upd = O.cpu_fwd_back_grad(cpu_size, n_views, workers, benchmark_poses(n_views))
    else:
        n_views = workers
    return {"value": upd / dt, "unit": UNIT, "cores": workers, "kind": "port",
            "sample": "%d^3 x %d views fwd+back+grad, views split over %d threads (np.array_split as in "
                      "recon/*_mpi.py), C port of src/ray_wt_grad.f90 loops at -O3, %.1f s"
                      % (cpu_size, n_views, workers, dt)}, dt


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    per_step = []
    info = None
    for i in range(a.warmup + a.steps):
        info, dt = cpu_sample(a.cpu_size, target_s=6.0)
        if i >= a.warmup:
            per_step.append((info["value"], dt))
    val = float(np.mean([v for v, _ in per_step]))
    ms = float(np.mean([d for _, d in per_step])) * 1e3
    info["value"] = val
    line = {"metric": METRIC, "value": val, "unit": UNIT, "impl": "reference", "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(a), "note": "each step is the bounded CPU sample described in "
                       "cpu_baseline.sample; throughput is size-independent to first order"},
            "cpu_baseline": info,
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.rows = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for ln in self.proc.stdout:
            self.rows.append(ln.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------------
# B200 arm
# ------------------------------------------------------------------------------------------------------
def run_b200(a):
    import torch
    import torch.distributed as dist
    from tomography_alignment_b200 import Geometry, ProjectionMatrix, pose_table
    from tomography_alignment_b200.cuda_backend import CudaBackend
    from tomography_alignment_b200.phantom import benchmark_poses, shepp3d
    from tomography_alignment_b200.sharding import SharedHostBuffer, adjoint_allreduce, shard_views

    log("torch imported")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    n, n_proj = a.size, a.views
    geo = Geometry(n_proj, np.array([n, n, n]), np.ones(3), np.array([n, n]), np.ones(2))
    phi, alpha, beta, xyz = benchmark_poses(n_proj)
    # interleaved shards (view i on rank i mod N): every rank sees the whole angular range, so the ranks stay balanced (the cost
    # of a view depends on its angle; with the reference's contiguous np.array_split blocks every all-reduce waits 2.7 ms of
    # 88 ms for the slowest rank at 8 GPUs, profiles/r2_shard_balance.json).  --shard contiguous restores the reference's split.
    mine = shard_views(n_proj, world, rank, a.shard)
    my_n = len(mine)
    poses = pose_table(np.array([phi, alpha, beta]).T, xyz, geo.cor_shift)
    # "true" poses generate the measured data; the current estimate (half the jitter) is what fwd/grad use
    be_true = CudaBackend(geo, dev)
    be_true.set_poses(poses[mine])
    est = poses.copy()
    est[:, 1:3] *= 0.5
    est[:, 3:6] *= 0.5
    be = CudaBackend(geo, dev)
    be.set_poses(est[mine])

    # small phantoms are built on the host (keeps profiler launch lists free of phantom kernels)
    vol_true = torch.as_tensor(shepp3d(n)).to(dev) if (n <= 256 or a.host_phantom) else shepp3d(n, device=dev)
    meas = be_true.forward(vol_true).clone()              # b = A_true phantom, (my_n, n, n)
    vol = (0.9 * vol_true).contiguous()                    # current reconstruction estimate
    proj = torch.empty_like(meas)
    bp = torch.empty((n, n, n), dtype=torch.float32, device=dev)
    table = torch.zeros((n_proj, 7), dtype=torch.float64, device=dev)
    idx = torch.as_tensor(mine, device=dev, dtype=torch.long)
    del be_true
    torch.cuda.synchronize()
    log("inputs ready (phantom, measured projections)")

    n_vox, n_det = float(n) ** 3, float(n) ** 2
    updates_per_step = 3.0 * n_vox * n_proj

    def step():
        be.forward(vol, out=proj)                                   # pad + forward, all local views
        res = meas - proj                                            # residual (elementwise, torch)
        # exact adjoint of all local views in x-slabs; the NCCL all-reduce of a finished slab (recon/sirt_mpi.py:103) runs
        # on NVLink while the next slab -- and, for the last slabs, the gradient kernel -- is computed
        _, works = adjoint_allreduce(be, res, bp, None, a.slabs, reduce=world > 1)
        out = be.proj_grad(vol, meas=meas, want_proj=False, want_dproj=False, repad=False)
        if world > 1:
            table.zero_()
            table[idx, :6] = out["grad6"]
            table[idx, 6] = out["cost"]
            dist.all_reduce(table)
        for w in works:                                              # the step ends with the summed volume complete
            w.wait()
        return out

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(a.warmup):
        step()
    sync_all()
    log("warm-up done")
    launches0 = be.launches
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        out = step()
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    clocks = sampler.stop() if rank == 0 else None
    launches = be.launches - launches0
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_per_step = t.item() / a.steps
    value = updates_per_step / (ms_per_step * 1e-3)

    # per-kernel timing for the roofline of the dominant kernel (rank 0's shard, same buffers)
    def time_kernel(fn, reps=2):
        fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        for _ in range(reps):
            fn()
        e.record()
        torch.cuda.synchronize()
        return s.elapsed_time(e) / reps

    volpad = be.pad(vol)
    import ctypes
    from tomography_alignment_b200 import _lib
    L = be.lib

    def k_fwd():
        _lib.check(L.tomo_forward_ex(be._g(), ctypes.c_void_p(be.views.data_ptr()), my_n, be.kinds, ctypes.c_void_p(volpad.data_ptr()),
                                     ctypes.c_void_p(proj.data_ptr()), be._stream()), "tomo_forward")

    def k_back():
        _lib.check(L.tomo_back_adjoint_slab(be._g(), ctypes.c_void_p(be.views.data_ptr()), my_n, be.kinds,
                                            ctypes.c_void_p(meas.data_ptr()), ctypes.c_void_p(bp.data_ptr()), 0, None, 0, 0, n,
                                            be._stream()), "tomo_back_adjoint")

    def k_grad():
        be.proj_grad(vol, meas=meas, want_proj=False, want_dproj=False, repad=False)

    log("timed steps done: %.1f ms/step" % ms_per_step)
    t_f, t_b, t_g = time_kernel(k_fwd), time_kernel(k_back), time_kernel(k_grad)
    log("per-kernel timing done")
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # algorithmic bytes per launch: every operator touches each fp32 voxel once per view
    # plus detector-side traffic
    bytes_f = my_n * (4 * n_vox + 4 * n_det)
    bytes_b = my_n * (4 * n_vox + 4 * n_det)
    bytes_g = my_n * (4 * n_vox + 4 * n_det) + 48 * my_n
    kernels = {"ray_kernel_forward": (t_f, bytes_f), "adjoint_tile_kernel": (t_b, bytes_b),
               "ray_kernel_gradient": (t_g, bytes_g)}
    dom = max(kernels, key=lambda k: kernels[k][0])
    ach = kernels[dom][1] / (kernels[dom][0] * 1e-3) / 1e9
    # measured DRAM traffic of the dominant kernel (ncu --set full capture of this exact workload, profiles/)
    traffic, traffic_src = None, None
    tpath = os.path.join(ROOT, "profiles", "r2_traffic_512x720.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        w = tj.get("workload", {})
        if (w.get("size"), w.get("views"), w.get("n_gpus")) == (n, n_proj, world):
            traffic = tj["kernels"].get(dom, {}).get("traffic_bytes")
            traffic_src = "profiles/r2_traffic_512x720.json"
    roofline = {"bound": "hbm", "kernel": dom, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic, "traffic_unit": "bytes per launch (dram read + write, ncu)", "traffic_source": traffic_src,
                "algorithmic_bytes_per_launch": kernels[dom][1], "peak_source": peak_src,
                "per_kernel_ms": {k: v[0] for k, v in kernels.items()},
                "per_kernel_algorithmic_gbs": {k: v[1] / (v[0] * 1e-3) / 1e9 for k, v in kernels.items()},
                "step_frac_of_peak": (bytes_f + bytes_b + bytes_g) * world / (ms_per_step * 1e-3) / 1e9 / peak / world}

    # secondary measurement: the same three kernels on UNTILTED poses (alpha = beta = 0, the default of
    # projection_matrix and every known-geometry reconstruction), where forward and adjoint take the separable kernels
    flat = est[mine].copy()
    flat[:, 1:3] = 0.0
    be.set_poses(flat)
    u_f, u_b, u_g = time_kernel(k_fwd), time_kernel(lambda: be.adjoint(meas, out=bp)), time_kernel(k_grad)
    be.set_poses(est[mine])
    untilted = {"per_kernel_ms": {"sep_forward_kernel": u_f, "sep_adjoint_kernel(+zgather)": u_b, "sep_gradient_kernel": u_g},
                "forward_frac_of_peak": bytes_f / (u_f * 1e-3) / 1e9 / peak,
                "adjoint_frac_of_peak": bytes_b / (u_b * 1e-3) / 1e9 / peak,
                "gradient_frac_of_peak": bytes_g / (u_g * 1e-3) / 1e9 / peak,
                "step_ms": u_f + u_b + u_g}
    log("untilted-pose timing done")
    # secondary measurement: the voxel-driven bilinear backprojector (src/back_projection.f90), TMA-staged kernel
    t_v = time_kernel(lambda: be.voxel_back(meas, out=bp))
    voxel_driven = {"kernel": "voxel_bilinear_tma_kernel", "ms": t_v, "algorithmic_gbs": bytes_b / (t_v * 1e-3) / 1e9,
                    "frac_of_peak": bytes_b / (t_v * 1e-3) / 1e9 / peak}
    log("voxel-driven backprojector timing done")

    # end to end through the public API with pinned host buffers
    e2e = None
    if not a.no_e2e:
        pm = ProjectionMatrix(geo, precision=np.float32, device=dev, backend=be)
        A = pm.projection_matrix(phi=est[mine, 0], alpha=est[mine, 1], beta=est[mine, 2], xyz_shift=est[mine, 3:6])
        h_vol = vol.cpu().pin_memory()
        h_meas = meas.cpu().pin_memory()
        h_proj = torch.empty((my_n, n, n), dtype=torch.float32).pin_memory()
        h_bp = torch.empty((n, n, n), dtype=torch.float32).pin_memory()
        log("pinned host buffers ready")

        d_vol = torch.empty((n, n, n), dtype=torch.float32, device=dev)
        g_tab = torch.zeros((n_proj, 7), dtype=torch.float64, device=dev)
        # shared page-locked host buffers up to 8 GB in total (larger registrations were refused on the 8-GPU box at 1024^3:
        # the per-process pinned path below is used then)
        shared_bytes = 4.0 * (2 * n ** 3 + n_proj * n * n)
        sharded_io = world > 1 and n % world == 0 and shared_bytes <= float(os.environ.get("TOMO_BENCH_SHARED_HOST_BYTES", 8e9))
        if sharded_io:
            # one box, one host memory: the volume, the backprojection and the projections live in host buffers every rank
            # sees (POSIX shared memory, page-locked in each process); each rank moves only its 1/N over its own PCIe link
            job = "tomo_b200_bench_%s" % os.environ.get("MASTER_PORT", "0")      # one name per job on the box
            sh_vol = SharedHostBuffer(job + "_vol", (n, n, n))
            sh_bp = SharedHostBuffer(job + "_bp", (n, n, n))
            per_rank = (n_proj + world - 1) // world
            sh_proj = SharedHostBuffer(job + "_proj", (world, per_rank, n, n))      # rank-major: block r = views of rank r
            if rank == 0:
                sh_vol.tensor.copy_(h_vol)
            dist.barrier()
            xs = n // world
            my_x = slice(rank * xs, (rank + 1) * xs)
            bp_slab = torch.empty((xs, n, n), dtype=torch.float32, device=dev)
            log("shared host buffers ready (pinned: %s)" % (sh_vol.pinned and sh_bp.pinned and sh_proj.pinned))

        def upload_volume():
            """The replicated volume, once per step: every rank uploads its x-slab from the shared host buffer and the slabs
            are all-gathered over NVLink (rank 0 uploading everything and broadcasting when nx does not divide)."""
            if world == 1:
                return None
            if sharded_io:
                d_vol[my_x].copy_(sh_vol.tensor[my_x], non_blocking=True)
                be.h2d_bytes += 4 * xs * n * n
                dist.all_gather_into_tensor(d_vol, d_vol[my_x])
                return d_vol
            if rank == 0:
                d_vol.copy_(h_vol, non_blocking=True)
                be.h2d_bytes += 4 * h_vol.numel()
            dist.broadcast(d_vol, src=0)
            return d_vol

        def e2e_step():
            # host buffers in, host buffers out; copies are issued inside the calls (view chunks, side stream)
            if world == 1:
                be.forward_host(h_vol, out_host=h_proj)                         # H2D volume, forward, D2H projections
                be.adjoint_host(h_meas, out_host=h_bp, wait=False)              # H2D projections, adjoint, D2H volume (queued)
                out = be.proj_grad_host(None, h_meas, vol_dev=be._buf("vol", be.vol_shape))   # H2D measured, D2H (n, 6) gradients
                be.sync_host()                                                  # the volume download (under the gradient kernel) has landed
                return out
            dv = upload_volume()                                                # H2D volume: once per step, 1/N per rank
            out_rows = sh_proj.tensor[rank, :my_n] if sharded_io else h_proj
            be.forward_host(None, out_host=out_rows, vol_dev=dv)                # forward, D2H of this rank's views
            v = be.adjoint_host(h_meas, out_host=None, to_host=False)           # H2D projections, adjoint (device volume)
            if sharded_io:                                                      # reduce-scatter instead of the Allreduce of
                rs = dist.reduce_scatter_tensor(bp_slab, v, async_op=True)      # recon/sirt_mpi.py:103, under the gradient kernel
            else:
                dist.all_reduce(v)
                if rank == 0:
                    h_bp.copy_(v)
                    be.d2h_bytes += 4 * v.numel()
            g6, c = be.proj_grad_host(None, h_meas, vol_dev=dv, to_host=False)
            if sharded_io:                                                      # each rank downloads its summed x-slab into
                rs.wait()                                                       # the shared host buffer
                sh_bp.tensor[my_x].copy_(bp_slab, non_blocking=True)
                be.d2h_bytes += 4 * bp_slab.numel()
            g_tab.zero_()
            g_tab[idx, :6] = g6
            g_tab[idx, 6] = c
            dist.all_reduce(g_tab)                                              # zero-padded all-reduce = all-gather
            out = g_tab.cpu() if rank == 0 else None
            if rank == 0:
                be.d2h_bytes += 8 * g_tab.numel()
            torch.cuda.current_stream().synchronize()                           # this rank's downloads have landed
            return out

        e2e_step()
        sync_all()
        t0 = time.perf_counter()
        reps = max(1, min(a.steps, 2))
        for _ in range(reps):
            e2e_step()
        sync_all()
        dt = (time.perf_counter() - t0) / reps
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        be.h2d_bytes = be.d2h_bytes = 0
        e2e_step()                                                   # untimed: counts the bytes one step moves
        sync_all()
        h2d, d2h = be.h2d_bytes, be.d2h_bytes
        if world > 1:                                                # whole-job bytes: every rank's copies
            bt = torch.tensor([h2d, d2h], dtype=torch.int64, device=dev)
            dist.all_reduce(bt)
            h2d, d2h = int(bt[0].item()), int(bt[1].item())
        if sharded_io:
            for b_ in (sh_vol, sh_bp, sh_proj):
                b_.close()
        log("e2e done")
        e2e = {"value": updates_per_step / tt.item(), "unit": UNIT, "h2d_bytes_per_step": int(h2d),
               "d2h_bytes_per_step": int(d2h), "ms_per_step": tt.item() * 1e3,
               "bytes_are": "summed over ranks", "host_buffers": "shared (POSIX shm, page-locked per rank): each rank moves 1/N"
                                                               if sharded_io else "per-process pinned"}

    if rank == 0:
        cpu = None
        if not a.no_cpu_baseline and world == 1:
            cpu, _ = cpu_sample(a.cpu_size)
            log("cpu baseline done")
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload_name(a), "n_vox": int(n_vox), "n_proj": n_proj,
                           "views_per_gpu": int(my_n),
                           "parallelism": "views sharded x%d (%s), volume replicated, backprojection all-reduced under the gradient kernel"
                                          % (world, a.shard),
                           "l2": "inputs exceed L2 (volume %d MiB, projections %d MiB per rank)"
                                 % (4 * n ** 3 // 2 ** 20, 4 * my_n * n * n // 2 ** 20),
                           "phantom": "shepp3d", "poses": "examples/generate_data.py jitter, seed 20240229"},
                "roofline": roofline, "untilted_poses": untilted, "voxel_driven_backprojector": voxel_driven, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches),
                "clocks": clocks, "impl": "b200"}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)
