"""Turn the `ncu --set full` captures that scripts/gpu_ncu_one.sh leaves in gpurun_out/ into what profiles/ keeps (no GPU needed):
  profiles/<tag>_ncu_<kernel>.csv     the raw page of each capture (`ncu -i ... --page raw --csv`)
  profiles/<tag>_traffic_512x720.json the counters bench.py and profiles/README.md quote (DRAM bytes per launch, issue, pipes, stalls)
  profiles/<tag>_launches_512x720.csv the launch list of the same command

    python scripts/ncu_summary.py <capture tag in gpurun_out> <tag in profiles> [kernel ...]
"""
import csv, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC, DST = sys.argv[1], sys.argv[2]
KERNELS = sys.argv[3:] or ["ray_kernel_forward", "ray_kernel_gradient", "adjoint_tile_kernel"]
KEYS = {
    "dram_bytes_read": "dram__bytes_read.sum", "dram_bytes_write": "dram__bytes_write.sum",
    "duration_ms_under_ncu": "gpu__time_duration.sum", "warp_inst_executed": "smsp__inst_executed.sum",
    "issue_active_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "pipe_fma_pct": "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "pipe_fmaheavy_pct": "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "pipe_alu_pct": "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "pipe_lsu_inst_pct": "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "lsu_wavefronts_pct": "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed",
    "shared_wavefronts": "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "l1_hit_pct": "l1tex__t_sector_hit_rate.pct", "l2_hit_pct": "lts__t_sector_hit_rate.pct",
    "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active", "registers": "launch__registers_per_thread",
    "eligible_warps_per_cycle": "smsp__warps_eligible.avg.per_cycle_active",
}
for st in ("long_scoreboard", "short_scoreboard", "math_pipe_throttle", "barrier", "wait", "not_selected"):
    KEYS["stall_" + st] = "smsp__average_warps_issue_stalled_%s_per_issue_active.ratio" % st
SCALE = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "ns": 1e-6, "s": 1e3}

out = {"source": "ncu --set full --clock-control none --import-source on, one capture per kernel and per gpurun call (-k regex:<kernel> "
                 "-s 1 -c 1) of `python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --host-phantom` (512^3 x 720 views, "
                 "1 GPU, scripts/gpu_ncu_one.sh); raw pages: profiles/%s_ncu_<kernel>.csv; launch list of the same command: "
                 "profiles/%s_launches_512x720.csv" % (DST, DST),
       "workload": {"size": 512, "views": 720, "n_gpus": 1}, "kernels": {}}
old = os.path.join(ROOT, "profiles", "%s_traffic_512x720.json" % DST)
if os.path.exists(old):                      # kernels not re-captured keep their entries
    out["kernels"] = json.load(open(old)).get("kernels", {})
for k in KERNELS:
    rep = os.path.join(ROOT, "gpurun_out", "prof_%s_%s.ncu-rep" % (SRC, k))
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True, check=True).stdout
    open(os.path.join(ROOT, "profiles", "%s_ncu_%s.csv" % (DST, k)), "w").write(raw)
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, val = rows[0], rows[1], rows[2]
    col = {h: i for i, h in enumerate(hdr)}
    rec = {}
    for key, metric in KEYS.items():
        v = float(val[col[metric]].replace(",", ""))
        rec[key] = v * SCALE.get(units[col[metric]], 1.0) if key.startswith(("dram", "duration")) else v
    rec["traffic_bytes"] = rec["dram_bytes_read"] + rec["dram_bytes_write"]
    out["kernels"][k] = rec
    print(k, "%.1f ms, dram %.2f GB, %.3g warp instr, issue %.1f %%, lsu wavefronts %.1f %%" % (
        rec["duration_ms_under_ncu"], rec["traffic_bytes"] / 1e9, rec["warp_inst_executed"], rec["issue_active_pct"], rec["lsu_wavefronts_pct"]))
json.dump(out, open(old, "w"), indent=1)
ll = os.path.join(ROOT, "gpurun_out", "%s_launches_512x720.csv" % SRC)
if os.path.exists(ll):
    shutil.copy(ll, os.path.join(ROOT, "profiles", "%s_launches_512x720.csv" % DST))
