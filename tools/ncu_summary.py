"""Summarise an .ncu-rep (raw page + source page) into text for profiles/.
    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [top_n]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size",
    "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "launch__occupancy_limit_warps", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "sm__maximum_warps_per_active_cycle_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.per_cycle_active", "smsp__inst_executed.sum", "sm__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__cycles_elapsed.max",
    "sm__cycles_active.avg", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__average_warp_latency_issue_stalled_barrier.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("== kernel:", d.get("Kernel Name"), "id", d.get("ID"))
    for k in KEYS:
        if k in d:
            print(f"  {k} = {d[k]} {units[hdr.index(k)]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
if not src.strip():
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
if rows:
    h = rows[0]
    print("source columns:", h[:12], "...")
    def col(name):
        for i, x in enumerate(h):
            if x.strip().lower() == name.lower():
                return i
        return None
    ci = col("Warp Stall Sampling (All Samples)") or col("# Samples") or col("Samples")
    cs = col("Source")
    if ci is not None and cs is not None:
        body = [r for r in rows[1:] if len(r) > max(ci, cs) and r[ci].replace(".", "").isdigit()]
        tot = sum(float(r[ci]) for r in body) or 1.0
        body.sort(key=lambda r: -float(r[ci]))
        print(f"-- top {top} lines by stall samples (total {tot:.0f})")
        for r in body[:top]:
            print(f"  {100 * float(r[ci]) / tot:5.1f}%  {r[cs].strip()[:150]}")
