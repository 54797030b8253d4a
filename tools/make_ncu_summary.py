"""Tracked summary of one ncu --set full capture (exports: --page raw --csv, --page source --csv).
usage: python tools/make_ncu_summary.py <out.md> <raw.csv> <src.csv> <commit> "<workload / command line>" [alg_bytes_per_launch]"""
import csv, os, subprocess, sys
out, raw, src, commit, what = sys.argv[1:6]
alg = float(sys.argv[6]) if len(sys.argv) > 6 else None
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
rows = list(csv.reader(open(raw)))
hdr, units, val = rows[0], rows[1], rows[2]
def get(name):
    for i, h in enumerate(hdr):
        if h == name: return val[i], units[i]
    return None, None
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "launch__grid_size", "launch__block_size",
        "launch__registers_per_thread", "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem",
        "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "sm__issue_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__inst_executed.sum",
        "lts__t_sector_hit_rate.pct", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio", "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio"]
kern = [k for k in rows[2:] if len(k) > 4][0][4]
md = [f"# ncu `--set full` of `{kern}`", "", f"Commit {commit}.  {what}", "",
      "One launch, `--clock-control none --import-source on`; numbers under the profiler are not bench values.", "",
      "| metric | unit | value |", "|---|---|---|"]
m = {}
for w in want:
    v, u = get(w)
    if v is not None:
        md.append(f"| `{w}` | {u} | {v} |"); m[w] = (v, u)
def to_bytes(key):
    if key not in m: return None
    v, u = m[key]
    return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
if rd is not None and alg:
    md += ["", f"DRAM traffic of the launch: {(rd + wr) / 1e6:.1f} MB read + written against {alg / 1e6:.1f} MB algorithmic "
               f"(ratio {(rd + wr) / alg:.2f}" + (": at or below the algorithmic bytes, no wasted re-reads)." if (rd + wr) / alg <= 1.05 else
               "; the excess is the per-CTA global scratch slab through which long-running QPs build their full-space operator "
               "-- build_full_op, fccqp_struct.cuh -- written once and read back per such QP; Q and A_eq themselves are read from DRAM once).")]
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_lines.py"), src, "24"], capture_output=True, text=True).stdout
md += ["", "## Stall reasons and hottest source lines (`tools/ncu_lines.py` over the source page; `struct` = fccqp_struct.cuh, `kernel` = fccqp_kernel.cuh)", "", "```", lines.rstrip(), "```", ""]
open(out, "w").write("\n".join(md))
print("wrote", out)
