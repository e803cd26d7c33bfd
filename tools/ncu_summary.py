"""Prints the metrics we care about from an .ncu-rep (raw page).  usage: ncu_summary.py file.ncu-rep [kernel-index]"""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
h = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum", "smsp__inst_executed.sum",
        "launch__registers_per_thread", "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_pipe_lsu.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
stall = [c for c in h if c.startswith("smsp__average_warps_issue_stalled") and c.endswith("_per_issue_active.ratio")] or [c for c in h if "warp_issue_stalled" in c and c.endswith("pct")]
for row in rows[2:]:
    print("=" * 100)
    for k in want:
        if k in h:
            print(f"{k:75s} {row[h.index(k)]} {rows[1][h.index(k)]}")
    st = sorted(((float(row[h.index(c)] or 0), c) for c in stall), reverse=True)[:8]
    for v, c in st:
        print(f"   stall {c.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:.3f}")
