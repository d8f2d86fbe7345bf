"""Summarise an .ncu-rep (raw page) per kernel: duration, pipes, stalls, memory.  Usage: ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h, units, data = rows[0], rows[1], rows[2:]
ix = {n: i for i, n in enumerate(h)}
keys = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sectors_op_red.sum", "lts__t_sectors_op_atom.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct"]
stalls = [k for k in h if k.startswith("smsp__average_warps_issue_stalled_") and k.endswith("_per_issue_active.ratio")]
for r in data:
    print("=" * 100); print(r[ix["Kernel Name"]][:110])
    for k in keys:
        if k in ix: print(f"  {k:75s} {r[ix[k]]:>16s} {units[ix[k]]}")
    st = sorted(((float(r[ix[k]] or 0), k.replace("smsp__average_warps_issue_stalled_", "").replace("_per_issue_active.ratio", "")) for k in stalls), reverse=True)
    print("  stalls/issue:", ", ".join(f"{n}={v:.2f}" for v, n in st[:9]))
