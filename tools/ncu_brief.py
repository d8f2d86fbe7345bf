"""Development aid: the handful of ncu counters the assembly / projection work is steered by, per kernel of a report."""
import csv, subprocess, sys
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
want = ['gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'smsp__inst_executed.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct']
stall = [h for h in hdr if 'smsp__average_warps_issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print("==", d.get('Kernel Name', '')[:70])
    for w in want:
        if w in d: print("  %-70s %s %s" % (w, d[w], rows[1][hdr.index(w)]))
    st = sorted([(float(d[h].replace(',', '')) if d[h] not in ('', 'n/a') else 0, h) for h in stall], reverse=True)[:7]
    print("  stalls per issue:", ", ".join("%s %.2f" % (h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), v) for v, h in st))
