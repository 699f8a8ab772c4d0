"""tools/ncu_summary.py -- print the handful of ncu metrics we track from a .ncu-rep (read on CPU).
usage: python tools/ncu_summary.py report.ncu-rep"""
import csv, subprocess, sys
rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
r = list(csv.reader(txt.splitlines()))
hdr = r[0]
rows = r[2:]
want = ['Kernel Name', 'launch__grid_size', 'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active']
for i, h in enumerate(hdr):
    if h in want:
        print("%-70s %s" % (h, [row[i][:48] for row in rows]))
st = []
for i, h in enumerate(hdr):
    if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio'):
        st.append((h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''), [row[i] for row in rows]))
st.sort(key=lambda x: -float(x[1][0] or 0))
print("stalls (warps per issue-active cycle):")
for h, v in st[:8]:
    print("  %-28s %s" % (h, v))
