"""One-screen summary of an ncu --set full report: python tools/ncu_summary.py rep.ncu-rep"""
import csv
import subprocess
import sys

WANT = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'launch__shared_mem_per_block_dynamic', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__pcsamp_warps_issue_stalled_long_scoreboard',
        'smsp__pcsamp_warps_issue_stalled_short_scoreboard', 'smsp__pcsamp_warps_issue_stalled_wait',
        'smsp__pcsamp_warps_issue_stalled_no_instructions', 'smsp__pcsamp_warps_issue_stalled_selected',
        'smsp__pcsamp_warps_issue_stalled_math_pipe_throttle', 'smsp__pcsamp_warps_issue_stalled_barrier',
        'smsp__pcsamp_warps_issue_stalled_branch_resolving', 'smsp__pcsamp_warps_issue_stalled_not_selected',
        'sass__inst_executed_local_loads', 'sass__inst_executed_local_stores']
out = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[0]
for w in WANT:
    for i in [i for i, h in enumerate(hdr) if h == w][:1]:
        print(hdr[i][:66].ljust(67) + ' | '.join(r[i][-42:] for r in rows[2:]))
