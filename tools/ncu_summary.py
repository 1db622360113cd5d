#!/usr/bin/env python
"""Summarise an .ncu-rep (raw page) into the handful of numbers the design discussion needs."""
import csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
h = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'launch__registers_per_thread', 'launch__grid_size', 'launch__block_size',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'smsp__inst_executed.sum',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_tensor.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.sum', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__inst_executed_pipe_xu.sum',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'sm__cycles_elapsed.max', 'lts__t_bytes.sum',
        'smsp__average_warp_latency_issue_stalled_barrier.ratio']
for r in rows[2:]:
    for w in want:
        for i, x in enumerate(h):
            if x == w:
                print(f"{w:80s} {r[i][:90]} {rows[1][i]}")
    # stall breakdown
    st = [(x, r[i]) for i, x in enumerate(h) if x.startswith('smsp__average_warps_issue_stalled_') and x.endswith('_per_issue_active.ratio')]
    st = sorted(((float(v), x) for x, v in st if v not in ('', 'n/a')), reverse=True)[:8]
    for v, x in st:
        print(f"   stall {x.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''):40s} {v:.3f}")
    print('-' * 60)
