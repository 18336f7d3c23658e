#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page) per captured launch: python scripts/ncu_summary.py REPORT.ncu-rep"""
import csv, subprocess, sys, io
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem',
        'sm__warps_active.avg.pct_of_peak_sustained_active',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct', 'lts__t_sector_hit_rate.pct',
        'sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum', 'launch__waves_per_multiprocessor']
for r in rows[2:]:
    print('-----')
    for w in want:
        if w in hdr:
            i = hdr.index(w); print('  %-70s %s %s' % (w, r[i], units[i]))
    st = []
    for i, h in enumerate(hdr):
        if h.startswith('smsp__pcsamp_warps_issue_stalled') and not h.endswith('not_issued'):
            try: st.append((float(r[i].replace(',', '')), h.replace('smsp__pcsamp_warps_issue_stalled_', '')))
            except ValueError: pass
    tot = sum(v for v, _ in st) or 1
    print('  stalls: ' + ', '.join('%s %.0f%%' % (h, 100 * v / tot) for v, h in sorted(st, reverse=True)[:7]))
