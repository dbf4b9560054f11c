#!/usr/bin/env python
"""Print the handful of ncu raw-page metrics that matter for a kernel: python tools/ncu_keys.py <report.ncu-rep>"""
import csv, subprocess, sys, io
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'launch__registers_per_thread', 'launch__occupancy_limit',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'sm__inst_executed_pipe_tensor', 'sm__pipe_tensor', 'smsp__pcsamp_warps_issue_stalled',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum',
        'l1tex__t_requests_pipe_lsu_mem_global_op_st.sum', 'lts__t_sectors_srcunit_tex_op_read.sum', 'lts__t_sectors_srcunit_tex_op_write.sum',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum', 'sm__cycles_elapsed.avg', 'sm__inst_executed_pipe']
for r in rows[2:]:
    print('==', r[hdr.index('Kernel Name')] if 'Kernel Name' in hdr else '')
    for h, u, v in zip(hdr, units, r):
        if any(h.startswith(w) for w in want) and 'pct_of_peak_sustained_elapsed' not in h and '.per_second' not in h:
            print(f'  {h} [{u}] = {v}')
