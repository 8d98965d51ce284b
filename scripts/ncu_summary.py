#!/usr/bin/env python
"""Prints the metrics of an .ncu-rep that the round notes quote (run here, no GPU needed): python scripts/ncu_summary.py file.ncu-rep"""
import csv, subprocess, sys, io
WANT = ['gpu__time_duration.sum','launch__registers_per_thread','launch__grid_size','launch__block_size','launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active','smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum','dram__bytes_write.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed','lts__t_sector_hit_rate.pct',
        'l1tex__t_sector_hit_rate.pct','lts__t_bytes.sum','l1tex__t_bytes.sum','sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum',
        'sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active',
        'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active','smsp__inst_executed_op_local_ld.sum','smsp__inst_executed_op_local_st.sum',
        'smsp__inst_executed_op_global_ld.sum','l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum','l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum',
        'l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum','l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum']
for path in sys.argv[1:]:
    out = subprocess.run(['ncu', '-i', path, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        print('=====', path, '|', d['Kernel Name'][:60])
        for k in WANT:
            if k in d:
                print(f'  {k} = {d[k]} {units[hdr.index(k)]}')
        for k in hdr:
            if 'warp_issue_stalled' in k and k.endswith('_per_warp_active.pct'):
                try:
                    v = float(d[k])
                except ValueError:
                    continue
                if v > 2:
                    print('  stall', k.replace('smsp__warp_issue_stalled_', '').replace('_per_warp_active.pct', ''), f'{v:.1f}')
