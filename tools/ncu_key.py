#!/usr/bin/env python
"""Key numbers of an ncu report (--set full), per captured launch: duration, warp instructions, issue utilisation,
occupancy, DRAM bytes, stall reasons per issued instruction.  python tools/ncu_key.py gpurun_out/x.ncu-rep"""
import csv, io, subprocess, sys
KEYS = ['gpu__time_duration.sum', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'launch__registers_per_thread', 'smsp__warps_eligible.avg.per_cycle_active',
        'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed',
        'smsp__thread_inst_executed_per_inst_executed.ratio', 'lts__t_sector_hit_rate.pct', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed']
raw = subprocess.run(['ncu', '-i', sys.argv[1], '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
for r in rows[2:]:
    d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
    print(d['Kernel Name'][:40], d.get('Grid Size'), d.get('Block Size'))
    for k in KEYS:
        if k in d: print('  %-90s %s %s' % (k, d[k], u[k]))
    st = [(float(d[k]), k) for k in d if 'issue_stalled' in k and k.endswith('per_issue_active.ratio') and d[k]]
    for v, k in sorted(st, reverse=True)[:9]:
        print('  stall %-40s %.2f' % (k.split('issue_stalled_')[1].split('_per_issue')[0], v))
