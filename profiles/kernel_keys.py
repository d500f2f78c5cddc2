#!/usr/bin/env python
"""Headline counters and top stall reasons of every launch in an .ncu-rep:  python profiles/kernel_keys.py <report> [name filter]
(the stand-alone kernels' summaries; the fused kernel has profiles/ncu_summary.py)."""
import csv, subprocess, sys, io
rep = sys.argv[1]
flt = sys.argv[2] if len(sys.argv) > 2 else ''
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['gpu__time_duration.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'dram__bytes_read.sum', 'dram__bytes_write.sum',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__inst_issued.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 
        
        'launch__registers_per_thread', 'launch__grid_size', 'launch__occupancy_limit_registers', 'sm__cycles_active.avg', 'sm__cycles_elapsed.avg',
        ]
seen = set()
for r in rows[2:]:
    name = r[hdr.index('Kernel Name')]
    if flt not in name or name in seen:
        continue
    seen.add(name)
    print('==', r[hdr.index('Kernel Name')][:90])
    for k in keys:
        if k in hdr:
            print('   %-75s %s %s' % (k, r[hdr.index(k)], units[hdr.index(k)]))
    stall = [i for i, h in enumerate(hdr) if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('per_issue_active.ratio')]
    st = sorted(((round(float(r[i]), 2), hdr[i].replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')) for i in stall), reverse=True)[:8]
    print('   stalls', st)
