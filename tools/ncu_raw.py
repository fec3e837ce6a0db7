#!/usr/bin/env python
"""Key raw metrics of an .ncu-rep:  python tools/ncu_raw.py rep [regex]"""
import csv, subprocess, sys, re
rep = sys.argv[1]
pat = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
out = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr, units = rows[0], rows[1]
DEFAULT = r'gpu__time_duration.sum|dram__bytes_(read|write).sum$|gpu__dram_throughput.avg.pct|sm__warps_active.avg.pct|launch__registers|launch__occupancy_limit|launch__grid_size|launch__block_size|smsp__thread_inst_executed_per_inst|sm__throughput.avg.pct|smsp__inst_executed.sum$|smsp__issue_active.avg.pct|smsp__average_warps_issue_stalled_.*_per_issue_active|l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum$|lts__t_sector_hit_rate.pct|sm__inst_executed_pipe_(alu|fma|lsu|xu|uniform|fp64).*sum$|l1tex__data_pipe_lsu_wavefronts_mem_shared.sum$|smsp__cycles_active.avg$|launch__shared_mem_per_block|sm__cycles_elapsed.max'
pat = pat or re.compile(DEFAULT)
for r in rows[2:]:
    print('---', r[hdr.index('Kernel Name')][:60])
    for i, h in enumerate(hdr):
        if pat.search(h):
            print(f'  {h:75s} {r[i]:>16s} {units[i]}')
