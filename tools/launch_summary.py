#!/usr/bin/env python
"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
   python tools/launch_summary.py gpurun_out/launches.csv [skip_regex] > profiles/rN_launches_x_summary.txt"""
import csv
import re
import sys
from collections import OrderedDict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
skip = re.compile(sys.argv[2]) if len(sys.argv) > 2 else None
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
tot = OrderedDict()
for r in rows[1:]:
    name = re.sub(r'\(.*', '', r[ki]).replace('<unnamed>::', '')
    name = re.sub(r'^void ', '', name)[:90]
    if skip and skip.search(name):
        continue
    ns = float(r[vi].replace(',', '')) * ({'ns': 1, 'us': 1e3, 'ms': 1e6}.get(r[ui], 1))
    t = tot.setdefault(name, [0, 0.0])
    t[0] += 1
    t[1] += ns
total = sum(v[1] for v in tot.values())
print(f'# {sys.argv[1]}: {sum(v[0] for v in tot.values())} launches, {total / 1e6:.3f} ms of kernel time (cold-cache, serialised under ncu)')
print(f'{"kernel":90s} {"launches":>8s} {"ms":>10s} {"share":>7s}')
for name, (n, ns) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f'{name:90s} {n:8d} {ns / 1e6:10.3f} {100 * ns / total:6.1f}%')
