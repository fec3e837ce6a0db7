#!/usr/bin/env python
"""Instruction / stall-sample share per source region of tally.cu:
   python tools/ncu_regions.py rep.ncu-rep "name:lo-hi" ...   (line ranges in tally.cu; other files are listed by file)"""
import csv
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
regions = []
for a in sys.argv[2:]:
    name, rng = a.split(':')
    lo, hi = rng.split('-')
    regions.append((name, int(lo), int(hi)))
out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                     capture_output=True, text=True).stdout
hdr = None
fname = ''
agg = defaultdict(lambda: [0, 0, 0])
for r in csv.reader(out.splitlines()):
    if not r:
        continue
    if r[0] == 'File Path':
        fname = r[1].split('/')[-1]
        continue
    if r[0] == 'Line No':
        hdr = r
        continue
    if hdr is None or r[0] in ('', 'Function Name'):
        continue
    d = dict(zip(hdr[4:], r[4:]))
    try:
        inst, tinst, smp = int(d['Instructions Executed']), int(d['Thread Instructions Executed']), int(d['# Samples'])
        ln = int(r[0])
    except (KeyError, ValueError):
        continue
    key = fname
    if fname == 'tally.cu':
        key = 'tally.cu:other'
        for name, lo, hi in regions:
            if lo <= ln <= hi:
                key = name
                break
    a = agg[key]
    a[0] += inst; a[1] += tinst; a[2] += smp
ti = sum(a[0] for a in agg.values()) or 1
ts = sum(a[2] for a in agg.values()) or 1
print(f'total warp-inst {ti} samples {ts}')
for k, a in sorted(agg.items(), key=lambda kv: -kv[1][0]):
    print(f'{100*a[0]/ti:6.2f}% inst {100*a[2]/ts:6.2f}% samples  {a[1]/max(a[0],1):5.1f} thr/inst  {k}')
