#!/usr/bin/env python
"""Per-source-line summary of an .ncu-rep (needs -lineinfo + --import-source on):
   python tools/ncu_lines.py gpurun_out/prof.ncu-rep [top_n]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
    out = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--csv', '--print-source', 'cuda,sass'],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = None
    fname = ''
    lines = []
    for r in rows:
        if not r:
            continue
        if r[0] == 'File Path':
            fname = r[1].split('/')[-1]
            continue
        if r[0] == 'Function Name':
            continue
        if r[0] == 'Line No':
            hdr = r
            continue
        if hdr is None or r[0] == '':
            continue
        d = dict(zip(hdr[4:], r[4:]))
        try:
            inst = int(d['Instructions Executed'])
            tinst = int(d['Thread Instructions Executed'])
            samples = int(d['# Samples'])
        except (KeyError, ValueError):
            continue
        lines.append((inst, tinst, samples, fname, r[0], r[1].strip()[:110]))
    tot_i = sum(x[0] for x in lines) or 1
    tot_s = sum(x[2] for x in lines) or 1
    print(f'total warp-inst {tot_i}  samples {tot_s}')
    print('  inst%  smpl%  thr/inst  file:line  source')
    for inst, tinst, samples, f, ln, src in sorted(lines, key=lambda x: -x[2])[:top]:
        print(f'{100*inst/tot_i:6.2f} {100*samples/tot_s:6.2f} {tinst/max(inst,1):6.1f}  {f}:{ln}  {src}')


if __name__ == '__main__':
    main()
