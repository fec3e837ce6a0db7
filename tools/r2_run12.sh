#!/bin/bash
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate|lz_kernel|index_kernel|parse_kernel|pack_kernel" -c 60 --csv --log-file gpurun_out/r2_inflate2_launches.csv python tools/bam_profile.py 16000000 > /dev/null 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_inflate2_launches.csv')) if len(r)>10]
h=rows[0]; k=h.index('Kernel Name'); v=h.index('Metric Value'); g=h.index('Grid Size')
for r in rows[1:40]:
    print(r[k].split('(')[0][-28:], r[g], float(r[v].replace(',',''))/1e6, 'ms')
PY
