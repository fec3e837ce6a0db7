#!/bin/bash
python tools/cons_prof.py 4000000 > gpurun_out/r2_cons_count.txt 2>&1; cat gpurun_out/r2_cons_count.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_cons_count_launches.csv python tools/cons_prof.py 4000000 > /dev/null 2>&1
python tools/launch_summary.py gpurun_out/r2_cons_count_launches.csv synth | head -12
python bench.py --c5 --skip-bam --skip-e2e --skip-cpu --skip-em --skip-consensus --skip-other-layout --skip-pipeline 2>/dev/null | python -c "
import json,sys
t=sys.stdin.read(); d=json.loads(t[t.index('{'):]); print(d['c5'])"
