#!/bin/bash
python tools/cons_prof.py 4000000 > gpurun_out/r2_cons_count.txt 2>&1; cat gpurun_out/r2_cons_count.txt
python -m pytest tests/test_gpu_consensus.py tests/test_consensus_operator.py tests/test_gpu_tally.py -x -q -m gpu 2>&1 | tail -3
python bench.py --skip-bam --skip-e2e --skip-cpu --skip-em --skip-other-layout --skip-pipeline 2>/dev/null | python -c "
import json,sys
t=sys.stdin.read(); d=json.loads(t[t.index('{'):]); print(d['consensus'])"
