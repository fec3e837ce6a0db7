#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_exchange.py -x -q -rs 2>&1 | tail -6
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 900 python bench.py --skip-bam --skip-e2e --skip-cpu --skip-em --skip-consensus --skip-other-layout 2>/dev/null | python -c "
import json,sys
t=sys.stdin.read(); d=json.loads(t[t.index('{'):]); print(d['pipeline']['ms'], d['pipeline'].get('pi_g'))"
