#!/bin/bash
for v in base t32c20 t32c24 t32c32 t64s4k; do
  if [ $v = base ]; then lib=dynast-release_b200/libdynast_b200.so; else lib=build/variants/lib_$v.so; fi
  DYNAST_B200_LIB=$PWD/$lib python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam --skip-pipeline --reads 30000000 2>/dev/null | python -c "
import json,sys
t=sys.stdin.read(); j=json.loads(t[t.index('{'):]); print('$v', 'lean %.3e reads/s' % j['value'], 'frac %.3f' % j['roofline']['frac'], 'full %.3e' % j['roofline']['other_layout']['reads_per_s'], 'frac %.3f' % j['roofline']['other_layout']['frac'])
"
done | tee gpurun_out/r2_tally_sweep2.txt
python -m pytest tests/test_gpu_tally.py -x -q -m gpu 2>&1 | tail -2
