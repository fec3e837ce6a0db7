#!/bin/bash
for v in base c16 c18 t128c8 s6c16; do
  if [ $v = base ]; then lib=dynast-release_b200/libdynast_b200.so; else lib=build/variants/lib_$v.so; fi
  DYNAST_B200_LIB=$PWD/$lib python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam --skip-pipeline --reads 30000000 2>/dev/null | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$v', '%.3e reads/s' % j['value'], 'frac %.3f' % j['roofline']['frac'])
"
done | tee gpurun_out/r2_tally_sweep.txt
