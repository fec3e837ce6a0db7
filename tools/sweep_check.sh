#!/bin/bash
# parity + speed of every tally variant in build/var: tools/sweep_check.sh
for f in build/var/*.so; do
  t=$(DYNAST_B200_LIB=$PWD/$f python -m pytest tests/test_gpu_tally.py -x -q 2>&1 | tail -1)
  r=$(DYNAST_B200_LIB=$PWD/$f python bench.py --skip-e2e --skip-cpu --skip-em --skip-pipeline --skip-consensus --steps 5 2>gpurun_out/sweep_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['frac'])")
  echo "$f | $t | $r" | tee -a gpurun_out/sweep3.log
done
