#!/bin/bash
export DYN_BENCH_IGNORE_STATUS=1
for f in build/var/*.so; do
 for extra in "" "--no-emit"; do
  r=$(DYNAST_B200_LIB=$PWD/$f python bench.py --skip-e2e --skip-cpu --skip-em --steps 5 $extra 2>gpurun_out/sweep_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['frac'])")
  echo "$f $extra $r" | tee -a gpurun_out/sweep.log
 done
done
