#!/bin/bash
for f in build/var/*.so; do
  r=$(DYNAST_B200_LIB=$PWD/$f python bench.py --skip-e2e --skip-cpu --skip-em --skip-pipeline --steps 5 2>gpurun_out/sweep_err.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value']/1e9, d['roofline']['frac'])")
  echo "$f $r" | tee -a gpurun_out/sweep2.log
done
