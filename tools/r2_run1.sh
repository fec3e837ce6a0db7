#!/bin/bash
# GPU tests, bench lines with both record layouts, ncu capture of the lean tally
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gputest_2.log
python bench.py --skip-consensus > gpurun_out/r2_bench_lean_a.json 2> gpurun_out/r2_bench_lean_a.err
python bench.py --full-records --skip-consensus --skip-em --skip-cpu > gpurun_out/r2_bench_full_a.json 2> gpurun_out/r2_bench_full_a.err
ncu --set full --clock-control none --import-source on -k tally_kernel -c 1 -o gpurun_out/r2_tally_lean -f \
  python bench.py --steps 1 --warmup 1 --reads 4000000 --skip-e2e --skip-cpu --skip-em --skip-pipeline --skip-consensus > gpurun_out/r2_ncu_lean.log 2>&1
tail -3 gpurun_out/r2_gputest_2.log; cut -c1-600 gpurun_out/r2_bench_lean_a.json; tail -2 gpurun_out/r2_bench_lean_a.err
