#!/bin/bash
# one GPU: full GPU test suite, the default bench line, C5 line + its launch list
python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --c5 --skip-bam --skip-e2e --skip-cpu --skip-em --skip-consensus --skip-other-layout --skip-pipeline > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err; tail -2 gpurun_out/r2_bench_c5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_c5_launches.csv python bench.py --c5 --skip-bam --skip-e2e --skip-cpu --skip-em --skip-consensus --skip-other-layout --skip-pipeline --steps 1 > gpurun_out/r2_c5_ncu.log 2>&1
tail -c 300 gpurun_out/r2_c5_ncu.log
