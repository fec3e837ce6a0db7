#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -4
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_pipeline_launches.csv python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam --steps 1 --warmup 1 > gpurun_out/r2_pipe_ncu.log 2>&1
python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam > gpurun_out/r2_bench_dedup.json 2> gpurun_out/r2_bench_dedup.err
python -c "
import json
j=json.load(open('gpurun_out/r2_bench_dedup.json'))
print(j['pipeline']['ms'], j['pipeline']['stage_ms'], j['pipeline']['rep_ms'])
"
