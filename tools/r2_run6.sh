#!/bin/bash
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam > gpurun_out/r2_bench_dedup.json 2> gpurun_out/r2_bench_dedup.err
tail -3 gpurun_out/r2_bench_dedup.err
python -c "
import json
j=json.load(open('gpurun_out/r2_bench_dedup.json'))
print(json.dumps(j['pipeline'], indent=1)[:1500])
"
