#!/bin/bash
python -m pytest tests/test_gpu_bamdevice.py tests/test_gpu_from_bam.py -x -q 2>&1 | tail -3
python bench.py --skip-consensus --skip-em --skip-cpu --skip-pipeline --skip-e2e > gpurun_out/r2_bench_bam_a.json 2> gpurun_out/r2_bench_bam_a.err
tail -5 gpurun_out/r2_bench_bam_a.err
python -c "
import json
j=json.load(open('gpurun_out/r2_bench_bam_a.json'))
print(json.dumps(j['e2e_bam'], indent=1))
"
