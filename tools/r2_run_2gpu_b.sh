#!/bin/bash
# two GPUs: the 2-GPU parity test (fused push exchange vs all-to-all), then the pipeline leg at N = 2 with each transport
timeout 600 python -m pytest tests/test_gpu_exchange.py -x -q -rs 2>&1 | tail -15 > gpurun_out/r2_2gpu_test_b.log
cat gpurun_out/r2_2gpu_test_b.log
for mode in push nccl; do
  DYN_EXCHANGE=$mode timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --skip-consensus --skip-e2e --skip-cpu --skip-em --skip-bam --skip-other-layout > gpurun_out/r2_bench_2gpu_$mode.json 2> gpurun_out/r2_bench_2gpu_$mode.err
  tail -3 gpurun_out/r2_bench_2gpu_$mode.err
  python - <<PY
import json
t=open('gpurun_out/r2_bench_2gpu_$mode.json').read()
j=json.loads(t[t.index('{'):])
print('$mode', 'bit_equal', j['multi_gpu_bit_equal'], 'pipeline', j['pipeline'] and (j['pipeline']['ms'], j['pipeline']['stage_ms']))
PY
done
