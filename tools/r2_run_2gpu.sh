#!/bin/bash
# two GPUs: the 2-GPU parity test, then the bench at N = 2
python -m pytest tests/test_gpu_exchange.py -x -q -rs 2>&1 | tail -8 > gpurun_out/r2_2gpu_test.log
cat gpurun_out/r2_2gpu_test.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --skip-consensus > gpurun_out/r2_bench_2gpu.json 2> gpurun_out/r2_bench_2gpu.err
tail -5 gpurun_out/r2_bench_2gpu.err
python - <<'PY'
import json
j=json.load(open('gpurun_out/r2_bench_2gpu.json'))
print('value %.3e' % j['value'], 'e2e', j['e2e'] and '%.3e' % j['e2e']['value'], 'bit_equal', j['multi_gpu_bit_equal'])
print('e2e_bam', j['e2e_bam'] and {k: j['e2e_bam'][k] for k in ('value', 's', 'rep_s', 'ingest', 'sharding')})
print('pipeline', j['pipeline'] and (j['pipeline']['ms'], j['pipeline']['stage_ms']))
print('em', j['em'] and (j['em']['value'], j['em']['ms']))
PY
