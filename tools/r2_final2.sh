#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_exchange.py tests/test_gpu_from_bam.py -x -q -rs 2>&1 | tail -4 > gpurun_out/r2_2gpu_test.log; cat gpurun_out/r2_2gpu_test.log
N=${1:-2}
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $2 > gpurun_out/r2_bench_${N}gpu.json 2> gpurun_out/r2_bench_${N}gpu.err
tail -3 gpurun_out/r2_bench_${N}gpu.err
python - <<PY
import json
t=open('gpurun_out/r2_bench_${N}gpu.json').read()
j=json.loads(t[t.index('{"metric'):])
print('value %.3e' % j['value'], 'e2e', j['e2e'] and j['e2e'].get('value'), 'bit_equal', j['multi_gpu_bit_equal'])
print('e2e_bam', j['e2e_bam'] and {k: j['e2e_bam'].get(k) for k in ('value', 's', 'rep_s', 'ingest', 'sharding', 'records')})
print('pipeline', j['pipeline'] and (j['pipeline']['ms'], j['pipeline']['stage_ms'], j['pipeline']['exchange']))
print('em', j['em'] and (j['em']['value'], j['em']['ms']))
print('cons', j['consensus'] and (j['consensus']['value'], j['consensus']['c3_value'])); print('c5', j.get('c5'))
PY
