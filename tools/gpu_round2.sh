#!/bin/bash
# tests + launch list of the whole device pipeline.  Usage: tools/gpu_round2.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_$tag.log
B="python bench.py --reads 4000000 --steps 2 --warmup 3 --skip-e2e --skip-cpu --em-barcodes 4000"
$B > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1
echo done
