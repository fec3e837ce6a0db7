#!/bin/bash
# One GPU round: tests, bench, launch list, one full ncu capture of the tally. Usage: tools/gpu_round.sh <tag>
tag=${1:-x}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_$tag.log 2>&1; echo "pytest rc=$?" | tee -a gpurun_out/pytest_$tag.log
tail -5 gpurun_out/pytest_$tag.log
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
cat gpurun_out/bench_$tag.json
B="python bench.py --reads 4000000 --steps 2 --warmup 1 --skip-e2e --skip-cpu --skip-pipeline --em-barcodes 2000"
$B > gpurun_out/plain_$tag.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_$tag.csv $B > gpurun_out/ncu_l_$tag.log 2>&1
$B > gpurun_out/plain2_$tag.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tally_kernel -s 3 -c 1 -f -o gpurun_out/prof_tally_$tag $B > gpurun_out/ncu_t_$tag.log 2>&1
echo done
