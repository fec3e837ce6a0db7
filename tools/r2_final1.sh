#!/bin/bash
# final 1-GPU evidence: tests, bench line, C5 line, launch lists (bench pipeline leg; from-BAM path), one ncu --set full of the Huffman pass
python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2_pytest_gpu.log; cat gpurun_out/r2_pytest_gpu.log
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
python bench.py --c5 --skip-bam --skip-e2e --skip-cpu --skip-em --skip-consensus --skip-other-layout --skip-pipeline > gpurun_out/r2_bench_c5.json 2> gpurun_out/r2_bench_c5.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2_pipeline_launches.csv python bench.py --skip-consensus --skip-em --skip-cpu --skip-e2e --skip-bam --skip-other-layout --steps 1 --warmup 3 > gpurun_out/r2_pipe_ncu.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"inflate|lz_kernel|index_kernel|parse_kernel|pack_kernel|tally" -c 80 --csv --log-file gpurun_out/r2_ingest_launches.csv python tools/bam_profile.py 16000000 > /dev/null 2>&1
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflate_lanes -c 1 -o gpurun_out/r2_inflate_lanes_final -f python tools/bam_profile.py 16000000 > gpurun_out/r2_inflate_lanes_ncu.log 2>&1
ls -la gpurun_out/r2_inflate_lanes_final.ncu-rep
