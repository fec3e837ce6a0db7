#!/bin/bash
python -m pytest tests/test_gpu_bamdevice.py tests/test_gpu_from_bam.py -x -q 2>&1 | tail -4
DYN_BAM_DEV_TIMING=1 python tools/ingest_bench2.py 12000000 6 512 > gpurun_out/r2_ingest_c.txt 2>&1
grep -v "^dyn_bam_dev_next" gpurun_out/r2_ingest_c.txt | tail -8; grep "^dyn_bam_dev_next" gpurun_out/r2_ingest_c.txt | head -8
ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -c 1 -o gpurun_out/r2_inflate_v2 -f python tools/ingest_bench2.py 4000000 6 512 > gpurun_out/r2_inflate_ncu2.log 2>&1
