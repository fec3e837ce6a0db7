#!/bin/bash
DYN_BAM_DEV_TIMING=1 python tools/ingest_bench2.py 8000000 6 256 > gpurun_out/r2_ingest_b.txt 2>&1
grep -v "^dyn_bam_dev_next" gpurun_out/r2_ingest_b.txt | tail -8; grep "^dyn_bam_dev_next" gpurun_out/r2_ingest_b.txt | head -12
ncu --set full --clock-control none --import-source on -k regex:inflate_kernel -c 1 -o gpurun_out/r2_inflate_v1 -f python tools/ingest_bench2.py 2000000 6 256 > gpurun_out/r2_inflate_ncu.log 2>&1
