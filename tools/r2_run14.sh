#!/bin/bash
timeout 300 python -m pytest tests/test_gpu_bamdevice.py tests/test_gpu_from_bam.py -x -q -m gpu 2>&1 | tail -2
for w in 3 8 12 16 24; do
  echo "== DYN_INF_WARPS=$w"
  DYN_INF_WARPS=$w DYN_BAM_DEV_TIMING=1 timeout 600 python tools/bam_profile.py 16000000 2>&1 | grep -v "^$" | tail -5
done > gpurun_out/r2_inflate_warps.txt 2>&1
cat gpurun_out/r2_inflate_warps.txt
