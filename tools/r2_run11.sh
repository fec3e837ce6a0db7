#!/bin/bash
# lane-per-block inflate: tests first (bounded by timeout), then the from-BAM profile with each kernel
timeout 600 python -m pytest tests/test_gpu_bamdevice.py tests/test_gpu_from_bam.py -x -q -m gpu 2>&1 | tail -4
for mode in lane warp; do
  echo "== DYN_INFLATE=$mode"
  DYN_INFLATE=$mode DYN_BAM_DEV_TIMING=1 timeout 600 python tools/bam_profile.py 16000000 2>&1 | grep -v "^$" | tail -14
done > gpurun_out/r2_inflate_lane_vs_warp.txt 2>&1
cat gpurun_out/r2_inflate_lane_vs_warp.txt
