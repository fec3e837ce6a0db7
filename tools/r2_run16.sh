#!/bin/bash
timeout 400 python -m pytest tests/test_gpu_bamdevice.py tests/test_gpu_from_bam.py -x -q -m gpu 2>&1 | tail -2
for cfg in "12 16" "8 16" "3 16"; do
  set -- $cfg
  echo "== DYN_INF_WARPS=$1 DYN_BAM_READ_THREADS=$2"
  DYN_INF_WARPS=$1 DYN_BAM_READ_THREADS=$2 DYN_BAM_DEV_TIMING=1 timeout 600 python tools/bam_profile.py 16000000 2>&1 | grep -E "wait for the reader|tally_bam|ingest only|on the stream" | tail -8
done > gpurun_out/r2_inflate_cfg2.txt 2>&1
cat gpurun_out/r2_inflate_cfg2.txt
