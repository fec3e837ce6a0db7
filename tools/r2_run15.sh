#!/bin/bash
for cfg in "3 4" "3 8" "3 16" "12 8" "12 16" "6 8"; do
  set -- $cfg
  echo "== DYN_INF_WARPS=$1 DYN_BAM_READ_THREADS=$2"
  DYN_INF_WARPS=$1 DYN_BAM_READ_THREADS=$2 DYN_BAM_DEV_TIMING=1 timeout 600 python tools/bam_profile.py 16000000 2>&1 | grep -E "wait for the reader [1-9]|tally_bam|ingest only" | tail -5
done > gpurun_out/r2_inflate_cfg.txt 2>&1
cat gpurun_out/r2_inflate_cfg.txt
