#!/bin/bash
for v in q8r8 q16r8 q24r8 q8r16 q12w128; do
  if [ $v = base ]; then lib=dynast-release_b200/libdynast_b200.so; else lib=build/variants/lib_$v.so; fi
  echo "== $v"
  DYNAST_B200_LIB=$PWD/$lib DYN_BAM_DEV_TIMING=1 timeout 600 python tools/bam_profile.py 16000000 2>&1 | grep -E "on the stream|tally_bam" | tail -5
done > gpurun_out/r2_inflate_lut_bits.txt 2>&1
cat gpurun_out/r2_inflate_lut_bits.txt
