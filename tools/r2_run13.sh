#!/bin/bash
timeout 1200 ncu --set full --clock-control none --import-source on -k regex:inflate_lanes -c 1 -o gpurun_out/r2_inflate_lanes_v1 -f python tools/bam_profile.py 8000000 > gpurun_out/r2_inflate_lanes_ncu.log 2>&1
tail -3 gpurun_out/r2_inflate_lanes_ncu.log
ls -la gpurun_out/r2_inflate_lanes_v1.ncu-rep
