#!/bin/bash
python tools/ingest_bench2.py 8000000 6 256 > gpurun_out/r2_ingest_a.txt 2>&1
cat gpurun_out/r2_ingest_a.txt | tail -12
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_ingest_launches.csv python tools/ingest_bench2.py 2000000 6 256 > gpurun_out/r2_ingest_ncu.log 2>&1
