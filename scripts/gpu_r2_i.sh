#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches_sample.csv \
  python scripts/sample_profile.py --frames 48 > gpurun_out/ncu_sample.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_sample.csv > gpurun_out/launch_summary_sample.txt 2>&1
head -16 gpurun_out/launch_summary_sample.txt
