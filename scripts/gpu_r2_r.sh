#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:core_inv_rows" -s 3 -c 1 \
  -f -o gpurun_out/prof_inv_rows python scripts/sample_profile.py --frames 8 > gpurun_out/ncu_inv.log 2>&1
echo "ncu exit $?"; tail -2 gpurun_out/ncu_inv.log
