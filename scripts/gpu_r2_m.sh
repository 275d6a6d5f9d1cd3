#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/wide_launches.csv python bench.py --variant wide-lstm --gemm bf16 --ncu-step > gpurun_out/ncu_wide.log 2>&1
tail -3 gpurun_out/ncu_wide.log
python scripts/ncu_summary.py gpurun_out/wide_launches.csv | head -25
