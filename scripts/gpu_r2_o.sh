#!/bin/bash
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_o.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_o.log | tail -12
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/wide_launches.csv python bench.py --variant wide-lstm --gemm bf16 --ncu-step > gpurun_out/ncu_wide.log 2>&1
python scripts/ncu_summary.py gpurun_out/wide_launches.csv > gpurun_out/r02_wide_lstm_launch_summary.txt; head -8 gpurun_out/r02_wide_lstm_launch_summary.txt
