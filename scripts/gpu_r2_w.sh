#!/bin/bash
# refresh the per-step launch list + DRAM traffic of the default path (after the fused input projection)
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/step_traffic_default.csv python bench.py --ncu-step > gpurun_out/ncu_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/step_traffic_default.csv gpurun_out/r02_step_traffic_default.json
python scripts/ncu_summary.py gpurun_out/step_traffic_default.csv > gpurun_out/r02_bf16x3_default_launch_summary.txt
head -14 gpurun_out/r02_bf16x3_default_launch_summary.txt
