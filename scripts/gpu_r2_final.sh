#!/bin/bash
# round 2 final visit: full GPU suite, smoke, bench (default), traffic of the default path, launch list
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_gpu.log | tail -20
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/step_traffic_default.csv python bench.py --ncu-step > gpurun_out/ncu_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/step_traffic_default.csv gpurun_out/r02_step_traffic_default.json | head -14
python scripts/ncu_summary.py gpurun_out/step_traffic_default.csv > gpurun_out/r02_default_launch_summary.txt 2>&1
timeout 1500 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_final.log 2>&1
echo "bench exit $?"; tail -1 gpurun_out/bench_final.log | cut -c1-1500
