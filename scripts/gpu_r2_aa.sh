#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_gemm_tc.py tests/test_gpu_pinned.py tests/test_gpu_paths.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_aa.log 2>&1
grep -E 'passed|failed|FAILED|ERROR' gpurun_out/pytest_aa.log | tail -3
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_aa.csv python bench.py --ncu-step > gpurun_out/ncu_aa.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_aa.csv > gpurun_out/launches_aa.txt; head -9 gpurun_out/launches_aa.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_aa.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_aa.log') if l.startswith('{')][-1]);print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'graphed', d.get('graphed_step',{}).get('ms_per_step'), 'gemm ms', d['roofline']['ms_per_launch'])"
