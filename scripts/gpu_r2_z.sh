#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pinned.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_z.log 2>&1
grep -E 'passed|failed|FAILED|ERROR' gpurun_out/pytest_z.log | tail -3
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/launches_z.csv python bench.py --ncu-step > gpurun_out/ncu_z.log 2>&1
python scripts/ncu_summary.py gpurun_out/launches_z.csv | head -9
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_z.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_z.log') if l.startswith('{')][-1]);print('ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'graphed', d.get('graphed_step',{}).get('ms_per_step'))"
