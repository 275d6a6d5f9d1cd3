#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pinned.py tests/test_gpu_paths.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_v.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|error' gpurun_out/pytest_v.log | tail -8
for v in 1 0 1 0; do
LFI_ENC_XFUSE=$v timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_x$v.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_x$v.log') if l.startswith('{')][-1]);print('XFUSE=$v ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'graphed', d.get('graphed_step',{}).get('ms_per_step', d.get('graphed_step')))" || tail -5 gpurun_out/bench_x$v.log
done
