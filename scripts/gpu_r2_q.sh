#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_graph.py -m gpu -q --timeout 600 -x > gpurun_out/pytest_q.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|error' gpurun_out/pytest_q.log | tail -12
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_q.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_q.log') if l.startswith('{')][-1]);print('ms/step', d['ms_per_step'], 'graphed', d.get('graphed_step'))" || tail -5 gpurun_out/bench_q.log
