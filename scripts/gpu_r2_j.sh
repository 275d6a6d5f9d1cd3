#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 600 -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_gpu.log | tail -12
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_j.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/bench_j.log').read().strip().splitlines()[-1]);print('default ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
LFI_ENC_NO_H32=0 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_j2.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/bench_j2.log').read().strip().splitlines()[-1]);print('with fp32 h copy ms/step', d['ms_per_step'])"
python scripts/step_phases.py 2>&1 | tail -1
