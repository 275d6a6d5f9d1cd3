#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_paths.py tests/test_gpu_parity.py -m gpu -q --timeout 900 -x > gpurun_out/pytest_n.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_n.log | tail -12
for v in wide-lstm wide-gru; do
timeout 900 python bench.py --variant $v --gemm bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_$v.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_$v.log') if l.startswith('{')][-1]);print('$v bf16: ms/step', d['ms_per_step'], 'frames/s', d['value'])" || tail -3 gpurun_out/bench_$v.log
done
