#!/bin/bash
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1500 $CS --tool memcheck --leak-check no --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_paths.py -m gpu -q --timeout 1200 -k "persistent_encoder and 100-bf16x3" > gpurun_out/sanitize_memcheck_ragged.log 2>&1
grep -E "ERROR SUMMARY|passed|failed|Invalid|Error" gpurun_out/sanitize_memcheck_ragged.log | sort | uniq -c | head
timeout 1500 $CS --tool racecheck --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_paths.py -m gpu -q --timeout 1200 -k "persistent_encoder and 100-bf16x3" > gpurun_out/sanitize_racecheck_ragged.log 2>&1
grep -E "RACECHECK SUMMARY|passed|failed|Race reported|hazard" gpurun_out/sanitize_racecheck_ragged.log | cut -c1-160 | sort | uniq -c | head
for v in wide-lstm wide-gru; do
timeout 900 python bench.py --variant $v --gemm bf16 --steps 5 --warmup 3 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_$v.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_$v.log') if l.startswith('{')][-1]);print('$v bf16: ms/step', d['ms_per_step'], 'frames/s', d['value'])"
done
