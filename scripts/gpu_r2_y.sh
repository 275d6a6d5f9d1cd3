#!/bin/bash
mkdir -p gpurun_out
for v in 0 1 0 1; do
LFI_GRU_TILE32=$v timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_t$v.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_t$v.log') if l.startswith('{')][-1]);print('TILE32=$v ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'graphed', d.get('graphed_step',{}).get('ms_per_step', d.get('graphed_step')))" || tail -5 gpurun_out/bench_t$v.log
done
LFI_GRU_TILE32=1 timeout 900 python -m pytest tests/test_gpu_pinned.py -m gpu -q --timeout 600 -x -k "every_mode or benchmarked" 2>&1 | tail -2
