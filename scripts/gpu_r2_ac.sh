#!/bin/bash
mkdir -p gpurun_out
for v in "LFI_GEMM_NARROW=1" "LFI_GEMM_NARROW=0" "LFI_ENC_STREAMS=0" "LFI_WGRAD_STREAM=0"; do
env $v timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/bench_ac.log 2>&1
python -c "
import json;d=json.loads([l for l in open('gpurun_out/bench_ac.log') if l.startswith('{')][-1]);print('$v ms/step', d['ms_per_step'], 'graphed', d.get('graphed_step',{}).get('ms_per_step'))" || tail -5 gpurun_out/bench_ac.log
done
