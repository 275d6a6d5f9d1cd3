#!/bin/bash
mkdir -p gpurun_out
for bs in 256 512 1024 2048; do
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bf16 --sample-seqs $bs --sample-frames 200 > gpurun_out/bench_s$bs.log 2>&1
python -c "import json;d=json.loads([l for l in open('gpurun_out/bench_s$bs.log') if l.startswith('{')][-1]);s=d['sample'];print('seqs $bs: ms', s['ms'], 'frames/s', s['value'], 'us/frame', s['roofline']['us_per_frame'])" || tail -3 gpurun_out/bench_s$bs.log
done
