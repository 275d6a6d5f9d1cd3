#!/bin/bash
# strong scaling at a global batch of 256 sequences (secondary number of SURVEY.md section 8(d) config 3)
mkdir -p gpurun_out
for n in 2 4 8; do
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2971$n bench.py --gpus $n --strong --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/strong_n$n.log 2>&1
grep '^{' gpurun_out/strong_n$n.log | tail -1 > gpurun_out/r02_strong_n$n.json
python -c "import json;d=json.load(open('gpurun_out/r02_strong_n$n.json'));print('N=$n ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), d['scaling'], d['config']['global_batch'])" || tail -3 gpurun_out/strong_n$n.log
done
