#!/bin/bash
# final weak-scaling run of round 2 on one 8-GPU box: N = 1, 2, 4, 8 (training + sampling), wide variant at 8
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv | head -3
timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 > gpurun_out/scale_n1.log 2>&1
for n in 2 4 8; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 20 --warmup 5 --no-cpu-baseline --no-bf16 > gpurun_out/scale_n$n.log 2>&1
done
for n in 1 2 4 8; do
grep '^{' gpurun_out/scale_n$n.log | tail -1 > gpurun_out/r02_scale_n$n.json
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/r02_scale_n$n.json'))
    print('N=$n ms/step', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'sample', round(d.get('sample',{}).get('value',0)))
except Exception as e:
    print('N=$n FAILED', e)
PY
done
for v in wide-lstm wide-gru; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29631 bench.py --gpus 8 --variant $v --gemm bf16 --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 --no-sample > gpurun_out/scale_$v.log 2>&1
grep '^{' gpurun_out/scale_$v.log | tail -1 > gpurun_out/r02_8gpu_$v.json
python -c "import json;d=json.load(open('gpurun_out/r02_8gpu_$v.json'));print('$v x8: ms/step', d['ms_per_step'], 'value', d['value'])" || tail -3 gpurun_out/scale_$v.log
done
