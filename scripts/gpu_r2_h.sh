#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -k "wide_variant or resident or sampl or inference or long_horizon or full_kat_final or seq_paths" > gpurun_out/pytest_h.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_h.log | tail -20
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/bench.log 2>&1
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print('step ms', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'], 'resident', d.get('e2e_resident'))
print('sample', d['sample']['value'], d['sample']['ms'], d['sample']['gpu_launches'], d['sample']['roofline'])
PY
LFI_SAMPLE_GRAPH=0 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-bf16 > gpurun_out/bench_nograph.log 2>&1
python -c "import json;d=json.loads(open('gpurun_out/bench_nograph.log').read().strip().splitlines()[-1]);print('sample no graph', d['sample']['value'], d['sample']['ms'])"
