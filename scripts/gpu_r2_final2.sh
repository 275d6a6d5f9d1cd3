#!/bin/bash
# final state of round 2: full GPU suite, smoke, default bench line
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_final2.log 2>&1
grep -E 'passed|failed|FAILED|ERROR' gpurun_out/pytest_final2.log | tail -5
timeout 600 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1500 python bench.py > gpurun_out/bench_final2.log 2>&1
grep '^{' gpurun_out/bench_final2.log | tail -1 > gpurun_out/r02_bench_1gpu.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_1gpu.json'))
print('ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], 'res', d['e2e_resident'].get('value'), 'graphed', d.get('graphed_step',{}).get('ms_per_step'))
print('sample', d['sample']['value'], 'bf16', d['bf16_mode']['ms_per_step'], 'bwd16', d['bf16_backward_mode']['ms_per_step'], 'cpu', d['cpu_baseline']['value'], 'clocks', d['clocks'])
print('roofline', d['roofline']['frac'], d['roofline']['step_traffic'])
PY
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_final2.log 2>&1
grep '^{' gpurun_out/bench_ref_final2.log | tail -1 | cut -c1-400
