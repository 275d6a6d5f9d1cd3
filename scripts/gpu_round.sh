#!/bin/bash
# One GPU-box visit: parity tests, smoke, a short bench, the ncu launch list.  Everything is bounded by `timeout`.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
if [ -z "${SKIP_TESTS}" ]; then
timeout 900 python -m pytest tests -m gpu -q --timeout 600 ${PYTEST_ARGS} > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E 'passed|failed|FAILED|ERROR|assert|Error' gpurun_out/pytest_gpu.log | tail -40
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
fi
timeout 900 python bench.py --steps ${BENCH_STEPS:-5} --warmup 3 ${BENCH_ARGS} > gpurun_out/bench.log 2>&1
echo "bench exit $?" >> gpurun_out/bench.log
tail -5 gpurun_out/bench.log
if [ -n "${NCU_LIST}" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python bench.py --ncu-step ${BENCH_ARGS} > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
head -40 gpurun_out/launch_summary.txt
fi
