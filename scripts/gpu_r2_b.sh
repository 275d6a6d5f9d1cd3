#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_paths.py -m gpu -q --timeout 300 -k "persistent_encoder" > gpurun_out/pytest_enc.log 2>&1
echo "pytest enc exit $?" >> gpurun_out/pytest_enc.log
grep -E 'passed|failed|FAILED|ERROR|assert|Error|timed out' gpurun_out/pytest_enc.log | tail -20
timeout 600 python -m pytest tests/test_gpu_gemm_tc.py -m gpu -q --timeout 300 -k "final_model_in_tensor" > gpurun_out/pytest_tc.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|timed out' gpurun_out/pytest_tc.log | tail -10
LFI_ENC_TIMING=1 timeout 300 python scripts/step_phases.py > gpurun_out/phases.log 2>&1
sort gpurun_out/phases.log | uniq -c | sort -rn | head -20
LFI_ENC_PERSIST=0 timeout 300 python scripts/step_phases.py 2>&1 | tail -1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/launches.csv \
  python bench.py --ncu-step > gpurun_out/ncu_bench.log 2>&1
echo "ncu exit $?"
python scripts/ncu_summary.py gpurun_out/launches.csv > gpurun_out/launch_summary.txt 2>&1
head -40 gpurun_out/launch_summary.txt
timeout 900 python -m pytest tests/test_gpu_pinned.py -m gpu -q --timeout 600 -k "long_horizon or three_optimizer or state_dict or benchmarked" > gpurun_out/pytest_pinned.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|timed out' gpurun_out/pytest_pinned.log | tail -30
