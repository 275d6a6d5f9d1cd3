#!/bin/bash
mkdir -p gpurun_out
LFI_ENC_TIMING=1 timeout 300 python scripts/step_phases.py > gpurun_out/phases.log 2>&1
grep -v "^ *$" gpurun_out/phases.log | awk '{k=$1" "$2" "$3" "$4" "$5" "$6" "$7; if(!(k in s)){s[k]=1; print}}' | head -20
timeout 600 python -m pytest tests/test_gpu_paths.py -m gpu -q --timeout 300 -k "persistent_encoder" > gpurun_out/pytest_enc.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|timed out' gpurun_out/pytest_enc.log | tail -8
timeout 600 python -m pytest tests/test_gpu_postprocess.py -m gpu -q --timeout 300 > gpurun_out/pytest_post.log 2>&1
grep -E 'passed|failed|FAILED|ERROR|assert|Error|timed out' gpurun_out/pytest_post.log | tail -8
