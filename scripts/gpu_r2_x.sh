#!/bin/bash
mkdir -p gpurun_out
SAN_TIMEOUT=900 bash scripts/sanitize.sh 2>&1 | tail -30
CS=/usr/local/cuda/bin/compute-sanitizer
timeout 1200 $CS --tool memcheck --leak-check no --print-limit 20 --error-exitcode 0 python -m pytest tests/test_gpu_paths.py -m gpu -q -x --timeout 1000 -k "hybrid" > gpurun_out/sanitize_hybrid_memcheck.log 2>&1
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_hybrid_memcheck.log | tail -3
