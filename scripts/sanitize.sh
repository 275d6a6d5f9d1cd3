#!/bin/bash
# compute-sanitizer over the hot path at a small shape: smoke() = final_model.yaml shapes, B=128, T=26, bf16x3:
# tcgen05 GEMMs (TMA, mbarrier rings, CTA pairs), persistent window-GRU kernels (DSMEM, cluster mbarriers), stage-pipelined flow
# core (release/acquire flags, DSMEM, TMEM), the row-owned inverse kernel (cp.async.bulk ring), fused clip+Adam.
# Usage: bash scripts/sanitize.sh   (on a GPU box; summaries land in gpurun_out/sanitize_*.txt)
mkdir -p gpurun_out
CS=/usr/local/cuda/bin/compute-sanitizer
for tool in memcheck synccheck racecheck initcheck; do
  extra=""
  [ "$tool" = "memcheck" ] && extra="--leak-check no"
  timeout ${SAN_TIMEOUT:-900} $CS --tool $tool $extra --print-limit 20 --error-exitcode 0 \
    python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "== $tool exit $?" | tee gpurun_out/sanitize_$tool.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|smoke ok|Error|hazard" gpurun_out/sanitize_$tool.log | sort | uniq -c | sort -rn | head -15 | tee -a gpurun_out/sanitize_$tool.txt
done
