#!/bin/bash
mkdir -p gpurun_out
for d in 1 17 21; do
echo "== LFI_ENC_TIMING=$d"
LFI_ENC_TIMING=$d timeout 300 python scripts/step_phases.py 2>&1 | grep "fwd E=256 hist=24\|forward" | awk '{k=$1" "$2" "$3" "$4" "$5" "$6" "$7; if(!(k in s)){s[k]=1; print}}' | cut -c1-160
done
