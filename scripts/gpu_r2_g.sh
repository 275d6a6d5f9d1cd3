#!/bin/bash
mkdir -p gpurun_out
LFI_CORE_TIMING=1 timeout 300 python scripts/step_phases.py 2>&1 | grep "core_\|forward" | awk '{k=$1" "$2" "$3; if(!(k in s)){s[k]=1; print}}' | cut -c1-400
