#!/bin/bash
# ncu --set full captures of the top kernels of one training step (bench.py --ncu-step brackets ONE step with cudaProfilerStart/Stop)
mkdir -p gpurun_out
K=${1:-"core_fwd_wave|core_bwd_wave"}
SKIP=${2:-40}
CNT=${3:-2}
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${K}" -s ${SKIP} -c ${CNT} \
  -f -o gpurun_out/${4:-prof_core} python bench.py --ncu-step ${BENCH_ARGS} > gpurun_out/ncu_full.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
