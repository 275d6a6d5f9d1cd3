#!/bin/bash
# Several ncu --set full captures of one training step (bench.py --ncu-step), one report per kernel family.
# usage: gpu_ncu_multi.sh "name|regex|skip|count" ...
mkdir -p gpurun_out
for spec in "$@"; do
  IFS='|' read -r name rx skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:${rx}" -s ${skip} -c ${cnt} \
    -f -o gpurun_out/${name} python bench.py --ncu-step ${BENCH_ARGS} > gpurun_out/ncu_${name}.log 2>&1
  echo "${name}: ncu exit $?"; tail -2 gpurun_out/ncu_${name}.log
done
ls -la gpurun_out/*.ncu-rep
