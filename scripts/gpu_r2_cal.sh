#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,sm__cycles_elapsed.max,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_tensor.sum,dram__bytes_read.sum,dram__bytes_write.sum \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/gemm_cal.csv python scripts/gemm_cal.py > gpurun_out/gemm_cal.log 2>&1
grep -v "^==" gpurun_out/gemm_cal.csv | python -c "
import csv,sys
for r in csv.DictReader(sys.stdin):
    print(r['Kernel Name'][:48].ljust(48), r['Metric Name'].ljust(70), r['Metric Unit'].ljust(8), r['Metric Value'])
"
