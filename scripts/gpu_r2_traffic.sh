#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --profile-from-start off --csv \
  --log-file gpurun_out/step_traffic.csv python bench.py --ncu-step > gpurun_out/ncu_traffic.log 2>&1
python scripts/ncu_traffic.py gpurun_out/step_traffic.csv gpurun_out/r02_step_traffic.json
# calibration of the tensor-pipe counter on cuBLAS (the peak the roofline divides by): bf16 8192^3
cat > /tmp/cublas_cal.py <<'PY'
import torch
a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16); b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
for _ in range(3): torch.matmul(a, b)
torch.cuda.synchronize(); torch.cuda.profiler.start(); torch.matmul(a, b); torch.cuda.synchronize(); torch.cuda.profiler.stop()
PY
timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.max,sm__inst_executed_pipe_tensor.sum \
  --clock-control none --profile-from-start off --csv --log-file gpurun_out/cublas_cal.csv python /tmp/cublas_cal.py > /dev/null 2>&1
grep -v "^==" gpurun_out/cublas_cal.csv | cut -d, -f5,13,14,15 | head -8
