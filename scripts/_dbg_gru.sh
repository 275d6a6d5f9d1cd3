for v in 0 1 2 4 8 16 31; do
  echo "== LFI_DBG_GRU=$v"
  LFI_DBG_GRU=$v timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/dbg_$v.csv python bench.py --ncu-step > /dev/null 2>&1
  python scripts/ncu_summary.py gpurun_out/dbg_$v.csv | grep "gemm_tc_kernel<1"
done
