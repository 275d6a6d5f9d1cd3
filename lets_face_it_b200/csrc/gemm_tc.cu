// tcgen05 / TMEM / TMA GEMM (placeholder until the tensor-core tiles land: reports "not handled").
#include "lfi_common.cuh"
namespace lfi {
size_t gemm_tc_ws_bytes() { return 0; }
int gemm_tc(int mode, const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st, bool *handled) {
  (void)mode; (void)g; (void)ws; (void)ws_bytes; (void)st;
  *handled = false;
  return LFI_OK;
}
}  // namespace lfi
