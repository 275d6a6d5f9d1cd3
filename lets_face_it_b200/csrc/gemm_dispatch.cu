// Routes a GEMM to the fp32 FFMA tiles or to the tcgen05 tiles (gemm_tc.cu) according to lfi_gemm_mode.
#include "lfi_common.cuh"

namespace lfi {
int gemm_tc(int mode, const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st, bool *handled);
size_t gemm_tc_ws_bytes();

int gemm_dispatch(int mode, const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (mode == LFI_GEMM_FP32) return gemm_simt(g, st);
  LFI_REQUIRE(mode == LFI_GEMM_BF16X3 || mode == LFI_GEMM_BF16, LFI_ERR_ARG, "unknown gemm mode %d", mode);
  bool handled = false;
  LFI_TRY(gemm_tc(mode, g, ws, ws_bytes, st, &handled));
  if (handled) return LFI_OK;
  return gemm_simt(g, st);  // shapes the tensor-core tiles do not cover (tiny K / N): exact fp32 tiles
}
}  // namespace lfi

extern "C" size_t lfi_gemm_ws_bytes(void) { return lfi::gemm_tc_ws_bytes(); }
