// Routes a GEMM to the fp32 FFMA tiles or to the tcgen05 tiles (gemm_tc.cu) according to lfi_gemm_mode.
#include "lfi_common.cuh"

namespace lfi {
int gemm_tc(int mode, const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st, bool *handled);
bool gemm_tc_wants(const GemmArgs &g);
namespace tc { size_t split_ws_bytes(const GemmArgs &g, int nplanes); }

int gemm_dispatch(int mode, const GemmArgs &g, void *ws, size_t ws_bytes, cudaStream_t st) {
  if (mode == LFI_GEMM_FP32) return gemm_simt(g, st);
  LFI_REQUIRE(mode == LFI_GEMM_BF16X3 || mode == LFI_GEMM_BF16, LFI_ERR_ARG, "unknown gemm mode %d", mode);
  bool handled = false;
  LFI_TRY(gemm_tc(mode, g, ws, ws_bytes, st, &handled));
  if (handled) return LFI_OK;
  return gemm_simt(g, st);  // shapes the tensor-core tiles do not cover (tiny K / N): exact fp32 tiles
}

// Operand-plane scratch one GEMM needs in the tensor-core modes (0 in fp32 mode or for shapes routed to the fp32 tiles).
size_t gemm_ws_bytes(int mode, const GemmArgs &g) {
  if (mode == LFI_GEMM_FP32 || !gemm_tc_wants(g)) return 0;
  return tc::split_ws_bytes(g, mode == LFI_GEMM_BF16X3 ? 2 : 1);
}
}  // namespace lfi

extern "C" size_t lfi_gemm_ws_bytes(int mode, int transA, int transB, int M, int N, int K, int batch) {
  lfi::GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.transA = transA; g.transB = transB; g.M = M; g.N = N; g.K = K; g.batch = batch;
  return lfi::gemm_ws_bytes(mode, g);
}
