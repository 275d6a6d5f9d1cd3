// Persistent windowed encoder GRU (ModalityEncoder `enc: rnn`, reference models.py:21-27, 55-69): ALL window steps of a tile of
// 128 windows in ONE launch, the GRU state resident in shared memory between the steps (enc_persist.cu).
#pragma once
#include "lfi_common.cuh"

namespace lfi {
namespace encp {

struct FwdArgs {
  int E, hist, B, T, t0, M;       // hidden size, window length, batch geometry (row m = t' * B + b), rows
  int nplanes;                    // 2 = split-bf16 (three products), 1 = bf16
  const float *xp;                // [B*T][3E]  x @ W_ih^T (no bias), computed once per raw frame
  const float *b_ih, *b_hh;       // [3E]
  const float *mask;              // [M][hist] frame-dropout mask (already scaled) or nullptr
  const void *whh_hi, *whh_lo;    // bf16 planes of W_hh [3E][E]
  // outputs; the per-step arrays are indexed [s][M][..] when stash != 0 and absent otherwise (sampling)
  int stash;
  float *hs;                      // [hist][M][E] fp32 state (nullable)
  void *hp_hi, *hp_lo;            // [hist][M][E] bf16 planes of the state (nullable; operand of dW_hh)
  void *gates; int gates16;       // [hist][M][3E] r, u, n after activation: fp32 or 16-bit fixed point (nullable)
  float *ahn;                     // [hist][M][E] h-side n pre-activation (nullable)
  float *cond; int cond_ld;       // final state -> cond[m * cond_ld + e]
};

bool fwd_supported(int E, int hist, size_t M, int mode);
int launch_fwd(const FwdArgs &a, cudaStream_t st);

}  // namespace encp

namespace tc {
// gemm_tc.cu: TMA descriptor of a bf16 plane [batch][rows][ldp], box = 64 columns x box_rows rows, 128-byte swizzle
int make_plane_map(void *map, const void *plane, int rows, int cols, int ldp, long stride, int batch, int box_rows);
}  // namespace tc
}  // namespace lfi
