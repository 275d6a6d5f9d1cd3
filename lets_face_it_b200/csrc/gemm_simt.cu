// fp32 FFMA tiled GEMM: the exact-fp32 path of the time-parallel contractions and the on-device
// cross-check of the tcgen05 path (gemm_tc.cu).  C[b] = op(A[b]) op(B[b]) with a fused epilogue.
//   transA = 0: A stored [M,K] (lda)   transA = 1: A stored [K,M] (lda)   (wgrad: reduction over rows)
//   transB = 0: B stored [K,N] (ldb)   transB = 1: B stored [N,K] (ldb)   (nn.Linear weight layout)
// 128x128x16 tiles, 256 threads, 8x8 register micro-tile, float4 global/shared accesses when the
// operand pitches allow it.  Split-K (atomicAdd) for accumulate-only epilogues with few output tiles.
#include "lfi_common.cuh"

namespace lfi {

namespace {

constexpr int BM = 128, BN = 128, BK = 16, LDS_ = BM + 4;

// Loads a [rows x BK] slab whose reduction index is contiguous in memory (stored [R, K]) and writes it
// transposed into s[k][r].
template <bool VEC>
__device__ __forceinline__ void load_k_contig(float *s, const float *g, int ld, int r0, int k0, int R, int K) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int r = (t >> 2) + 64 * i, kq = (t & 3) * 4;
    const int gr = r0 + r, gk = k0 + kq;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (gr < R) {
      const float *p = g + (size_t)gr * ld + gk;
      if (VEC && gk + 3 < K) {
        float4 q = *reinterpret_cast<const float4 *>(p);
        v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e) if (gk + e < K) v[e] = p[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) s[(kq + e) * LDS_ + r] = v[e];
  }
}

// Loads a [BK x cols] slab whose non-reduction index is contiguous (stored [K, R]) into s[k][r].
template <bool VEC>
__device__ __forceinline__ void load_r_contig(float *s, const float *g, int ld, int r0, int k0, int R, int K) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 2; ++i) {
    const int k = (t >> 5) + 8 * i, r4 = (t & 31) * 4;
    const int gk = k0 + k, gr = r0 + r4;
    float4 q = make_float4(0.f, 0.f, 0.f, 0.f);
    if (gk < K) {
      const float *p = g + (size_t)gk * ld + gr;
      if (VEC && gr + 3 < R) {
        q = *reinterpret_cast<const float4 *>(p);
      } else {
        if (gr + 0 < R) q.x = p[0];
        if (gr + 1 < R) q.y = p[1];
        if (gr + 2 < R) q.z = p[2];
        if (gr + 3 < R) q.w = p[3];
      }
    }
    *reinterpret_cast<float4 *>(&s[k * LDS_ + r4]) = q;
  }
}

template <bool TA, bool TB, bool VEC>
__global__ void __launch_bounds__(256) gemm_simt_kernel(GemmArgs g, int splitk) {
  __shared__ __align__(16) float As[2][BK * LDS_];
  __shared__ __align__(16) float Bs[2][BK * LDS_];
  const int bz = blockIdx.z / splitk, sk = blockIdx.z % splitk;
  const float *A = g.A + (size_t)bz * g.sA;
  const float *B = g.B + (size_t)bz * g.sB;
  float *C = g.C + (size_t)bz * g.sC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;

  // K range of this split (multiples of BK)
  const int ktiles = (g.K + BK - 1) / BK;
  const int per = (ktiles + splitk - 1) / splitk;
  const int kt0 = sk * per, kt1 = min(ktiles, kt0 + per);

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load = [&](int buf, int kt) {
    const int k0 = kt * BK;
    if (TA) load_r_contig<VEC>(As[buf], A, g.lda, m0, k0, g.M, g.K);
    else load_k_contig<VEC>(As[buf], A, g.lda, m0, k0, g.M, g.K);
    if (TB) load_k_contig<VEC>(Bs[buf], B, g.ldb, n0, k0, g.N, g.K);
    else load_r_contig<VEC>(Bs[buf], B, g.ldb, n0, k0, g.N, g.K);
  };

  if (kt0 < kt1) {
    load(0, kt0);
    __syncthreads();
    for (int kt = kt0; kt < kt1; ++kt) {
      const int buf = (kt - kt0) & 1;
      if (kt + 1 < kt1) load(buf ^ 1, kt + 1);
      const float *as = As[buf], *bs = Bs[buf];
#pragma unroll
      for (int k = 0; k < BK; ++k) {
        const float4 a0 = *reinterpret_cast<const float4 *>(&as[k * LDS_ + ty * 4]);
        const float4 a1 = *reinterpret_cast<const float4 *>(&as[k * LDS_ + 64 + ty * 4]);
        const float4 b0 = *reinterpret_cast<const float4 *>(&bs[k * LDS_ + tx * 4]);
        const float4 b1 = *reinterpret_cast<const float4 *>(&bs[k * LDS_ + 64 + tx * 4]);
        const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
        const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
    }
  }

  const float *bias = g.bias ? g.bias + (size_t)bz * g.sBias : nullptr;
  const float *aux = g.aux ? g.aux + (size_t)bz * g.sAux : nullptr;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= g.N) continue;
      float v = acc[i][j];
      float *c = C + (size_t)m * g.ldc + n;
      if (g.epi & LFI_EPI_ACCUM_PRE) v += *c;
      if (g.epi & LFI_EPI_BIAS) v += bias[n];
      if (g.epi & LFI_EPI_LRELU) v = v > 0.f ? v : kLeaky * v;
      if (g.epi & LFI_EPI_LRELU_BWD) v *= (aux[(size_t)m * g.ldaux + n] > 0.f ? 1.f : kLeaky);
      if (splitk > 1) atomicAdd(c, v);
      else if (g.epi & LFI_EPI_ACCUM) *c += v;
      else *c = v;
    }
  }
}

template <bool TA, bool TB>
int launch(const GemmArgs &g, cudaStream_t st, bool vec, int splitk) {
  dim3 grid((g.N + BN - 1) / BN, (g.M + BM - 1) / BM, g.batch * splitk);
  if (vec) gemm_simt_kernel<TA, TB, true><<<grid, 256, 0, st>>>(g, splitk);
  else gemm_simt_kernel<TA, TB, false><<<grid, 256, 0, st>>>(g, splitk);
  LFI_LAUNCH_CHECK();
  return LFI_OK;
}

}  // namespace

int gemm_simt(const GemmArgs &g, cudaStream_t st) {
  LFI_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0 && g.batch > 0, LFI_ERR_SHAPE, "gemm: empty problem %dx%dx%d b=%d", g.M, g.N, g.K, g.batch);
  LFI_REQUIRE(g.A && g.B && g.C, LFI_ERR_ARG, "gemm: null operand");
  LFI_REQUIRE(!(g.epi & LFI_EPI_BIAS) || g.bias, LFI_ERR_ARG, "gemm: bias epilogue without bias");
  LFI_REQUIRE(!(g.epi & LFI_EPI_LRELU_BWD) || g.aux, LFI_ERR_ARG, "gemm: lrelu-bwd epilogue without aux");
  auto al = [](const void *p) { return ((uintptr_t)p & 15) == 0; };
  const bool vec = al(g.A) && al(g.B) && g.lda % 4 == 0 && g.ldb % 4 == 0 && g.sA % 4 == 0 && g.sB % 4 == 0;
  // split-K only for pure accumulation with few output tiles and a long reduction
  int splitk = 1;
  if (g.epi == LFI_EPI_ACCUM) {
    const long tiles = (long)((g.N + BN - 1) / BN) * ((g.M + BM - 1) / BM) * g.batch;
    const int ktiles = (g.K + BK - 1) / BK;
    if (tiles < 148 && ktiles >= 64) {
      splitk = (int)((2 * 148 + tiles - 1) / tiles);
      if (splitk > ktiles / 16) splitk = ktiles / 16;
      if (splitk < 1) splitk = 1;
    }
  }
  if (!g.transA && !g.transB) return launch<false, false>(g, st, vec, splitk);
  if (!g.transA && g.transB) return launch<false, true>(g, st, vec, splitk);
  if (g.transA && !g.transB) return launch<true, false>(g, st, vec, splitk);
  return launch<true, true>(g, st, vec, splitk);
}

}  // namespace lfi
