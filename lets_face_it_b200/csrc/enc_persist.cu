// Persistent windowed encoder GRU, forward (ModalityEncoder `enc: rnn`: a one-layer batch-first nn.GRU run from a zero state over
// the `history` frames of every window, reference models.py:21-27 and 55-69; gate equations of torch nn.GRU, order r, z, n).
//
// Round 1 ran one launch per window step (recurrent product + gate math), so the state h, its operand planes and the input
// projections made a round trip through HBM / L2 24 + 2 + 16 times per training step.  Here ONE launch per modality walks all
// window steps:
//   * a 2-CTA thread-block cluster owns a tile of 128 windows (rows m = t' * B + b) for the whole window; CTA c owns the hidden
//     units [c * E/2, (c+1) * E/2), i.e. the r, u, n columns of those units (3 E/2 <= 384 fp32 accumulator columns in TMEM);
//   * the state h lives in shared memory as split-bf16 (hi, lo) operand planes in the canonical K-major 128-byte-swizzled UMMA
//     layout: it is the A operand of the next step's tcgen05.mma and (hi + lo, exact to 2^-17) the h_{s-1} of the state update.
//     Each CTA writes its half of the new state into its own planes and, through distributed shared memory, into its peer's;
//   * W_hh (this CTA's 3 E/2 rows, both planes) streams from L2 through a TMA ring (cp.async.bulk.tensor, 128-byte swizzle) that
//     runs ahead across step boundaries; one elected thread issues the MMAs (a = a_hi + a_lo: a_hi b_hi + a_lo b_hi + a_hi b_lo,
//     fp32 accumulation in TMEM), eight warps do the gate math straight from TMEM (thread = window x 16-unit chunks);
//   * the input projections xp are read from L2 (63 MB for the widest modality, re-used by every window that contains the
//     frame), the stash for the backward pass (16-bit gates, h-side n pre-activation, state planes) is the only HBM traffic.
// Synchronisation per step: mbarriers only (no cluster-wide barrier): `a_ready` (16 warp arrivals: both CTAs have written
// their halves of h_s into THIS CTA's planes and this CTA's accumulator reads are done) gates the next product; `mma_done`
// (2 arrivals: this CTA's and the peer's products of the step have completed, so nobody reads h_{s-1} any more) gates the
// state writes.  All waits are bounded (trap after 4 s).
#include "enc_persist.cuh"
#include "tc_ptx.cuh"

#include <cooperative_groups.h>
#include <cuda.h>

namespace cg = cooperative_groups;

namespace lfi {
namespace encp {

using namespace tcp;

namespace {

constexpr int kRows = 128;            // windows per cluster tile (UMMA M)
constexpr int kEpiWarps = 8;
constexpr int kThreads = 32 * (kEpiWarps + 2);  // + TMA producer warp + MMA issuer warp
constexpr int kKB = 64;               // bf16 elements per k-block (one 128-byte swizzle span)
constexpr int kAKB = kRows * 128;     // bytes of one k-block of one A plane
constexpr int kTmemCols = 512;
constexpr int kMaxStages = 4;

__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// arrive (release, cluster scope) on a barrier of any CTA of the cluster, given its shared::cluster address
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar_cluster) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(bar_cluster) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_cluster(uint64_t *bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// bounded wait with cluster-scope acquire (the arrivals come from both CTAs of the cluster)
__device__ __forceinline__ void mbar_wait_cluster(uint64_t *bar, uint32_t parity) {
  if (mbar_try_wait_cluster(bar, parity)) return;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  while (!mbar_try_wait_cluster(bar, parity)) {
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 4000000000ull) {
      printf("lfi enc_persist: mbarrier wait timed out (block %d,%d thread %d)\n", blockIdx.x, blockIdx.y, threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_3d(void *smem, const CUtensorMap *map, uint64_t *bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(smem_u32(smem)), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ float bf_lo(uint32_t w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(uint32_t w) { return __uint_as_float(w & 0xffff0000u); }
__device__ __forceinline__ void unpack8(const uint4 &hi, const uint4 &lo, float *v) {  // v[i] = hi[i] + lo[i]
  v[0] = bf_lo(hi.x) + bf_lo(lo.x); v[1] = bf_hi(hi.x) + bf_hi(lo.x);
  v[2] = bf_lo(hi.y) + bf_lo(lo.y); v[3] = bf_hi(hi.y) + bf_hi(lo.y);
  v[4] = bf_lo(hi.z) + bf_lo(lo.z); v[5] = bf_hi(hi.z) + bf_hi(lo.z);
  v[6] = bf_lo(hi.w) + bf_lo(lo.w); v[7] = bf_hi(hi.w) + bf_hi(lo.w);
}
__device__ __forceinline__ void ld16(const float *p, float *v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4 *>(p + 4 * i);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void ldg16(const float *p, float *v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = __ldg(reinterpret_cast<const float4 *>(p + 4 * i));
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void st16(float *p, const float *v) {
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4 *>(p + 4 * i) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ uint32_t pack_u16(unsigned short a, unsigned short b) { return (uint32_t)a | ((uint32_t)b << 16); }

struct SmemPlan {
  int a_plane;      // bytes of one A plane (E/64 k-blocks)
  int b_off, stage_bytes, stages;
  int bar_off, total;
};
__host__ __device__ inline SmemPlan plan(int E) {
  SmemPlan p;
  const int UH = E / 2;
  p.a_plane = (E / kKB) * kAKB;
  p.b_off = 2 * p.a_plane;
  p.stage_bytes = 3 * UH * 128;
  const int budget = 227 * 1024 - 1024 - 256 - p.b_off;
  p.stages = budget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.bar_off = p.b_off + p.stages * p.stage_bytes;
  p.total = p.bar_off + 256 + 1024;  // barriers + alignment reserve
  return p;
}

}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
enc_gru_fwd_persist(const FwdArgs a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(16) uint8_t smraw[];
  uint8_t *smb = (uint8_t *)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, UH = E >> 1, nkb = E / kKB, hist = a.hist;
  const SmemPlan pl = plan(E);
  uint8_t *sA = smb, *sB = smb + pl.b_off;
  uint64_t *bars = (uint64_t *)(smb + pl.bar_off);
  uint64_t *full = bars, *empty = bars + kMaxStages, *mma_local = bars + 2 * kMaxStages, *mma_done = mma_local + 1, *a_ready = mma_local + 2;
  uint32_t *tmem_slot = (uint32_t *)(mma_local + 3);
  uint8_t *peerA = cluster.map_shared_rank(smb, c ^ 1);

  if (tid == 0) {
    for (int s = 0; s < pl.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_init(mma_local, 1);
    mbar_init(mma_done, 2);                // this CTA's relay + the peer's
    mbar_init(a_ready, 2 * kEpiWarps);     // every epilogue warp of both CTAs
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, kTmemCols);
  fence_before();
  __syncthreads();
  cluster.sync();  // both CTAs run and their barriers are initialised before any remote access
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tile = blockIdx.y;
  const uint32_t my_rank = (uint32_t)c, peer_rank = (uint32_t)(c ^ 1);

  if (warp == kEpiWarps) {
    // ============================== TMA producer: this CTA's rows of W_hh, plane by plane, k-block by k-block ==============
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int s = 1; s < hist; ++s)
        for (int kb = 0; kb < nkb; ++kb)
          for (int p = 0; p < a.nplanes; ++p) {
            mbar_wait(&empty[st], ph ^ 1);
            mbar_expect_tx(&full[st], (uint32_t)pl.stage_bytes);
            uint8_t *dst = sB + (size_t)st * pl.stage_bytes;
            const CUtensorMap *mp = p ? &map_lo : &map_hi;
#pragma unroll
            for (int g = 0; g < 3; ++g) tma_load_3d(dst + g * UH * 128, mp, &full[st], kb * kKB, g * E + c * UH, 0);
            if (++st == pl.stages) { st = 0; ph ^= 1; }
          }
    }
  } else if (warp == kEpiWarps + 1) {
    // ============================== MMA issuer ==============================================================================
    if (lane == 0) {
      const uint32_t sAu = smem_u32(sA), sBu = smem_u32(sB);
      const uint32_t done_own = mapa_u32(smem_u32(mma_done), my_rank), done_peer = mapa_u32(smem_u32(mma_done), peer_rank);
      int st = 0; uint32_t ph = 0;
      for (int s = 1; s < hist; ++s) {
        mbar_wait_cluster(a_ready, (uint32_t)((s - 1) & 1));  // h_{s-1} complete in this CTA's planes, accumulators drained
        fence_after();
        fence_async_smem();
        for (int kb = 0; kb < nkb; ++kb)
          for (int p = 0; p < a.nplanes; ++p) {
            mbar_wait(&full[st], ph);
            fence_after();
            const uint32_t bbase = sBu + (uint32_t)st * pl.stage_bytes;
            // plane 0 (b_hi): a_hi b_hi (+ a_lo b_hi in the split mode); plane 1 (b_lo): a_hi b_lo
            const int na = (p == 0 && a.nplanes == 2) ? 2 : 1;
            for (int pa = 0; pa < na; ++pa) {
              const uint32_t abase = sAu + (uint32_t)pa * pl.a_plane + (uint32_t)kb * kAKB;
#pragma unroll
              for (int ks = 0; ks < kKB / 16; ++ks) {
                const uint32_t acc = (kb > 0 || p > 0 || pa > 0 || ks > 0) ? 1u : 0u;
                for (int n0 = 0; n0 < 3 * UH; n0 += 256) {
                  const int n = min(256, 3 * UH - n0);
                  umma_bf16(tmem + n0, make_sdesc(abase + ks * 32, 1024, kSw128), make_sdesc(bbase + n0 * 128 + ks * 32, 1024, kSw128),
                            idesc_bf16_m128(n), acc);
                }
              }
            }
            umma_commit(&empty[st]);  // the stage is free once the products above have read it
            if (++st == pl.stages) { st = 0; ph ^= 1; }
          }
        umma_commit(mma_local);
        mbar_wait(mma_local, (uint32_t)((s - 1) & 1));  // this CTA's products of step s are complete ...
        mbar_arrive_cluster(done_own);                  // ... tell this CTA's and the peer's gate warps
        mbar_arrive_cluster(done_peer);
      }
    }
  } else {
    // ============================== gate math: thread = window (TMEM lane) x a quarter of the cluster's hidden units ==========
    const int q = warp & 3, sub = warp >> 2;
    const int L = 32 * q + lane;                // row of the tile = TMEM lane
    const size_t m_raw = (size_t)tile * kRows + L;
    const bool row_ok = m_raw < (size_t)a.M;
    const size_t m = row_ok ? m_raw : (size_t)a.M - 1;  // clamped: rows beyond M compute on valid addresses and store nothing
    const int b = (int)(m % a.B), tp = (int)(m / a.B);
    const int UQ = UH >> 1, nch = UQ / 16;
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    const uint32_t ready_own = mapa_u32(smem_u32(a_ready), my_rank), ready_peer = mapa_u32(smem_u32(a_ready), peer_rank);
    const bool cond_vec = a.cond && (((uintptr_t)a.cond & 15) == 0) && (a.cond_ld % 4 == 0);
    const size_t ME = (size_t)a.M * E;
    for (int s = 0; s < hist; ++s) {
      const float mk = a.mask ? __ldg(a.mask + m * hist + s) : 1.0f;
      const int tau = a.t0 + tp - hist + 1 + s;
      const float *xrow = a.xp + ((size_t)b * a.T + tau) * 3 * E;
      const bool lastStep = (s == hist - 1);
      float xr[16], xu[16], xn[16];
      {  // operands of the first chunk are requested before the wait for the products
        const int ug0 = c * UH + sub * UQ;
        ld16(xrow + ug0, xr); ld16(xrow + E + ug0, xu); ld16(xrow + 2 * E + ug0, xn);
      }
      if (s > 0) {
        mbar_wait_cluster(mma_done, (uint32_t)((s - 1) & 1));
        fence_after();
      }
      for (int j = 0; j < nch; ++j) {
        const int ul = sub * UQ + 16 * j, ug = c * UH + ul;
        if (j > 0) { ld16(xrow + ug, xr); ld16(xrow + E + ug, xu); ld16(xrow + 2 * E + ug, xn); }
        float ar[16], au[16], an[16], hp[16];
        const uint32_t aoff = (uint32_t)(ug >> 6) * kAKB;       // k-block of these units inside a plane
        const uint32_t o0 = aoff + sw128_off(L, (ug & 63) >> 3), o1 = aoff + sw128_off(L, ((ug & 63) >> 3) + 1);
        if (s > 0) {
          tmem_ld16(tlane + ul, ar);
          tmem_ld16(tlane + UH + ul, au);
          tmem_ld16(tlane + 2 * UH + ul, an);
          const uint4 h0 = *reinterpret_cast<const uint4 *>(sA + o0), l0 = *reinterpret_cast<const uint4 *>(sA + pl.a_plane + o0);
          const uint4 h1 = *reinterpret_cast<const uint4 *>(sA + o1), l1 = *reinterpret_cast<const uint4 *>(sA + pl.a_plane + o1);
          unpack8(h0, l0, hp); unpack8(h1, l1, hp + 8);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 16; ++i) { ar[i] = 0.f; au[i] = 0.f; an[i] = 0.f; hp[i] = 0.f; }
        }
        // gate by gate, so that only one pair of bias vectors is live at a time (register budget: 204 per thread)
        float hn[16];
        {
          float bi[16], bh[16];
          ldg16(a.b_ih + ug, bi); ldg16(a.b_hh + ug, bh);
#pragma unroll
          for (int i = 0; i < 16; ++i) ar[i] = fast_sigmoid(mk * xr[i] + bi[i] + (ar[i] + bh[i]));            // r
          ldg16(a.b_ih + E + ug, bi); ldg16(a.b_hh + E + ug, bh);
#pragma unroll
          for (int i = 0; i < 16; ++i) au[i] = fast_sigmoid(mk * xu[i] + bi[i] + (au[i] + bh[i]));            // u (z in torch's naming)
          ldg16(a.b_ih + 2 * E + ug, bi); ldg16(a.b_hh + 2 * E + ug, bh);
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            an[i] += bh[i];                                                                                  // h-side n pre-activation
            xn[i] = fast_tanh(mk * xn[i] + bi[i] + ar[i] * an[i]);                                           // n
            hn[i] = xn[i] + au[i] * (hp[i] - xn[i]);                                                         // h' = (1 - u) n + u h
          }
        }
        uint4 hi0, lo0, hi1, lo1;
        {
          const float v0[8] = {hn[0], hn[1], hn[2], hn[3], hn[4], hn[5], hn[6], hn[7]};
          const float v1[8] = {hn[8], hn[9], hn[10], hn[11], hn[12], hn[13], hn[14], hn[15]};
          split8(v0, hi0, lo0);
          split8(v1, hi1, lo1);
        }
        if (!lastStep) {  // new state into the operand planes of both CTAs (nobody reads h_{s-1} any more: mma_done)
          *reinterpret_cast<uint4 *>(sA + o0) = hi0; *reinterpret_cast<uint4 *>(sA + pl.a_plane + o0) = lo0;
          *reinterpret_cast<uint4 *>(sA + o1) = hi1; *reinterpret_cast<uint4 *>(sA + pl.a_plane + o1) = lo1;
          *reinterpret_cast<uint4 *>(peerA + o0) = hi0; *reinterpret_cast<uint4 *>(peerA + pl.a_plane + o0) = lo0;
          *reinterpret_cast<uint4 *>(peerA + o1) = hi1; *reinterpret_cast<uint4 *>(peerA + pl.a_plane + o1) = lo1;
        }
        if (row_ok) {
          const size_t o1e = (a.stash ? (size_t)s * ME : 0) + m * E + ug;
          if (a.stash) {
            if (a.hs) st16(a.hs + o1e, hn);
            if (a.hp_hi) {
              __nv_bfloat16 *ph_ = (__nv_bfloat16 *)a.hp_hi + o1e;
              *reinterpret_cast<uint4 *>(ph_) = hi0; *reinterpret_cast<uint4 *>(ph_ + 8) = hi1;
              if (a.hp_lo) {
                __nv_bfloat16 *pl_ = (__nv_bfloat16 *)a.hp_lo + o1e;
                *reinterpret_cast<uint4 *>(pl_) = lo0; *reinterpret_cast<uint4 *>(pl_ + 8) = lo1;
              }
            }
            if (a.gates) {
              const size_t o3 = ((size_t)s * a.M + m) * 3 * E + ug;
              if (a.gates16) {
                unsigned short *gq = reinterpret_cast<unsigned short *>(a.gates) + o3;
                uint32_t w[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_u16(q_unorm16(ar[2 * i]), q_unorm16(ar[2 * i + 1]));
                *reinterpret_cast<uint4 *>(gq) = make_uint4(w[0], w[1], w[2], w[3]); *reinterpret_cast<uint4 *>(gq + 8) = make_uint4(w[4], w[5], w[6], w[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_u16(q_unorm16(au[2 * i]), q_unorm16(au[2 * i + 1]));
                *reinterpret_cast<uint4 *>(gq + E) = make_uint4(w[0], w[1], w[2], w[3]); *reinterpret_cast<uint4 *>(gq + E + 8) = make_uint4(w[4], w[5], w[6], w[7]);
#pragma unroll
                for (int i = 0; i < 8; ++i) w[i] = pack_u16(q_snorm16(xn[2 * i]), q_snorm16(xn[2 * i + 1]));
                *reinterpret_cast<uint4 *>(gq + 2 * E) = make_uint4(w[0], w[1], w[2], w[3]); *reinterpret_cast<uint4 *>(gq + 2 * E + 8) = make_uint4(w[4], w[5], w[6], w[7]);
              } else {
                float *gf = reinterpret_cast<float *>(a.gates) + o3;
                st16(gf, ar); st16(gf + E, au); st16(gf + 2 * E, xn);
              }
            }
            if (a.ahn) st16(a.ahn + o1e, an);
          }
          if (lastStep && a.cond) {
            float *cd = a.cond + m * a.cond_ld + ug;
            if (cond_vec) st16(cd, hn);
            else {
#pragma unroll
              for (int i = 0; i < 16; ++i) cd[i] = hn[i];
            }
          }
        }
      }
      if (!lastStep) {
        fence_before();      // this warp's accumulator reads are complete before the next products overwrite them
        fence_async_smem();  // the new state is visible to the tensor cores of both CTAs
        __syncwarp();
        if (lane == 0) { mbar_arrive_cluster(ready_own); mbar_arrive_cluster(ready_peer); }
      }
    }
  }

  fence_before();
  __syncthreads();
  cluster.sync();  // no CTA leaves while its peer may still write into its planes or signal its barriers
  if (warp == kEpiWarps + 1) {
    fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

// ------------------------------------------------------------------------------------------------
bool fwd_supported(int E, int hist, size_t M, int mode) {
  if (mode == LFI_GEMM_FP32 || !env_flag("LFI_ENC_PERSIST", true)) return false;
  if (!(E == 128 || E == 256) || hist < 1 || M < 128) return false;
  const SmemPlan p = plan(E);
  return p.stages >= 2 && p.total <= 227 * 1024;
}

int launch_fwd(const FwdArgs &a, cudaStream_t st) {
  LFI_REQUIRE(a.E == 128 || a.E == 256, LFI_ERR_SHAPE, "enc_persist: E=%d unsupported", a.E);
  LFI_REQUIRE(a.xp && a.b_ih && a.b_hh && a.whh_hi && (a.nplanes == 1 || a.whh_lo), LFI_ERR_ARG, "enc_persist: null argument");
  auto al16 = [](const void *p) { return p == nullptr || ((uintptr_t)p & 15) == 0; };
  LFI_REQUIRE(al16(a.xp) && al16(a.b_ih) && al16(a.b_hh) && al16(a.hs) && al16(a.hp_hi) && al16(a.hp_lo) && al16(a.gates) && al16(a.ahn),
              LFI_ERR_ARG, "enc_persist: misaligned pointer");
  const SmemPlan p = plan(a.E);
  LFI_REQUIRE(p.stages >= 2 && p.total <= 227 * 1024, LFI_ERR_SHAPE, "enc_persist: shared-memory plan does not fit");
  CUtensorMap mhi, mlo;
  LFI_TRY(tc::make_plane_map(&mhi, a.whh_hi, 3 * a.E, a.E, a.E, 0, 1, a.E / 2));
  if (a.nplanes == 2) LFI_TRY(tc::make_plane_map(&mlo, a.whh_lo, 3 * a.E, a.E, a.E, 0, 1, a.E / 2));
  else mlo = mhi;
  static bool attr_set = false;
  if (!attr_set) {
    LFI_CUDA(cudaFuncSetAttribute(enc_gru_fwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int ntiles = (a.M + kRows - 1) / kRows;
  // > half of the SM's shared memory: never two CTAs (each wanting all 512 TMEM columns) on one SM
  const int smem = p.total < 120 * 1024 ? 120 * 1024 : p.total;
  dim3 grid(2, ntiles, 1);
  enc_gru_fwd_persist<<<grid, kThreads, smem, st>>>(a, mhi, mlo);
  LFI_LAUNCH_CHECK();
  return LFI_OK;
}

}  // namespace encp
}  // namespace lfi
