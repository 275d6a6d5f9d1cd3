// Persistent windowed encoder GRU, forward and backward (ModalityEncoder `enc: rnn`: a one-layer batch-first nn.GRU run from a zero state over
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
//     frame), the stash for the backward pass (16-bit gates, h-side n pre-activation, state) is the only HBM traffic.  Both use
//     the row-interleaved layouts of enc_persist.cuh, so that the thread-per-window accesses of a warp are contiguous (the first
//     version used row-major arrays: 32 lines per warp access, 62 k cycles of gate phase per step against 18 k of products).
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

constexpr int kStageBytes = 192 * 128;  // forward ring stage: one 64-unit half (r, u, n rows) of one k-block of one W_hh plane
constexpr int kMaxStages = 8;

struct SmemPlan {
  int a_plane;      // bytes of one A plane (E/64 k-blocks)
  int b_off, stages;
  int bar_off, total;
};
__host__ __device__ inline SmemPlan plan(int E) {
  SmemPlan p;
  p.a_plane = (E / kKB) * kAKB;
  p.b_off = 2 * p.a_plane;
  const int budget = 227 * 1024 - 1024 - 256 - p.b_off;
  p.stages = budget / kStageBytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.bar_off = p.b_off + p.stages * kStageBytes;
  p.total = p.bar_off + 256 + 1024;  // barriers + alignment reserve
  return p;
}

// 16 consecutive columns n0 .. n0+15 of row `row` of a tiled fp32 array of width W (enc_persist.cuh)
__device__ __forceinline__ void ldt16(const float *base, size_t row, int n0, int W, float *v) {
  const float *p = base + ((row >> 5) * (size_t)(W >> 2) + (size_t)(n0 >> 2)) * 128 + (row & 31) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4 t = *reinterpret_cast<const float4 *>(p + i * 128);
    v[4 * i] = t.x; v[4 * i + 1] = t.y; v[4 * i + 2] = t.z; v[4 * i + 3] = t.w;
  }
}
__device__ __forceinline__ void stt16(float *base, size_t row, int n0, int W, const float *v) {
  float *p = base + ((row >> 5) * (size_t)(W >> 2) + (size_t)(n0 >> 2)) * 128 + (row & 31) * 4;
#pragma unroll
  for (int i = 0; i < 4; ++i) *reinterpret_cast<float4 *>(p + i * 128) = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
}
__device__ __forceinline__ unsigned short *u16_ptr(void *base, size_t row, int n0, int W) {
  return reinterpret_cast<unsigned short *>(base) + ((row >> 5) * (size_t)(W >> 3) + (size_t)(n0 >> 3)) * 256 + (row & 31) * 8;
}

}  // namespace

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
enc_gru_fwd_persist(const FwdArgs a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(16) uint8_t smraw[];
  uint8_t *smb = (uint8_t *)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  cg::cluster_group cluster = cg::this_cluster();
  const int c = (int)cluster.block_rank();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, UH = E >> 1, NH = UH >> 6, nkb = E / kKB, hist = a.hist;
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
    mbar_init(a_ready, 2 * kEpiWarps);     // every gate warp of both CTAs
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
    // ============================== TMA producer: this CTA's rows of W_hh: k-block, plane, 64-unit half ====================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int s = 1; s < hist; ++s)
        for (int kb = 0; kb < nkb; ++kb)
          for (int p = 0; p < a.nplanes; ++p)
            for (int hf = 0; hf < NH; ++hf) {
              mbar_wait(&empty[st], ph ^ 1);
              mbar_expect_tx(&full[st], (uint32_t)kStageBytes);
              uint8_t *dst = sB + (size_t)st * kStageBytes;
              const CUtensorMap *mp = p ? &map_lo : &map_hi;
#pragma unroll
              for (int g = 0; g < 3; ++g) tma_load_3d(dst + g * 64 * 128, mp, &full[st], kb * kKB, g * E + c * UH + hf * 64, 0);
              if (++st == pl.stages) { st = 0; ph ^= 1; }
            }
    }
  } else if (warp == kEpiWarps + 1) {
    // ============================== MMA issuer ==============================================================================
    if (lane == 0) {
      const uint32_t sAu = smem_u32(sA), sBu = smem_u32(sB);
      const uint32_t done_own = mapa_u32(smem_u32(mma_done), my_rank), done_peer = mapa_u32(smem_u32(mma_done), peer_rank);
      const uint32_t idesc = idesc_bf16_m128(192);
      int st = 0; uint32_t ph = 0;
      const bool tm = a.timing && blockIdx.x == 0 && blockIdx.y == 0;
      long long t_wait = 0, t_mma = 0;
      for (int s = 1; s < hist; ++s) {
        const long long c0 = tm ? clock64() : 0;
        mbar_wait_cluster(a_ready, (uint32_t)((s - 1) & 1));  // h_{s-1} complete in this CTA's planes, accumulators drained
        const long long c1 = tm ? clock64() : 0;
        t_wait += c1 - c0;
        fence_after();
        fence_async_smem();
        for (int kb = 0; kb < nkb; ++kb)
          for (int p = 0; p < a.nplanes; ++p)
            for (int hf = 0; hf < NH; ++hf) {
              mbar_wait(&full[st], ph);
              fence_after();
              const uint32_t bbase = sBu + (uint32_t)st * kStageBytes;
              // plane 0 (b_hi): a_hi b_hi (+ a_lo b_hi in the split mode); plane 1 (b_lo): a_hi b_lo
              const int na = (p == 0 && a.nplanes == 2) ? 2 : 1;
              for (int pa = 0; pa < na; ++pa) {
                const uint32_t abase = sAu + (uint32_t)pa * pl.a_plane + (uint32_t)kb * kAKB;
#pragma unroll
                for (int ks = 0; ks < kKB / 16; ++ks) {
                  const uint32_t acc = (kb > 0 || p > 0 || pa > 0 || ks > 0) ? 1u : 0u;
                  umma_bf16(tmem + hf * 192, make_sdesc(abase + ks * 32, 1024, kSw128), make_sdesc(bbase + ks * 32, 1024, kSw128), idesc, acc);
                }
              }
              umma_commit(&empty[st]);  // the stage is free once the products above have read it
              if (++st == pl.stages) { st = 0; ph ^= 1; }
            }
        umma_commit(mma_local);
        mbar_wait(mma_local, (uint32_t)((s - 1) & 1));  // this CTA's products of step s are complete ...
        mbar_arrive_cluster(done_own);                  // ... tell this CTA's and the peer's gate warps
        mbar_arrive_cluster(done_peer);
        if (tm) t_mma += clock64() - c1;
      }
      if (tm && hist > 1) printf("enc_persist fwd E=%d hist=%d: MMA warp per step: wait a_ready %lld cycles, products %lld cycles\n", E, hist, t_wait / (hist - 1), t_mma / (hist - 1));
    }
  } else {
    // ============================== gate math: thread = window (TMEM lane) x a quarter of the cluster's hidden units ==========
    const int q = warp & 3, sub = warp >> 2;
    const int L = 32 * q + lane;                // row of the tile = TMEM lane
    const size_t m_raw = (size_t)tile * kRows + L;
    const bool row_ok = m_raw < (size_t)a.M;
    const size_t m = row_ok ? m_raw : (size_t)a.M - 1;  // clamped: rows beyond M compute on valid addresses and store nothing
    const int UQ = UH >> 1, nch = UQ / 16;
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    const uint32_t ready_own = mapa_u32(smem_u32(a_ready), my_rank), ready_peer = mapa_u32(smem_u32(a_ready), peer_rank);
    const bool cond_vec = a.cond && (((uintptr_t)a.cond & 15) == 0) && (a.cond_ld % 4 == 0);
    const size_t Mp = tiled_rows((size_t)a.M);
    const bool tm = a.timing && blockIdx.x == 0 && blockIdx.y == 0 && tid == 0;
    long long t_wait = 0, t_epi = 0;
    for (int s = 0; s < hist; ++s) {
      const long long c0 = tm ? clock64() : 0;
      const float mk = a.mask ? __ldg(a.mask + m * hist + s) : 1.0f;
      const size_t xr_row = m + (size_t)(a.t0 - hist + 1 + s) * a.B;  // raw frame of this window at this step (time-major rows)
      const bool lastStep = (s == hist - 1);
      float xr[16], xu[16], xn[16];
      {  // operands of the first chunk are requested before the wait for the products
        const int ug0 = c * UH + sub * UQ;
        ldt16(a.xp, xr_row, ug0, 3 * E, xr); ldt16(a.xp, xr_row, E + ug0, 3 * E, xu); ldt16(a.xp, xr_row, 2 * E + ug0, 3 * E, xn);
      }
      if (s > 0) {
        mbar_wait_cluster(mma_done, (uint32_t)((s - 1) & 1));
        fence_after();
      }
      const long long c1 = tm ? clock64() : 0;
      t_wait += c1 - c0;
      for (int j = 0; j < nch; ++j) {
        const int ul = sub * UQ + 16 * j, ug = c * UH + ul;
        if (j > 0) { ldt16(a.xp, xr_row, ug, 3 * E, xr); ldt16(a.xp, xr_row, E + ug, 3 * E, xu); ldt16(a.xp, xr_row, 2 * E + ug, 3 * E, xn); }
        float ar[16], au[16], an[16], hp[16];
        const uint32_t aoff = (uint32_t)(ug >> 6) * kAKB;       // k-block of these units inside a plane
        const uint32_t o0 = aoff + sw128_off(L, (ug & 63) >> 3), o1 = aoff + sw128_off(L, ((ug & 63) >> 3) + 1);
        if (s > 0) {
          const uint32_t tcol = (uint32_t)((ul >> 6) * 192 + (ul & 63));  // accumulator columns: per 64-unit half [r | u | n]
          tmem_ld16(tlane + tcol, ar);
          tmem_ld16(tlane + tcol + 64, au);
          tmem_ld16(tlane + tcol + 128, an);
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
          if (a.stash) {
            if (a.hs) stt16(a.hs + (size_t)s * Mp * E, m, ug, E, hn);
            if (a.hp_hi) {
              const size_t o1e = ((size_t)s * a.M + m) * E + ug;
              __nv_bfloat16 *ph_ = (__nv_bfloat16 *)a.hp_hi + o1e;
              *reinterpret_cast<uint4 *>(ph_) = hi0; *reinterpret_cast<uint4 *>(ph_ + 8) = hi1;
              if (a.hp_lo) {
                __nv_bfloat16 *pl_ = (__nv_bfloat16 *)a.hp_lo + o1e;
                *reinterpret_cast<uint4 *>(pl_) = lo0; *reinterpret_cast<uint4 *>(pl_ + 8) = lo1;
              }
            }
            if (a.gates) {
              float *gstep = reinterpret_cast<float *>(a.gates) + (size_t)s * Mp * 3 * E;  // per-step blocks of Mp * 3E floats in both formats
              if (a.gates16) {
                uint32_t w[8];
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                  const float *src = g == 0 ? ar : (g == 1 ? au : xn);
#pragma unroll
                  for (int i = 0; i < 8; ++i)
                    w[i] = g < 2 ? pack_u16(q_unorm16(src[2 * i]), q_unorm16(src[2 * i + 1])) : pack_u16(q_snorm16(src[2 * i]), q_snorm16(src[2 * i + 1]));
                  unsigned short *gq = u16_ptr(gstep, m, g * E + ug, 3 * E);
                  *reinterpret_cast<uint4 *>(gq) = make_uint4(w[0], w[1], w[2], w[3]);
                  *reinterpret_cast<uint4 *>(gq + 256) = make_uint4(w[4], w[5], w[6], w[7]);
                }
              } else {
                stt16(gstep, m, ug, 3 * E, ar); stt16(gstep, m, E + ug, 3 * E, au); stt16(gstep, m, 2 * E + ug, 3 * E, xn);
              }
            }
            if (a.ahn) stt16(a.ahn + (size_t)s * Mp * E, m, ug, E, an);
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
      if (tm) t_epi += clock64() - c1;
    }
    if (tm) printf("enc_persist fwd E=%d hist=%d: gate warp 0 per step: wait products %lld cycles, gate math + stores %lld cycles\n", E, hist, t_wait / hist, t_epi / hist);
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
  LFI_TRY(tc::make_plane_map(&mhi, a.whh_hi, 3 * a.E, a.E, a.E, 0, 1, 64));
  if (a.nplanes == 2) LFI_TRY(tc::make_plane_map(&mlo, a.whh_lo, 3 * a.E, a.E, a.E, 0, 1, 64));
  else mlo = mhi;
  static const bool timing = env_flag("LFI_ENC_TIMING", false);
  FwdArgs at = a;
  at.timing = timing ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    LFI_CUDA(cudaFuncSetAttribute(enc_gru_fwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int ntiles = (a.M + kRows - 1) / kRows;
  // > half of the SM's shared memory: never two CTAs (each wanting all 512 TMEM columns) on one SM
  const int smem = p.total < 120 * 1024 ? 120 * 1024 : p.total;
  dim3 grid(2, ntiles, 1);
  enc_gru_fwd_persist<<<grid, kThreads, smem, st>>>(at, mhi, mlo);
  LFI_LAUNCH_CHECK();
  return LFI_OK;
}

// ================================================================================================
// Backward: one CTA per tile of 128 windows walks the window steps in reverse.  The hidden units are processed in chunks of 32:
// the gate warps (thread = window x 16 units) turn dh_s into the gate gradients of the chunk and write them as split-bf16
// operand planes [128 windows x (da_r | da_u | da_n r) of 32 units] into one of two shared-memory buffers; the MMA warp
// accumulates dh_{s-1} += dA_h(chunk) W_hh[chunk rows, :] into the other of two TMEM accumulators ([128 x E] fp32 each) while
// the gate warps work on the next chunk.  W_hh streams from L2 as MN-major operand blocks (TMA boxes of 32 rows x 64 columns).
// The direct part dh_s * u_s travels through an L2-resident tiled scratch (same thread writes and re-reads it); the gate
// gradients leave for the weight-gradient GEMMs through a per-warp shared-memory transpose, so that the global stores cover
// whole 96-byte row segments instead of one 16-byte piece per line.
namespace {

constexpr int kCU = 32;                          // hidden units per chunk
constexpr int kAGate = kRows * 64;               // bytes of one gate block [128 windows x 32 units] of one plane (64-byte swizzle)
constexpr int kABuf = 2 * 3 * kAGate;            // one chunk buffer: 2 planes x 3 gates
constexpr int kStgRow = 112;                     // staging row pitch in bytes (96 + 16: conflict-free 128-bit accesses)
constexpr int kStgWarp = 32 * kStgRow;

struct BwdPlan {
  int acc_off, stg_off, b_off, stage_bytes, stages, bar_off, total;
};
__host__ __device__ inline BwdPlan bwd_plan(int E) {
  BwdPlan p;
  p.acc_off = 2 * kABuf;                          // bias-gradient accumulators [4][E] fp32
  p.stg_off = p.acc_off + 4 * E * 4;
  p.b_off = (p.stg_off + kEpiWarps * kStgWarp + 1023) & ~1023;
  p.stage_bytes = (E / 64) * 32 * 128;            // W_hh rows [g*E + 32j, +32) x all E columns of one plane: E/64 boxes of [32][64]
  const int budget = 227 * 1024 - 1024 - 256 - p.b_off;
  p.stages = budget / p.stage_bytes;
  if (p.stages > kMaxStages) p.stages = kMaxStages;
  p.bar_off = p.b_off + p.stages * p.stage_bytes;
  p.total = p.bar_off + 256 + 1024;
  return p;
}

// MN-major operand block (rows = K index, 64 MN elements = 128 bytes per row, 128-byte swizzle): LBO = stride between
// 64-element MN chunks, SBO = 8 K-rows (gemm_tc.cu: make_sdesc with b_mn)
__device__ __forceinline__ uint64_t make_sdesc_mn(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= 1ull << 46;
  d |= 2ull << 61;  // SWIZZLE_128B
  return d;
}

// Sum over the 32 lanes of 64 values per lane; lane l ends up with the totals of indices 2l and 2l+1 (in v[0], v[1]).
__device__ __forceinline__ void warp_reduce64(float (&v)[64], int lane) {
#pragma unroll
  for (int half = 32, bit = 16; half >= 2; half >>= 1, bit >>= 1) {
    const bool upper = (lane & bit) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = upper ? v[i] : v[i + half];
      const float keep = upper ? v[i + half] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
    }
  }
}

}  // namespace

__global__ void __launch_bounds__(kThreads, 1)
enc_gru_bwd_persist(const BwdArgs a, const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo) {
  extern __shared__ __align__(16) uint8_t smraw[];
  uint8_t *smb = (uint8_t *)(((uintptr_t)smraw + 1023) & ~(uintptr_t)1023);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int E = a.E, hist = a.hist, nch = E / kCU;
  const BwdPlan pl = bwd_plan(E);
  uint8_t *sA = smb, *sB = smb + pl.b_off;
  float *bacc = (float *)(smb + pl.acc_off);
  uint64_t *bars = (uint64_t *)(smb + pl.bar_off);
  uint64_t *full = bars, *empty = bars + kMaxStages, *a_full = bars + 2 * kMaxStages, *a_empty = a_full + 2, *acc_full = a_full + 4;
  uint32_t *tmem_slot = (uint32_t *)(a_full + 5);

  if (tid == 0) {
    for (int s = 0; s < pl.stages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int b = 0; b < 2; ++b) { mbar_init(&a_full[b], kEpiWarps); mbar_init(&a_empty[b], 1); }
    mbar_init(acc_full, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int i = tid; i < 4 * E; i += kThreads) bacc[i] = 0.f;
  if (warp == kEpiWarps + 1) tmem_alloc(tmem_slot, kTmemCols);
  fence_before();
  __syncthreads();
  fence_after();
  const uint32_t tmem = *tmem_slot;
  const int tile = blockIdx.x;
  const int nit = hist - 1;  // iterations with a product (it = hist-1-s for s >= 1)

  if (warp == kEpiWarps) {
    // ============================== TMA producer: W_hh rows of (chunk, plane, gate) as MN-major blocks ===================
    if (lane == 0) {
      int st = 0; uint32_t ph = 0;
      for (int it = 0; it < nit; ++it)
        for (int j = 0; j < nch; ++j)
          for (int p = 0; p < a.nplanes; ++p)
            for (int g = 0; g < 3; ++g) {
              mbar_wait(&empty[st], ph ^ 1);
              mbar_expect_tx(&full[st], (uint32_t)pl.stage_bytes);
              uint8_t *dst = sB + (size_t)st * pl.stage_bytes;
              const CUtensorMap *mp = p ? &map_lo : &map_hi;
              for (int i = 0; i < E / 64; ++i) tma_load_3d(dst + i * 4096, mp, &full[st], 64 * i, g * E + kCU * j, 0);
              if (++st == pl.stages) { st = 0; ph ^= 1; }
            }
    }
  } else if (warp == kEpiWarps + 1) {
    // ============================== MMA issuer ==============================================================================
    if (lane == 0) {
      const uint32_t sAu = smem_u32(sA), sBu = smem_u32(sB);
      // D fp32, A / B bf16, A K-major, B MN-major, N = E, M = 128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 16) | ((uint32_t)(E >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      int st = 0; uint32_t ph = 0;
      int cc = 0;
      for (int it = 0; it < nit; ++it) {
        const uint32_t d_tmem = tmem + (uint32_t)(((it + 1) & 1) * 256);
        uint32_t acc = 0;
        for (int j = 0; j < nch; ++j, ++cc) {
          const int buf = cc & 1;
          mbar_wait(&a_full[buf], (uint32_t)((cc >> 1) & 1));  // the chunk's gate gradients are in the operand planes
          fence_after();
          fence_async_smem();
          for (int p = 0; p < a.nplanes; ++p)
            for (int g = 0; g < 3; ++g) {
              mbar_wait(&full[st], ph);
              fence_after();
              const uint32_t bbase = sBu + (uint32_t)st * pl.stage_bytes;
              const int na = (p == 0 && a.nplanes == 2) ? 2 : 1;
              for (int pa = 0; pa < na; ++pa) {
                const uint32_t abase = sAu + (uint32_t)buf * kABuf + (uint32_t)pa * (3 * kAGate) + (uint32_t)g * kAGate;
#pragma unroll
                for (int ks = 0; ks < kCU / 16; ++ks) {
                  umma_bf16(d_tmem, make_sdesc(abase + ks * 32, 512, kSw64), make_sdesc_mn(bbase + ks * 2048, 4096, 1024), idesc, acc);
                  acc = 1;
                }
              }
              umma_commit(&empty[st]);
              if (++st == pl.stages) { st = 0; ph ^= 1; }
            }
          umma_commit(&a_empty[buf]);  // the chunk buffer may be overwritten once the products above have read it
        }
        umma_commit(acc_full);         // dh_{s-1} (product part) is complete
      }
    }
  } else {
    // ============================== gate backward: thread = window x 16 units of the chunk ===================================
    const int q = warp & 3, sub = warp >> 2;
    const int L = 32 * q + lane;
    const size_t m_raw = (size_t)tile * kRows + L;
    const bool row_ok = m_raw < (size_t)a.M;
    const size_t m = row_ok ? m_raw : (size_t)a.M - 1;
    const size_t Mp = tiled_rows((size_t)a.M);
    const uint32_t tlane = tmem + ((uint32_t)(32 * q) << 16);
    uint8_t *stg = smb + pl.stg_off + warp * kStgWarp;
    const size_t row0 = (size_t)tile * kRows + 32 * q;  // first window of this warp
    const bool extra_vec = a.dh_extra && (((uintptr_t)a.dh_extra & 15) == 0) && (a.dh_extra_ld % 4 == 0);
    const bool tm = a.timing && blockIdx.x == 0 && tid == 0;
    long long t_wait = 0, t_gate = 0;
    int cc = 0;
    for (int it = 0; it < hist; ++it) {
      const int s = hist - 1 - it;
      const long long c0 = tm ? clock64() : 0;
      if (it > 0) {
        mbar_wait(acc_full, (uint32_t)((it - 1) & 1));
        fence_after();
      }
      const long long c1 = tm ? clock64() : 0;
      t_wait += c1 - c0;
      const uint32_t tcur = tlane + (uint32_t)((it & 1) * 256);
      const float *gstep = reinterpret_cast<const float *>(a.gates) + (size_t)s * Mp * 3 * E;
      for (int j = 0; j < nch; ++j) {
        const int ug = kCU * j + 16 * sub;
        float rg[16], ugt[16], ng[16], an[16], hp[16], dh[16];
        // ---- loads (tiled: coalesced) ----
        if (a.gates16) {
          const unsigned short *base = reinterpret_cast<const unsigned short *>(gstep);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const unsigned short *gq = base + ((m >> 5) * (size_t)((3 * E) >> 3) + (size_t)((g * E + ug) >> 3)) * 256 + (m & 31) * 8;
            const uint4 w0 = __ldg(reinterpret_cast<const uint4 *>(gq)), w1 = __ldg(reinterpret_cast<const uint4 *>(gq + 256));
            const uint32_t w[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
            float *dst = g == 0 ? rg : (g == 1 ? ugt : ng);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              if (g < 2) { dst[2 * i] = dq_unorm16((unsigned short)(w[i] & 0xffff)); dst[2 * i + 1] = dq_unorm16((unsigned short)(w[i] >> 16)); }
              else { dst[2 * i] = dq_snorm16((unsigned short)(w[i] & 0xffff)); dst[2 * i + 1] = dq_snorm16((unsigned short)(w[i] >> 16)); }
            }
          }
        } else {
          ldt16(gstep, m, ug, 3 * E, rg); ldt16(gstep, m, E + ug, 3 * E, ugt); ldt16(gstep, m, 2 * E + ug, 3 * E, ng);
        }
        ldt16(a.ahn + (size_t)s * Mp * E, m, ug, E, an);
        if (s > 0) ldt16(a.hs + (size_t)(s - 1) * Mp * E, m, ug, E, hp);
        else {
#pragma unroll
          for (int i = 0; i < 16; ++i) hp[i] = 0.f;
        }
        if (it > 0) {
          tmem_ld16(tcur + ug, dh);                 // (dA_h(s+1) W_hh)[:, ug..]
          float dd[16];
          ldt16(a.dhd, m, ug, E, dd);              // + dh_{s+1} * u_{s+1}
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 16; ++i) dh[i] += dd[i];
        } else {
          if (a.dh_extra) {
            const float *ex = a.dh_extra + m * a.dh_extra_ld + ug;
            if (extra_vec) ld16(ex, dh);
            else {
#pragma unroll
              for (int i = 0; i < 16; ++i) dh[i] = ex[i];
            }
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) dh[i] = 0.f;
          }
        }
        // ---- gate backward (aux::enc_gate_bwd2 semantics) ----
        float dar[16], dau[16], dan[16], danr[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float dn = dh[i] * (1.0f - ugt[i]), du = dh[i] * (hp[i] - ng[i]);
          dan[i] = dn * (1.0f - ng[i] * ng[i]);
          dau[i] = du * ugt[i] * (1.0f - ugt[i]);
          dar[i] = dan[i] * an[i] * rg[i] * (1.0f - rg[i]);
          danr[i] = dan[i] * rg[i];
          dh[i] *= ugt[i];                          // direct part for step s-1
        }
        if (s > 0 && row_ok) stt16(a.dhd, m, ug, E, dh);
        // ---- operand planes of the chunk for dh_{s-1} += dA_h W_hh (K index = unit inside the chunk) ----
        if (s > 0) {
          const int buf = cc & 1;
          mbar_wait(&a_empty[buf], (uint32_t)(((cc >> 1) & 1) ^ 1));
          uint8_t *ab = sA + buf * kABuf;
          const uint32_t o0 = sw64_off(L, 2 * sub), o1 = sw64_off(L, 2 * sub + 1);
#pragma unroll
          for (int g = 0; g < 3; ++g) {
            const float *src = g == 0 ? dar : (g == 1 ? dau : danr);
            const float v0[8] = {src[0], src[1], src[2], src[3], src[4], src[5], src[6], src[7]};
            const float v1[8] = {src[8], src[9], src[10], src[11], src[12], src[13], src[14], src[15]};
            uint4 h0, l0, h1, l1;
            split8(v0, h0, l0); split8(v1, h1, l1);
            *reinterpret_cast<uint4 *>(ab + g * kAGate + o0) = h0; *reinterpret_cast<uint4 *>(ab + g * kAGate + o1) = h1;
            *reinterpret_cast<uint4 *>(ab + 3 * kAGate + g * kAGate + o0) = l0; *reinterpret_cast<uint4 *>(ab + 3 * kAGate + g * kAGate + o1) = l1;
          }
          fence_before();
          fence_async_smem();
          __syncwarp();
          if (lane == 0) {
            asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(&a_full[buf])) : "memory");
          }
        }
        ++cc;
        // ---- gate gradients for the weight-gradient GEMMs: gate-interleaved columns, transposed through shared memory ----
        {
          uint4 hi6[6], lo6[6];
#pragma unroll
          for (int k = 0; k < 6; ++k) {
            float v[8];
#pragma unroll
            for (int x = 0; x < 8; ++x) {
              const int e = 8 * k + x, u = e / 3, g = e - 3 * u;
              v[x] = g == 0 ? dar[u] : (g == 1 ? dau[u] : danr[u]);
            }
            split8(v, hi6[k], lo6[k]);
          }
          __nv_bfloat16 *planes[2] = {(__nv_bfloat16 *)a.dah3_hi, (__nv_bfloat16 *)a.dah3_lo};
#pragma unroll
          for (int pp = 0; pp < 2; ++pp) {
            if (planes[pp] == nullptr) continue;
#pragma unroll
            for (int k = 0; k < 6; ++k) *reinterpret_cast<uint4 *>(stg + lane * kStgRow + 16 * k) = pp ? lo6[k] : hi6[k];
            __syncwarp();
#pragma unroll
            for (int k2 = 0; k2 < 6; ++k2) {
              const int pc = k2 * 32 + lane, r = pc / 6, piece = pc - 6 * r;
              const uint4 val = *reinterpret_cast<const uint4 *>(stg + r * kStgRow + 16 * piece);
              const size_t mr = row0 + r;
              if (mr < (size_t)a.M)
                *reinterpret_cast<uint4 *>(planes[pp] + ((size_t)s * a.M + mr) * 3 * E + 3 * ug + 8 * piece) = val;
            }
            __syncwarp();
          }
          // da_n: 16 units = 2 pieces per plane; staged as [hi0 hi1 lo0 lo1] per window
          {
            const float v0[8] = {dan[0], dan[1], dan[2], dan[3], dan[4], dan[5], dan[6], dan[7]};
            const float v1[8] = {dan[8], dan[9], dan[10], dan[11], dan[12], dan[13], dan[14], dan[15]};
            uint4 h0, l0, h1, l1;
            split8(v0, h0, l0); split8(v1, h1, l1);
            *reinterpret_cast<uint4 *>(stg + lane * kStgRow) = h0; *reinterpret_cast<uint4 *>(stg + lane * kStgRow + 16) = h1;
            *reinterpret_cast<uint4 *>(stg + lane * kStgRow + 32) = l0; *reinterpret_cast<uint4 *>(stg + lane * kStgRow + 48) = l1;
            __syncwarp();
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const int pc = k2 * 32 + lane, r = pc >> 2, piece = pc & 3;
              const uint4 val = *reinterpret_cast<const uint4 *>(stg + r * kStgRow + 16 * piece);
              const size_t mr = row0 + r;
              __nv_bfloat16 *dst = (__nv_bfloat16 *)(piece < 2 ? a.dan_hi : a.dan_lo);
              if (mr < (size_t)a.M && dst) *reinterpret_cast<uint4 *>(dst + ((size_t)s * a.M + mr) * E + ug + 8 * (piece & 1)) = val;
            }
            __syncwarp();
          }
        }
        // ---- bias gradients: column sums over the warp's windows, then shared-memory accumulators ----
        {
          float v[64];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            v[i] = row_ok ? dar[i] : 0.f; v[16 + i] = row_ok ? dau[i] : 0.f; v[32 + i] = row_ok ? dan[i] : 0.f; v[48 + i] = row_ok ? danr[i] : 0.f;
          }
          warp_reduce64(v, lane);
          const int idx = 2 * lane, type = idx >> 4, u = idx & 15;
          atomicAdd(&bacc[type * E + ug + u], v[0]);
          atomicAdd(&bacc[type * E + ug + u + 1], v[1]);
        }
      }
      if (tm) t_gate += clock64() - c1;
    }
    if (tm) printf("enc_persist bwd E=%d hist=%d: gate warp 0 per step: wait products %lld cycles, gate backward + stores %lld cycles\n", E, hist, t_wait / hist, t_gate / hist);
  }

  fence_before();
  __syncthreads();
  for (int i = tid; i < 4 * E; i += kThreads) {
    const int type = i / E, u = i - type * E;
    const float v = bacc[i];
    if (type == 0) { atomicAdd(a.gb_ih + u, v); atomicAdd(a.gb_hh + u, v); }
    else if (type == 1) { atomicAdd(a.gb_ih + E + u, v); atomicAdd(a.gb_hh + E + u, v); }
    else if (type == 2) atomicAdd(a.gb_ih + 2 * E + u, v);
    else atomicAdd(a.gb_hh + 2 * E + u, v);
  }
  if (warp == kEpiWarps + 1) {
    fence_after();
    tmem_dealloc(tmem, kTmemCols);
  }
}

bool bwd_supported(int E, int hist, size_t M, int mode) {
  if (!fwd_supported(E, hist, M, mode) || !env_flag("LFI_ENC_PERSIST_BWD", true)) return false;
  const BwdPlan p = bwd_plan(E);
  return p.stages >= 3 && p.total <= 227 * 1024;
}

int launch_bwd(const BwdArgs &a, cudaStream_t st) {
  LFI_REQUIRE(a.E == 128 || a.E == 256, LFI_ERR_SHAPE, "enc_persist bwd: E=%d unsupported", a.E);
  LFI_REQUIRE(a.hs && a.gates && a.ahn && a.whh_hi && (a.nplanes == 1 || a.whh_lo) && a.dhd && a.dah3_hi && a.dan_hi && a.gb_ih && a.gb_hh,
              LFI_ERR_ARG, "enc_persist bwd: null argument");
  const BwdPlan p = bwd_plan(a.E);
  LFI_REQUIRE(p.stages >= 3 && p.total <= 227 * 1024, LFI_ERR_SHAPE, "enc_persist bwd: shared-memory plan does not fit");
  CUtensorMap mhi, mlo;
  LFI_TRY(tc::make_plane_map(&mhi, a.whh_hi, 3 * a.E, a.E, a.E, 0, 1, 32));
  if (a.nplanes == 2) LFI_TRY(tc::make_plane_map(&mlo, a.whh_lo, 3 * a.E, a.E, a.E, 0, 1, 32));
  else mlo = mhi;
  static const bool timing = env_flag("LFI_ENC_TIMING", false);
  BwdArgs at = a;
  at.timing = timing ? 1 : 0;
  static bool attr_set = false;
  if (!attr_set) {
    LFI_CUDA(cudaFuncSetAttribute(enc_gru_bwd_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set = true;
  }
  const int ntiles = (a.M + kRows - 1) / kRows;
  const int smem = p.total < 120 * 1024 ? 120 * 1024 : p.total;
  enc_gru_bwd_persist<<<ntiles, kThreads, smem, st>>>(at, mhi, mlo);
  LFI_LAUNCH_CHECK();
  return LFI_OK;
}

// ------------------------------------------------------------------------------------------------
__global__ void add_deinterleaved_rows_kernel(float *out, const float *in, int E, int ngates, int cols) {
  const size_t n = (size_t)ngates * E * cols;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
    const int j = (int)(i % cols);
    const size_t r = i / cols;
    const int g = (int)(r / E), u = (int)(r - (size_t)g * E);
    out[i] += in[((size_t)3 * u + g) * cols + j];
  }
}
int add_deinterleaved_rows(float *out, const float *in, int E, int ngates, int cols, cudaStream_t st) {
  const size_t n = (size_t)ngates * E * cols;
  int blocks = (int)((n + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  add_deinterleaved_rows_kernel<<<blocks, 256, 0, st>>>(out, in, E, ngates, cols);
  LFI_LAUNCH_CHECK();
  return LFI_OK;
}

}  // namespace encp
}  // namespace lfi
