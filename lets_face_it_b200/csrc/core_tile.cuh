// Building blocks of the sequential flow core (core_fwd.cu, core_bwd.cu, core_sample.cu).
//
// One CTA owns a tile of R = 4*RPT sequences ("rows") and evaluates one flow-step cell
// (step k, frame t) — ActNorm -> 1x1 conv -> coupling RNN -> affine coupling
// (reference: FlowStep.normal_flow / reverse_flow, models.py:311-373) — as a chain of small
// [R x Kin] @ [Kin x N] products whose weights stream from L2 through a double-buffered cp.async
// stage while the activations stay in shared memory.
#pragma once
#include "lfi_common.cuh"

namespace lfi {
namespace core {

constexpr int NT = 256;        // threads per CTA
constexpr int TX = 64;         // threads along output columns
constexpr int RG = NT / TX;    // row groups
constexpr int CPT = 6;         // columns per thread per chunk
constexpr int WCH = TX * CPT;  // 384 columns per chunk
constexpr int KC16 = 16;       // reduction rows per weight stage (default ring)
constexpr int NSTG = 4;        // weight stages in flight (ring): covers the L2 latency of the stream

__device__ __forceinline__ void cp_async16(float *smem, const float *g) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(g) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

template <int RPT> __device__ __forceinline__ void load_rows(float (&a)[RPT], const float *p) {
  if constexpr (RPT == 8) {
    const float4 u = *reinterpret_cast<const float4 *>(p), v = *reinterpret_cast<const float4 *>(p + 4);
    a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w; a[4] = v.x; a[5] = v.y; a[6] = v.z; a[7] = v.w;
  } else if constexpr (RPT == 4) {
    const float4 u = *reinterpret_cast<const float4 *>(p);
    a[0] = u.x; a[1] = u.y; a[2] = u.z; a[3] = u.w;
  } else if constexpr (RPT == 2) {
    const float2 u = *reinterpret_cast<const float2 *>(p);
    a[0] = u.x; a[1] = u.y;
  } else {
    a[0] = p[0];
  }
}

// Activation ("act") layout in shared memory: act[i * RS + r], i = reduction index, r = row in tile,
// RS = R + 4 (keeps the float4 row reads 16-byte aligned and spreads transposed writes over banks).
template <int RPT> struct Tile {
  static constexpr int R = RG * RPT;
  static constexpr int RS = R + 4;
};

// out(r, j) = sum_i act[map(i)][r] * Wg[i*ldw + j]   for r < R, j < N.
// Wg: global, row pitch ldw (multiple of 4 floats, 16-byte aligned base, zero padded to a multiple of 4
// columns).  map(i) = i < split ? i + shift_lo : i + shift_hi  lets one product read two disjoint act row
// ranges (GRU backward) or a ring buffer (autoregressive window of the sampler).
// Every thread of the CTA must call this (it contains __syncthreads); act must have been written before
// the call (the first internal barrier orders it).  epi(r, j, v) is called once per owned output.
// KC = reduction rows per ring stage: 16, or 8 (half the ring: 48 KB instead of 96 KB, which lets the wide shapes run
// twice the rows per CTA; the bytes in flight, 3 stages x 12 KB, still cover the L2 latency of one SM's stream).
template <int RPT, int KC = KC16, class Epi>
__device__ __forceinline__ void tile_gemm(const float *act, const float *__restrict__ Wg, int ldw, int Kin, int N,
                                          float *wst, Epi epi, int split = 1 << 30, int shift_hi = 0, int shift_lo = 0) {
  constexpr int RS = Tile<RPT>::RS;
  const int tid = threadIdx.x, tx = tid % TX, ry = tid / TX;
  const int nst = (Kin + KC - 1) / KC;
  for (int n0 = 0; n0 < N; n0 += WCH) {
    const int nw = min(WCH, N - n0);
    const int nw4 = (nw + 3) >> 2;
    float acc[RPT][CPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r)
#pragma unroll
      for (int c = 0; c < CPT; ++c) acc[r][c] = 0.f;

    auto load_stage = [&](int buf, int s) {
      const int k0 = s * KC;
      const int kk = min(KC, Kin - k0);
      float *dst = wst + buf * (KC * WCH);
      for (int idx = tid; idx < kk * nw4; idx += NT) {
        const int i = idx / nw4, c4 = idx - i * nw4;
        cp_async16(dst + i * WCH + c4 * 4, Wg + (size_t)(k0 + i) * ldw + n0 + c4 * 4);
      }
    };
#pragma unroll
    for (int s = 0; s < NSTG - 1; ++s) {
      if (s < nst) load_stage(s, s);
      cp_async_commit();
    }
    for (int s = 0; s < nst; ++s) {
      cp_async_wait<NSTG - 2>();   // stage s has landed
      __syncthreads();             // ... for every thread; and everyone is done with the buffer of stage s-1
      if (s + NSTG - 1 < nst) load_stage((s + NSTG - 1) % NSTG, s + NSTG - 1);
      cp_async_commit();
      const float *w = wst + (s % NSTG) * (KC * WCH);
      const int k0 = s * KC;
      const int kk = min(KC, Kin - k0);
#pragma unroll 4
      for (int i = 0; i < kk; ++i) {
        int ai = k0 + i;
        ai = ai < split ? ai + shift_lo : ai + shift_hi;
        float a[RPT];
        load_rows<RPT>(a, act + ai * RS + ry * RPT);
        float wv[CPT];
#pragma unroll
        for (int c = 0; c < CPT; ++c) wv[c] = w[i * WCH + tx + TX * c];
#pragma unroll
        for (int r = 0; r < RPT; ++r)
#pragma unroll
          for (int c = 0; c < CPT; ++c) acc[r][c] = fmaf(a[r], wv[c], acc[r][c]);
      }
    }
    __syncthreads();  // the ring may be refilled (next column chunk / next product)
    cp_async_wait<0>();
#pragma unroll
    for (int c = 0; c < CPT; ++c) {
      const int j = n0 + tx + TX * c;
      if (j < N) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) epi(ry * RPT + r, j, acc[r][c]);
      }
    }
  }
}

// Skinny product: out(r, j) = sum_i act[i][r] * Wg[i*ldw + j] for N <= 64 output columns and a long reduction (the
// backward dz1 = dA_i W_ih[:, :Ci] with Kin = GH, LinearZeros with Kin = H).  tile_gemm would walk Kin / KC ring stages with a
// CTA barrier each for a handful of columns; here the reduction is split over the warps instead: thread = (column j, slice
// of Kin), weights straight from L2 (one coalesced row segment per warp and reduction index), activations as broadcast
// float4 reads, partial sums combined through `scr` (>= (NT / NP) * R * NP floats; the weight ring is free at that point).
// Every thread of the CTA must call this; act must be complete before the call (first barrier inside).
template <int RPT, class Epi>
__device__ __forceinline__ void skinny_gemm(const float *act, const float *__restrict__ Wg, int ldw, int Kin, int N, float *scr, Epi epi) {
  constexpr int R = Tile<RPT>::R, RS = Tile<RPT>::RS;
  const int tid = threadIdx.x;
  const int NP = N <= 32 ? 32 : 64, NS = NT / NP;
  const int j = tid % NP, sl = tid / NP;
  const int per = (Kin + NS - 1) / NS, k0 = sl * per, k1 = min(Kin, k0 + per);
  float acc[R];
#pragma unroll
  for (int r = 0; r < R; ++r) acc[r] = 0.f;
  __syncthreads();
  if (j < N) {
    for (int i0 = k0; i0 < k1; i0 += 8) {  // eight weight rows requested before the first is used
      float wv[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) wv[u] = (i0 + u < k1) ? __ldg(Wg + (size_t)(i0 + u) * ldw + j) : 0.f;
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const float *ap = act + min(i0 + u, k1 - 1) * RS;
#pragma unroll
        for (int q = 0; q < R / 4; ++q) {
          const float4 a4 = *reinterpret_cast<const float4 *>(ap + 4 * q);
          acc[4 * q] = fmaf(a4.x, wv[u], acc[4 * q]); acc[4 * q + 1] = fmaf(a4.y, wv[u], acc[4 * q + 1]);
          acc[4 * q + 2] = fmaf(a4.z, wv[u], acc[4 * q + 2]); acc[4 * q + 3] = fmaf(a4.w, wv[u], acc[4 * q + 3]);
        }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) scr[(sl * R + r) * NP + j] = acc[r];
  __syncthreads();
  for (int e = tid; e < R * N; e += NT) {
    const int r = e / N, c = e - r * N;
    float v = 0.f;
    for (int s2 = 0; s2 < NS; ++s2) v += scr[(s2 * R + r) * NP + c];
    epi(r, c, v);
  }
  __syncthreads();  // scr (the weight ring) may be refilled
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Per-step weight views inside the derived cache (all pitches multiples of 4 floats).
struct StepWeights {
  const float *an_bias, *an_logs;  // [C]
  const float *Wfwd;               // [C][Cp]    W[i][j]        (z = y @ W)
  const float *WT;                 // [C][Cp]    W[j][i]        (dy = dz @ W^T)
  const float *Winv;               // [C][Cp]    W^-1[i][j]
  const float *WzT;                // [Ci][GH]   W_ih[j][i], i < Ci
  const float *WihZ;               // [GH][Cip]  W_ih[j][i], i < Ci  (dz1 = dA_i @ W_ih[:, :Ci])
  const float *WhhT;               // [H][GH]    W_hh[j][i]
  const float *Whh;                // [GH][H]    raw
  const float *b_ih, *b_hh;        // [GH]
  const float *WfT;                // [H][Cop]   Wf[j][i]
  const float *Wf;                 // [Co][H]    raw
  const float *bf, *lf;            // [Co]
  const float *WcArT;              // [Far][D]   Wc[d][i], i < Far       (sampler: AR part of cond_transform)
  const float *WihCT;              // [D][GH]    W_ih[j][Ci + d]         (sampler: gate-ih on c)
};

// Whole-model table handed to the kernels by value.
struct DerivedView {
  const float *an_bias, *an_logs, *Wfwd, *WT, *Winv, *WzT, *WihZ, *WhhT, *Whh, *b_ih, *b_hh, *WfT, *Wf, *bf, *lf,
      *WcArT, *WihCT;
  __device__ __forceinline__ StepWeights step(const Dims &d, int k) const {
    StepWeights w;
    w.an_bias = an_bias + (size_t)k * d.C; w.an_logs = an_logs + (size_t)k * d.C;
    w.Wfwd = Wfwd + (size_t)k * d.C * d.Cp; w.WT = WT + (size_t)k * d.C * d.Cp;
    w.Winv = Winv ? Winv + (size_t)k * d.C * d.Cp : nullptr;
    w.WzT = WzT + (size_t)k * d.Ci * d.GH; w.WihZ = WihZ + (size_t)k * d.GH * d.Cip;
    w.WhhT = WhhT + (size_t)k * d.H * d.GH; w.Whh = Whh + (size_t)k * d.GH * d.H;
    w.b_ih = b_ih + (size_t)k * d.GH; w.b_hh = b_hh + (size_t)k * d.GH;
    w.WfT = WfT + (size_t)k * d.H * d.Cop; w.Wf = Wf + (size_t)k * d.Co * d.H;
    w.bf = bf + (size_t)k * d.Co; w.lf = lf + (size_t)k * d.Co;
    w.WcArT = WcArT ? WcArT + (size_t)k * d.Far * d.D : nullptr;
    w.WihCT = WihCT ? WihCT + (size_t)k * d.D * d.GH : nullptr;
    return w;
  }
};

// Shared-memory carve-up shared by the forward / inverse / backward cell kernels (offsets in floats).
struct SmemPlan {
  int wst, xs, zact, zrow, hp, cp, S, ahn, orow, ldacc;  // common
  int cact, hist;                                         // sampler (autoregressive) only
  int dact, dxr, dhr, prod, cn;                           // backward only
  int total;
};

// odd pitch => conflict-free column walks over row-major [R][pitch] arrays
__host__ __device__ inline int odd(int n) { return n | 1; }

__host__ __device__ inline SmemPlan plan_smem(const Dims &d, int R, bool bwd, bool sampler, int kc = KC16) {
  const int RS = R + 4;
  const int Cm = d.Co > d.C ? d.Co : d.C;
  SmemPlan p;
  int o = 0;
  auto take = [&](int n) { int r = o; o += round_up(n, 4); return r; };
  p.wst = take(NSTG * kc * WCH);
  p.xs = take(Cm * RS);                        // act: ActNorm output y / inverse: coupling output / bwd: dlin
  p.zact = take(d.C * RS);                     // act: z1 (fwd) / dzf (bwd)
  p.zrow = take(R * odd(d.C));                 // row-major z (fwd) / zf -> dzf (bwd)
  p.hp = take(d.H * RS);                       // act: h_prev -> h_new
  p.cp = d.G == 4 ? take(d.H * RS) : 0;        // act: c_prev -> c_new (LSTM)
  p.S = take(R * odd(d.GH));                   // row-major gate pre-activations -> gates -> dA_i
  p.ahn = take(R * odd(d.H));                  // row-major GRU h-side n pre-activation -> (dA_h)_n; LSTM bwd: dc
  p.orow = take(R * odd(Cm));                  // row-major LinearZeros output -> dO
  p.ldacc = take(R);                           // per-row running log-det (inverse kernel)
  p.cact = sampler ? take(d.D * RS) : 0;       // act: c = LeakyReLU(cond_transform)
  p.hist = sampler ? take(d.Far * RS) : 0;     // act: ring of the last hist[0] generated frames
  p.dact = bwd ? take((d.GH + d.H) * RS) : 0;  // act: dA_i rows [0,GH) + GRU (dA_h)_n rows [GH,GH+H)
  p.dxr = bwd ? take(R * odd(d.C)) : 0;        // row-major d(output of the step) -> d(input of the step)
  p.dhr = bwd ? take(R * odd(d.H)) : 0;        // row-major dh -> dh_prev
  p.prod = bwd ? take(R * odd(Cm)) : 0;        // row-major products for the logs gradients
  p.cn = (bwd && d.G == 4) ? take(d.H * RS) : 0;  // act: c_new (LSTM backward)
  p.total = o;
  return p;
}

}  // namespace core
}  // namespace lfi
