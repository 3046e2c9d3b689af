// Persistent AR decode kernel: the whole token loop of models/vallex.py:528-598 in ONE cooperative launch.
//
// One CTA per SM (512 threads), all resident; every decode step walks
//     S0  sample + append + embed (CTA b serves utterance b)              models/vallex.py:529-531,569-598
//     12 x [ P1 LN1 + QKV rows | P2 attention (KV append, split-K over CTAs, ticket combine)
//            | P3 out-proj + residual | P4 LN2 + FFN1 + ReLU | P5 FFN2 (split-K x4, partial sums) ]
//     PF  final LN + predict rows -> logits                               modules/transformer.py:337-347,470-471
// separated by grid-wide barriers (release/acquire counter in L2), 62 per step.  Nothing returns to the host
// between tokens: sampling, EOS/budget bookkeeping and the all-finished early exit live on the device.
//
// Data path (HBM bound): each CTA owns a fixed, contiguous slice of the rows of every weight matrix
// (row-major, so a slice is one contiguous 40-60 KB range); a warp streams one row at a time with 128-bit
// coalesced loads (4-8 in flight per lane), activations (fp32, B x 1024) sit in shared memory, the reduction
// is a warp shuffle.  While a phase runs, the CTA's slice for the NEXT phase is pulled into L2 with
// cp.async.bulk.prefetch.L2 so HBM keeps streaming across the barriers.  The KV cache ([b][h][pos][64], 128 B
// rows in bf16) is read once per step with 128-bit loads; this step's key/value are appended in place.
// FFN2 is split over K so every GEMV phase has the same shape (~21-28 rows x 1024); the four partial sums are
// folded (fixed order, deterministic) by the consumer's prologue together with bias and residual.
//
// Algorithmic bytes per step: 304,412,672 + 49,152 * sum_b(L_b + 1) at bf16.
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "decode_persist.h"
#include "tc_common.cuh"

namespace vx {

constexpr int PNT = 512;            // threads per CTA
constexpr int PNW = PNT / 32;       // warps
constexpr int PCHUNK = 2048;        // max keys per attention unit
constexpr int PSORT = 2048;

// ---- grid barrier -------------------------------------------------------------------------------------
// `prof` (nullable): two clock64 stamps per barrier -- arrival of this CTA, release of the grid
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch, long long*& prof) {
  __syncthreads();
  if (threadIdx.x == 0) {
    if (prof) { prof[0] = 1; prof[1] = clock64(); prof += 2; }
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int v;
    unsigned int spins = 0;
    do {  // relaxed polling (no per-iteration L1 invalidate)
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 24)) __trap();  // a lost CTA must not hang the box
    } while ((int)(v - epoch) < 0);
    // no acquire fence: every cross-CTA mutable value is read with ld.global.cg (L2), never through L1
    if (prof) { prof[0] = 2; prof[1] = clock64(); prof += 2; }
  }
  __syncthreads();
}
#define PSTAMP(id_) do { if (prof) { prof[0] = (id_); prof[1] = clock64(); prof += 2; } } while (0)

__device__ __forceinline__ void l2_prefetch(const void* p, size_t bytes) {
  const char* c = reinterpret_cast<const char*>(p);
  // 16-byte aligned ranges, chunks of <= 32 KB
  size_t start = reinterpret_cast<size_t>(c) & ~size_t(15);
  size_t end = (reinterpret_cast<size_t>(c) + bytes + 15) & ~size_t(15);
  while (start < end) {
    const unsigned int n = (unsigned int)((end - start) < 32768 ? (end - start) : 32768);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(start), "r"(n) : "memory");
    start += n;
  }
}

__device__ __forceinline__ int ld_cg(const int32_t* p) { return __ldcg(p); }
// KV rows are written during the kernel (append) -> read them through L2 (ld.global.cg), never the .nc path
__device__ __forceinline__ void ldkv8(const bf16* p, float (&o)[8]) {
  const uint4 r = __ldcg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ldkv8(const float* p, float (&o)[8]) {
  const float4 a = __ldcg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldcg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

__device__ __forceinline__ void slice(int total, int& r0, int& r1) {
  r0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  r1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);
}

// 128-bit streaming weight loads (bypass L1: each byte is used once per step)
__device__ __forceinline__ void ldw8(const bf16* p, float (&o)[8]) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void ldw8(const float* p, float (&o)[8]) {
  uint4 a, b;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p));
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(p + 4));
  o[0] = __uint_as_float(a.x); o[1] = __uint_as_float(a.y); o[2] = __uint_as_float(a.z); o[3] = __uint_as_float(a.w);
  o[4] = __uint_as_float(b.x); o[5] = __uint_as_float(b.y); o[6] = __uint_as_float(b.z); o[7] = __uint_as_float(b.w);
}

// ---- weight-streaming GEMV with cross-barrier register prefetch ---------------------------------------------
// A CTA's row slice of one matrix is at most 2 rows per warp (148 CTAs x 16 warps), so a phase's weights fit in
// registers: they are requested BEFORE the preceding grid barrier (the weights do not depend on it) and consumed
// after it, which takes the HBM latency of every GEMV phase off the critical path.
template <typename WT, int MR>
struct WPre {
  static constexpr int kRows = (sizeof(WT) == 2 && MR <= 8) ? 2 : 1;
  static constexpr int kVec = (sizeof(WT) == 2) ? 4 : 8;  // uint4 per row per lane (K = 1024)
  uint4 r[kRows][kVec];
  float bias[kRows];  // constant addend of each prefetched row (bias [+ previous layer's FFN2 bias]), fetched with the row
};
struct LnPre {        // LayerNorm parameters (+ previous layer's FFN2 bias) of this thread's two columns
  float g0, g1, b0, b1, f0, f1;
};
// Before a barrier: pull the NEXT LayerNorm's vectors (gamma, beta, previous FFN2 bias; 12 KB) towards L2 -- no
// registers are held across the barrier.  ln_load (after the barrier) then hits L2 and overlaps the activation loads.
__device__ __forceinline__ void ln_issue(LnPre&, const float* __restrict__ g, const float* __restrict__ b,
                                         const float* __restrict__ fb) {
  const int t = threadIdx.x;
  if (t < 32) asm volatile("prefetch.global.L2 [%0];" ::"l"(g + t * 32));
  else if (t < 64) asm volatile("prefetch.global.L2 [%0];" ::"l"(b + (t - 32) * 32));
  else if (t < 96 && fb) asm volatile("prefetch.global.L2 [%0];" ::"l"(fb + (t - 64) * 32));
}
__device__ __forceinline__ void ln_load(LnPre& lp, const float* __restrict__ g, const float* __restrict__ b,
                                        const float* __restrict__ fb) {
  const int t = threadIdx.x;
  lp.g0 = g[t]; lp.g1 = g[t + PNT]; lp.b0 = b[t]; lp.b1 = b[t + PNT];
  lp.f0 = fb ? fb[t] : 0.f; lp.f1 = fb ? fb[t + PNT] : 0.f;
}

__device__ __forceinline__ uint4 ldw_raw(const void* p) {
  uint4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
template <typename WT>
__device__ __forceinline__ void raw_to_f32(const uint4* v, int it, float (&o)[8]);
template <>
__device__ __forceinline__ void raw_to_f32<bf16>(const uint4* v, int it, float (&o)[8]) {
  const uint32_t w[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
template <>
__device__ __forceinline__ void raw_to_f32<float>(const uint4* v, int it, float (&o)[8]) {
  o[0] = __uint_as_float(v[2 * it].x); o[1] = __uint_as_float(v[2 * it].y); o[2] = __uint_as_float(v[2 * it].z); o[3] = __uint_as_float(v[2 * it].w);
  o[4] = __uint_as_float(v[2 * it + 1].x); o[5] = __uint_as_float(v[2 * it + 1].y); o[6] = __uint_as_float(v[2 * it + 1].z); o[7] = __uint_as_float(v[2 * it + 1].w);
}

template <typename WT, int MR>
__device__ __forceinline__ void load_row_raw(const WT* __restrict__ W, size_t ldw, int n, uint4 (&dst)[WPre<WT, MR>::kVec]) {
  const int lane = threadIdx.x & 31;
  const WT* p = W + (size_t)n * ldw + lane * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if constexpr (sizeof(WT) == 2) dst[it] = ldw_raw(p + it * 256);
    else { dst[2 * it] = ldw_raw(p + it * 256); dst[2 * it + 1] = ldw_raw(p + it * 256 + 4); }
  }
}

// request this warp's first rows of slice [r0,r1)
template <typename WT, int MR>
__device__ __forceinline__ void gemv_issue(WPre<WT, MR>& pre, const WT* __restrict__ W, size_t ldw, int r0, int r1,
                                           const float* __restrict__ bias, const float* __restrict__ bias2) {
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < WPre<WT, MR>::kRows; ++j) {
    const int n = r0 + warp + j * PNW;
    pre.bias[j] = 0.f;
    if (n < r1) {
      load_row_raw<WT, MR>(W, ldw, n, pre.r[j]);
      if (bias) pre.bias[j] = bias[n];
      if (bias2) pre.bias[j] += bias2[n];
    }
  }
}

template <typename WT, int MR>
__device__ __forceinline__ void row_dot(const uint4 (&raw)[WPre<WT, MR>::kVec], const float* __restrict__ sx, float (&acc)[MR]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    float w[8];
    raw_to_f32<WT>(raw, it, w);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
      acc[m] = fmaf(w[0], x0.x, acc[m]); acc[m] = fmaf(w[1], x0.y, acc[m]); acc[m] = fmaf(w[2], x0.z, acc[m]); acc[m] = fmaf(w[3], x0.w, acc[m]);
      acc[m] = fmaf(w[4], x1.x, acc[m]); acc[m] = fmaf(w[5], x1.y, acc[m]); acc[m] = fmaf(w[6], x1.z, acc[m]); acc[m] = fmaf(w[7], x1.w, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = warp_sum(acc[m]);
}

// consume the prefetched rows (and any further rows of the slice).  addend(n, m) is evaluated by lane m BEFORE the
// dot products (its loads overlap them); emit(n, m, value) is called by lane m for batch row m < B.
template <typename WT, int MR, typename Add, typename Emit>
__device__ __forceinline__ void gemv_finish(WPre<WT, MR>& pre, const WT* __restrict__ W, size_t ldw, int r0, int r1,
                                            const float* __restrict__ sx, int B, const float* __restrict__ bias,
                                            const float* __restrict__ bias2, Add addend, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kRows = WPre<WT, MR>::kRows;
  float add[kRows];
#pragma unroll
  for (int j = 0; j < kRows; ++j) {
    const int n = r0 + warp + j * PNW;
    add[j] = (n < r1 && lane < B) ? addend(n, lane) : 0.f;
  }
  if constexpr (kRows == 2) {
    // both prefetched rows in ONE pass over x: the shared-memory reads of the activations (the bottleneck of a
    // 16-warp GEMV phase) are shared by the two rows
    const int n0 = r0 + warp, n1 = r0 + warp + PNW;
    if (n0 < r1) {  // warp-uniform
      float a0[MR], a1[MR];
#pragma unroll
      for (int m = 0; m < MR; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        float w0[8], w1[8];
        raw_to_f32<WT>(pre.r[0], it, w0);
        raw_to_f32<WT>(pre.r[1], it, w1);
#pragma unroll
        for (int m = 0; m < MR; ++m) {
          const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
          const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
          const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
          for (int i = 0; i < 8; ++i) { a0[m] = fmaf(w0[i], xv[i], a0[m]); a1[m] = fmaf(w1[i], xv[i], a1[m]); }
        }
      }
      float m0 = 0.f, m1 = 0.f;
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float s0 = warp_sum(a0[m]), s1 = warp_sum(a1[m]);
        if (lane == m) { m0 = s0; m1 = s1; }
      }
      if (lane < B) {
        emit(n0, lane, m0 + (add[0] + pre.bias[0]));
        if (n1 < r1) emit(n1, lane, m1 + (add[1] + pre.bias[1]));
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < kRows; ++j) {
      const int n = r0 + warp + j * PNW;
      if (n < r1) {
        float acc[MR];
        row_dot<WT, MR>(pre.r[j], sx, acc);
        float mine = 0.f;
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (lane == m) mine = acc[m];
        if (lane < B) emit(n, lane, mine + (add[j] + pre.bias[j]));
      }
    }
  }
  for (int n = r0 + warp + kRows * PNW; n < r1; n += PNW) {  // only when the grid is smaller than 148 CTAs
    uint4 raw[WPre<WT, MR>::kVec];
    load_row_raw<WT, MR>(W, ldw, n, raw);
    const float ad = ((lane < B) ? addend(n, lane) : 0.f) + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f);
    float acc[MR];
    row_dot<WT, MR>(raw, sx, acc);
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m)
      if (lane == m) mine = acc[m];
    if (lane < B) emit(n, lane, mine + ad);
  }
}

// Stage B rows into shared memory with LayerNorm (eps 1e-5), all 512 threads, two-pass variance.
// value(m, c) is the dynamic part of the stream; lp.f* (previous layer's FFN2 bias) is added here.
template <int MR, typename Val>
__device__ __forceinline__ void stage_ln(float* sx, int B, Val value, const LnPre& lp, float* red /*[2][PNW][8]*/) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = MR < 8 ? MR : 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v0[R], v1[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      v0[m] = (m < nb) ? value(mb + m, tid) + lp.f0 : 0.f;
      v1[m] = (m < nb) ? value(mb + m, tid + PNT) + lp.f1 : 0.f;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float s = warp_sum(v0[m] + v1[m]);
      if (lane == 0) red[warp * 8 + m] = s;
    }
    __syncthreads();
    float mean[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PNW; ++w) s += red[w * 8 + m];
      mean[m] = s / (float)kD;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float d0 = v0[m] - mean[m], d1 = v1[m] - mean[m];
      const float s = warp_sum(d0 * d0 + d1 * d1);
      if (lane == 0) red[PNW * 8 + warp * 8 + m] = s;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) s += red[PNW * 8 + w * 8 + m];
        const float rstd = 1.0f / sqrtf(s / (float)kD + 1e-5f);
        sx[(mb + m) * kD + tid] = __fadd_rn(__fmul_rn((v0[m] - mean[m]) * rstd, lp.g0), lp.b0);
        sx[(mb + m) * kD + tid + PNT] = __fadd_rn(__fmul_rn((v1[m] - mean[m]) * rstd, lp.g1), lp.b1);
      }
    }
    __syncthreads();  // red[] is reused by the next chunk; sx complete for the consumers
  }
}
template <typename Val>
__device__ __forceinline__ void stage_plain(float* sx, int B, Val value) {
  for (int i = threadIdx.x; i < B * kD; i += PNT) sx[i] = value(i / kD, i % kD);
  __syncthreads();
}

// residual stream entering a layer: xa (+ previous layer's FFN2 bias and 4 split-K partials, fixed order)
// (the previous layer's FFN2 bias, a constant, is added from prefetched registers by the callers)
__device__ __forceinline__ float stream_dyn(const PParams& p, int l, int b, int c, int B) {
  float v = __ldcg(p.xa + (size_t)b * kD + c);
  if (l > 0) {
    float s4 = 0.f;
#pragma unroll
    for (int s = 0; s < 4; ++s) s4 += __ldcg(p.pf + ((size_t)s * B + b) * kD + c);
    v += s4;
  }
  return v;
}

// ---- sampling (same arithmetic as sample_kernel in kernels_basic.cu; models/vallex.py:791-853) --------
__device__ __forceinline__ float p_philox(uint64_t seed, uint32_t c0, uint32_t c1) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = 0x9E3779B9u, x3 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = __umulhi(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
    const uint32_t h1 = __umulhi(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
    x0 = h1 ^ x1 ^ k0; x1 = l1; x2 = h0 ^ x3 ^ k1; x3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ((float)(x0 >> 8) + 0.5f) * (1.0f / 16777216.0f);
}

__device__ void sample_utterance(const PParams& p, int b, float* smem) {
  float* v = smem;                          // [1032]
  float* sv = smem + 1032;                  // [2048]
  int16_t* sidx = reinterpret_cast<int16_t*>(sv + PSORT);  // [2048]
  float* scan = reinterpret_cast<float*>(sidx + PSORT);    // [512]
  float* red = scan + PNT;                  // [32]
  int* s_tok = reinterpret_cast<int*>(red + 32);
  const Batch& bt = p.bt;
  const int tid = threadIdx.x;
  const float* lg = p.logits + (size_t)b * kARLogits;
  for (int i = tid; i < kARLogits; i += PNT) {
    const float x = __ldcg(lg + i);
    v[i] = (p.temperature != 1.0f) ? x / p.temperature : x;
  }
  __syncthreads();
  const bool need_sort = (p.top_k > 1) || (p.top_p < 1.0f);
  const bool greedy = (p.top_k == 1);
  int g_tok = -1;
  float g_logprob = 0.f;
  if (greedy) {
    // top_k == 1 keeps only the maximum (models/vallex.py:803-809): the draw is its index (first one on an exact tie)
    // and log_softmax over the kept entries is -log(#ties).
    float m = -INFINITY;
    for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
    m = block_max(m, red);
    int first = 0x7fffffff, cnt = 0;
    for (int i = tid; i < kARLogits; i += PNT)
      if (v[i] == m) { first = min(first, i); ++cnt; }
    first = -(int)block_max((float)(-first), red);   // indices < 2^24: exact in fp32
    cnt = (int)block_sum((float)cnt, red);
    g_tok = first;
    g_logprob = -logf((float)cnt);
  }
  if (need_sort) {
    for (int i = tid; i < PSORT; i += PNT) { sv[i] = (i < kARLogits) ? v[i] : -INFINITY; sidx[i] = (int16_t)i; }
    __syncthreads();
    for (int k = 2; k <= PSORT; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < PSORT; i += PNT) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const bool desc = ((i & k) == 0);
            const float a = sv[i], c = sv[ixj];
            if (desc ? (a < c) : (a > c)) {
              sv[i] = c; sv[ixj] = a;
              const int16_t t = sidx[i]; sidx[i] = sidx[ixj]; sidx[ixj] = t;
            }
          }
        }
        __syncthreads();
      }
    if (p.top_k > 1) {
      const float thr = sv[min(p.top_k, kARLogits) - 1];
      for (int i = tid; i < kARLogits; i += PNT)
        if (v[i] < thr) v[i] = -INFINITY;
      for (int i = tid; i < PSORT; i += PNT)
        if (sv[i] < thr) sv[i] = -INFINITY;
      __syncthreads();
    }
    if (p.top_p < 1.0f) {
      const float m = sv[0];
      float part = 0.f;
      for (int i = tid; i < kARLogits; i += PNT) part += __expf(sv[i] - m);
      const float Z = block_sum(part, red);
      if (tid == 0) {
        float c = 0.f;
        bool drop_next = false;
        for (int i = 0; i < kARLogits; ++i) {
          const bool drop = drop_next;
          c += __expf(sv[i] - m) / Z;
          drop_next = (c > p.top_p);
          if (drop) v[sidx[i]] = -INFINITY;
        }
      }
      __syncthreads();
    }
  }
  float m = -INFINITY;
  if (!greedy) {
  for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
  m = block_max(m, red);
  }
  constexpr int CH = (kARLogits + PNT - 1) / PNT;  // 3
  float Z = 1.f;
  const int Tb = ld_cg(bt.T + b);
  if (!greedy) {
  float loc[CH], s = 0.f;
#pragma unroll
  for (int q = 0; q < CH; ++q) {
    const int i = tid * CH + q;
    loc[q] = (i < kARLogits) ? expf(v[i] - m) : 0.f;
    s += loc[q];
  }
  scan[tid] = s;
  __syncthreads();
  if (tid < 32) {
    constexpr int PER = PNT / 32;
    float run = 0.f, mine[PER];
#pragma unroll
    for (int q = 0; q < PER; ++q) { mine[q] = scan[tid * PER + q]; run += mine[q]; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += n;
    }
    float ex = incl - run;
#pragma unroll
    for (int q = 0; q < PER; ++q) { scan[tid * PER + q] = ex; ex += mine[q]; }
    if (tid == 31) red[0] = incl;
  }
  if (tid == 0) *s_tok = -1;
  __syncthreads();
  Z = red[0];
  const float u = p_philox(p.seed, (uint32_t)Tb, (uint32_t)b) * Z;
  {
    float c = scan[tid];
    int cand = -1;
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      const float n = c + loc[q];
      if (cand < 0 && loc[q] > 0.f && u <= n && u > c) cand = tid * CH + q;
      c = n;
    }
    if (cand >= 0) atomicMax(s_tok, cand);
  }
  __syncthreads();
  } else {
    if (tid == 0) *s_tok = g_tok;
    __syncthreads();
  }
  if (tid == 0) {
    int tok = *s_tok;
    if (tok < 0)
      for (int i = kARLogits - 1; i >= 0; --i)
        if (v[i] > -INFINITY) { tok = i; break; }
    const float logprob = greedy ? g_logprob : (v[tok] - m) - logf(Z);
    const int prev = ld_cg(bt.y + (size_t)b * bt.ymax + bt.P[b] + Tb);
    if (prev != kEOS) bt.sum_logprobs[b] += logprob;
    p.out_tokens[(size_t)b * p.out_stride + Tb] = tok;
    const bool eos = (tok == kEOS) && !p.ignore_eos;
    if (eos) {
      bt.finished[b] = 1;
      atomicSub(bt.n_active, 1);
    } else {
      const int fed = p.forced ? p.forced[(size_t)b * p.forced_stride + Tb] : tok;
      bt.y[(size_t)b * bt.ymax + bt.P[b] + 1 + Tb] = fed;
      bt.T[b] = Tb + 1;
      if (Tb + 1 >= bt.budget[b]) {
        bt.finished[b] = 1;
        atomicSub(bt.n_active, 1);
      } else {
        // next input row: emb[token] + alpha * pe[y position]   (models/vallex.py:529-531)
        *s_tok = fed;
      }
    }
    if (bt.finished[b]) *s_tok = -1;   // written by this thread just above
    s_tok[1] = bt.T[b];
  }
  __syncthreads();
  const int fed = *s_tok;
  if (fed >= 0) {
    const int j = bt.P[b] + s_tok[1];
    const float a = p.alpha[0];
    for (int c = tid; c < kD; c += PNT)
      p.xa[(size_t)b * kD + c] = __fadd_rn(__fmul_rn(p.audio_emb[(size_t)fed * kD + c], 1.0f), __fmul_rn(a, p.pe[(size_t)j * kD + c]));
  }
  __syncthreads();
}

// ---- decode attention unit (b, h, split): same arithmetic as attn_decode_kernel ------------------------
template <typename KVT> __device__ __forceinline__ float kv_round(float v);
template <> __device__ __forceinline__ float kv_round<float>(float v) { return v; }
template <> __device__ __forceinline__ float kv_round<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ float att_combined(const PParams& p, int b, int c);

template <typename KVT>
__device__ void attn_unit(const PParams& p, int layer, int b, int h, int split, float* smem) {
  float* sq = smem;                  // [64]
  float* sknew = smem + 64;          // [64]
  float* svnew = smem + 128;         // [64]
  float* red = smem + 192;           // [32]
  int* s_last = reinterpret_cast<int*>(smem + 224);
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();  // smem reuse across units
  if (tid < kDH) {
    sq[tid] = __ldcg(row + tid) * 0.125f;
    sknew[tid] = __ldcg(row + kD + tid);
    svnew[tid] = __ldcg(row + 2 * kD + tid);
  }
  __syncthreads();
  KVT* kb = reinterpret_cast<KVT*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  KVT* vb = reinterpret_cast<KVT*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<KVT>(sknew[d]), s);
    } else {
      const KVT* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<KVT>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
  // partial (m, l, o[64]) of this split.  The LAST split of (b, h) to arrive (ticket) folds all of them in fixed split
  // order -- deterministic whoever is last -- into att[b][h*64 ..], so the out-proj phase stages 4 KB per utterance
  // instead of every CTA gathering all partials.
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  if (p.dbg & 2) return;   // bisecting aid: the out-proj phase folds the splits itself
  __syncthreads();
  if (tid == 0) {
    __threadfence();       // the CTA's partial is visible device-wide before the ticket is taken
    *s_last = (atomicAdd(p.tickets + b * kH + h, 1) == p.nsplit - 1);
  }
  __syncthreads();
  if (*s_last) {
    __threadfence();
    if (tid < kDH) p.att[(size_t)b * kD + h * kDH + tid] = att_combined(p, b, h * kDH + tid);
    if (tid == 0) p.tickets[b * kH + h] = 0;   // next use is behind a grid barrier
  }
}

// softmax-combine the key splits of head c/64 for column c (fixed order -> deterministic)
__device__ __forceinline__ float att_combined(const PParams& p, int b, int c) {
  const int h = c >> 6, d = c & 63;
  const int ns = p.nsplit;
  const float* p0 = p.part + ((size_t)b * kH + h) * ns * (kDH + 2);
  constexpr int U = 12;  // splits folded with every load in flight at once (one L2 round trip instead of 3*ns)
  float mv[U], lv[U], ov[U];
#pragma unroll
  for (int s = 0; s < U; ++s) {
    const bool ok = s < ns;
    mv[s] = ok ? __ldcg(p0 + s * (kDH + 2)) : -INFINITY;
    lv[s] = ok ? __ldcg(p0 + s * (kDH + 2) + 1) : 0.f;
    ov[s] = ok ? __ldcg(p0 + s * (kDH + 2) + 2 + d) : 0.f;
  }
  float M = -INFINITY;
#pragma unroll
  for (int s = 0; s < U; ++s) M = fmaxf(M, mv[s]);
  for (int s = U; s < ns; ++s) M = fmaxf(M, __ldcg(p0 + s * (kDH + 2)));
  float L = 0.f, O = 0.f;
#pragma unroll
  for (int s = 0; s < U; ++s) {
    const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
    L += w * lv[s];
    O += w * ov[s];
  }
  for (int s = U; s < ns; ++s) {
    const float ms = __ldcg(p0 + s * (kDH + 2));
    const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
    L += w * __ldcg(p0 + s * (kDH + 2) + 1);
    O += w * __ldcg(p0 + s * (kDH + 2) + 2 + d);
  }
  return O / L;
}

// ---- the kernel -------------------------------------------------------------------------------------------
template <typename WT, int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_persistent_kernel(const PParams p) {
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* scratch = dsm + MR * kD;        // attention / sampling / LN-reduction scratch
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;
  const size_t wsz = sizeof(WT);

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // FFN2 split-K: CTA -> (k chunk kc, row slice)
  const int per = gridDim.x / 4;                   // CTAs per k chunk (the last chunk takes the remainder)
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  WPre<WT, MR> pre;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  // everything that does not depend on a barrier is requested BEFORE it: weight rows, their biases, LN parameters
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  // optional timeline: CTAs 0, grid/2 and grid-1 stamp every barrier of steps [prof_step0, prof_step0 + 4)
  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed -------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      // ---- P1: LN1 + QKV rows; pull this CTA's K/V ranges of the coming attention phase into L2 ----------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        if (tid == 0) {
          const int split = u % p.nsplit, bh = u / p.nsplit, b = bh / kH, h = bh % kH;
          const int n = bt.S[b] + bt.P[b] + ld_cg(bt.T + b) + 1;
          int chunk = (n + p.nsplit - 1) / p.nsplit;
          chunk = (chunk + 31) & ~31;
          const int j0 = split * chunk, j1 = min(n - 1, j0 + chunk);
          if (j1 > j0) {
            const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * wsz;
            l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * wsz);
            l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * wsz);
          }
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, lp, scratch);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, Wo, kD, o0, o1, L.out_b, fb_prev);   // out-proj rows ride through the attention phase
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention ---------------------------------------------------------------------------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        if (sizeof(WT) == 2) attn_unit<bf16>(p, l, b, h, split, scratch);
        else attn_unit<float>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: (combine key splits) + out-proj + residual -> xa ------------------------------------------------
      stage_plain(sx, B, [&](int m, int c) {
        if (ld_cg(bt.finished + m)) return 0.f;
        return (p.dbg & 2) ? att_combined(p, m, c) : __ldcg(p.att + (size_t)m * kD + c);
      });
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wo, kD, o0, o1, sx, B, L.out_b, fb_prev, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                          [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      ln_issue(lp, L.n2_g, L.n2_b, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU -------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, nullptr);
      stage_ln<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W2, kFF, g0, g1, nullptr, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2, split-K: rows [g0,g1) x k chunk kc -> partial pf[kc] --------------------------------------
      stage_plain(sx, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W2, kFF, g0, g1, sx, B, nullptr, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
      PSTAMP(4);
      if (l + 1 < kL) {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1, p.layer[l + 1].qkv_b, nullptr);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: final LayerNorm + predict rows -> logits ---------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, lp, scratch);
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
  (void)wsz;
}

// =================================================================================================================
// Ring variant (bf16 weights, B <= 8): the CTA's weight-row slice of every GEMV phase -- plus that phase's LayerNorm
// vectors and biases -- is staged into shared memory by the TMA engine (cp.async.bulk -> mbarrier, 1-D bulk copies of
// the contiguous slice) TWO phases ahead through a 2-slot ring, so HBM streams continuously across the grid barriers
// and no phase waits on DRAM latency; the dot products read weights from shared memory with conflict-free 128-bit loads.
// =================================================================================================================

constexpr int RW_ROWS = 32;
constexpr int RW_BYTES = RW_ROWS * 2048;             // weight rows (K = 1024 bf16)
constexpr int RLN_OFF = RW_BYTES;                    // gamma | beta | previous FFN2 bias (3 x 4 KB)
constexpr int RBIAS_OFF = RLN_OFF + 3 * 4096;        // bias | bias2 (2 x 256 B)
constexpr int RSTAGE_BYTES = RBIAS_OFF + 512;        // 78,336 B per ring slot

struct RSlice {
  const bf16* W; size_t ldw; int r0, r1, N;
  const float *bias, *bias2, *lg, *lb, *lf;
};

// thread 0: arm the slot's mbarrier and launch the bulk copies of slice `s`
__device__ __forceinline__ void ring_issue(uint8_t* slot, uint64_t* bar, const RSlice& s) {
  using namespace tc;
  const int rows = s.r1 - s.r0;
  const int a0 = s.r0 & ~3;
  int nb = ((s.r1 - a0) + 3) & ~3;
  if (a0 + nb > s.N) nb = s.N - a0;
  uint32_t bytes = (uint32_t)rows * 2048u;
  if (s.lg) bytes += 8192u;
  if (s.lf) bytes += 4096u;
  if (s.bias) bytes += (uint32_t)nb * 4u;
  if (s.bias2) bytes += (uint32_t)nb * 4u;
  fence_proxy_async();  // the slot was last read through the generic proxy
  mbar_expect_tx(bar, bytes);
  if (s.ldw == (size_t)kD) {
    bulk_g2s(slot, s.W + (size_t)s.r0 * kD, (uint32_t)rows * 2048u, bar);
  } else {
    for (int r = 0; r < rows; ++r) bulk_g2s(slot + r * 2048, s.W + (size_t)(s.r0 + r) * s.ldw, 2048u, bar);
  }
  if (s.lg) { bulk_g2s(slot + RLN_OFF, s.lg, 4096u, bar); bulk_g2s(slot + RLN_OFF + 4096, s.lb, 4096u, bar); }
  if (s.lf) bulk_g2s(slot + RLN_OFF + 8192, s.lf, 4096u, bar);
  if (s.bias) bulk_g2s(slot + RBIAS_OFF, s.bias + a0, (uint32_t)nb * 4u, bar);
  if (s.bias2) bulk_g2s(slot + RBIAS_OFF + 256, s.bias2 + a0, (uint32_t)nb * 4u, bar);
}

template <int MR>
__device__ __forceinline__ void ring_row_dot(const uint8_t* wrow, const float* __restrict__ sx, float (&acc)[MR]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint4 r = *reinterpret_cast<const uint4*>(wrow + it * 512 + lane * 16);
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
    float w[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[2 * i] = __uint_as_float(w4[i] << 16); w[2 * i + 1] = __uint_as_float(w4[i] & 0xffff0000u); }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
      float a = acc[m], c = 0.f;  // two chains halve the FMA dependency depth
      a = fmaf(w[0], x0.x, a); c = fmaf(w[1], x0.y, c); a = fmaf(w[2], x0.z, a); c = fmaf(w[3], x0.w, c);
      a = fmaf(w[4], x1.x, a); c = fmaf(w[5], x1.y, c); a = fmaf(w[6], x1.z, a); c = fmaf(w[7], x1.w, c);
      acc[m] = a + c;
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = warp_sum(acc[m]);
}

// rows of the slice from the ring slot; addend(n, m) evaluated by lane m before the dots; emit(n, m, v) by lane m.
template <int MR, typename Add, typename Emit>
__device__ __forceinline__ void ring_finish(const uint8_t* slot, const RSlice& s, const float* __restrict__ sx, int B,
                                            Add addend, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = s.r0 & ~3;
  const float* sb = reinterpret_cast<const float*>(slot + RBIAS_OFF);
  const float* sb2 = reinterpret_cast<const float*>(slot + RBIAS_OFF + 256);
  for (int n = s.r0 + warp; n < s.r1; n += PNW) {
    float ad = (lane < B) ? addend(n, lane) : 0.f;
    if (s.bias) ad += sb[n - a0];
    if (s.bias2) ad += sb2[n - a0];
    float acc[MR];
    ring_row_dot<MR>(slot + (size_t)(n - s.r0) * 2048, sx, acc);
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m)
      if (lane == m) mine = acc[m];
    if (lane < B) emit(n, lane, mine + ad);
  }
}

// LayerNorm staging with the parameters taken from the ring slot
template <int MR, typename Val>
__device__ __forceinline__ void ring_stage_ln(float* sx, int B, Val value, const uint8_t* slot, bool has_f, float* red) {
  const float* lnp = reinterpret_cast<const float*>(slot + RLN_OFF);
  LnPre lp;
  const int t = threadIdx.x;
  lp.g0 = lnp[t]; lp.g1 = lnp[t + PNT]; lp.b0 = lnp[1024 + t]; lp.b1 = lnp[1024 + t + PNT];
  lp.f0 = has_f ? lnp[2048 + t] : 0.f; lp.f1 = has_f ? lnp[2048 + t + PNT] : 0.f;
  stage_ln<MR>(sx, B, value, lp, red);
}

// out-proj input: softmax-combine the key splits.  (m, l) of every (b, head, split) first go to shared memory, then
// each thread folds its columns with all split loads in flight at once.
template <int MR>
__device__ __forceinline__ void stage_att(float* sx, const PParams& p, int B, float* sml /*[B*16*nsplit*2]*/) {
  const int tid = threadIdx.x;
  const int npairs = B * kH * p.nsplit;
  for (int i = tid; i < npairs * 2; i += PNT) {
    const int pair = i >> 1, which = i & 1;
    sml[i] = __ldcg(p.part + (size_t)pair * (kDH + 2) + which);
  }
  __syncthreads();
  for (int i = tid; i < B * kD; i += PNT) {
    const int b = i / kD, c = i % kD, h = c >> 6, d = c & 63;
    const float* ml = sml + (size_t)(b * kH + h) * p.nsplit * 2;
    const float* po = p.part + ((size_t)(b * kH + h) * p.nsplit) * (kDH + 2) + 2 + d;
    float o[16];
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) o[s2] = (s2 < p.nsplit) ? __ldcg(po + (size_t)s2 * (kDH + 2)) : 0.f;
    float M = -INFINITY;
    for (int s2 = 0; s2 < p.nsplit; ++s2) M = fmaxf(M, ml[2 * s2]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      if (s2 < p.nsplit) {
        const float ms = ml[2 * s2];
        const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
        L += w * ml[2 * s2 + 1];
        O += w * o[s2];
      }
    }
    for (int s2 = 16; s2 < p.nsplit; ++s2) {  // nsplit > 16 only for very long contexts
      const float ms = ml[2 * s2];
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
      L += w * ml[2 * s2 + 1];
      O += w * __ldcg(po + (size_t)s2 * (kDH + 2));
    }
    sx[i] = ld_cg(p.bt.finished + b) ? 0.f : O / L;
  }
  __syncthreads();
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_ring_kernel(const PParams p) {
  using namespace tc;
  extern __shared__ __align__(128) uint8_t rsm[];
  uint8_t* ring = rsm;                                                   // [2][RSTAGE_BYTES]
  float* sx = reinterpret_cast<float*>(rsm + 2 * RSTAGE_BYTES);          // [MR][1024]
  float* scratch = sx + MR * kD;                                         // attention / sampling / LN scratch
  uint64_t* rbar = reinterpret_cast<uint64_t*>(rsm + 2 * RSTAGE_BYTES + (size_t)MR * kD * 4 + 32768);
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int per = gridDim.x / 4;
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  // GEMV phase c of the 49-phase cycle of one step: 4 per layer (qkv, out, ff1, ff2) + the predict layer
  auto slice_of = [&](int c) {
    RSlice s{};
    if (c == 4 * kL) {
      s.W = reinterpret_cast<const bf16*>(p.predict_w); s.ldw = kD; s.r0 = pr0; s.r1 = pr1; s.N = kARLogits;
      s.lg = p.norm_g; s.lb = p.norm_b; s.lf = p.layer[kL - 1].ff2_b;
      return s;
    }
    const int l = c >> 2, kind = c & 3;
    const PLayer& L = p.layer[l];
    const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
    if (kind == 0) { s.W = (const bf16*)L.qkv_w; s.ldw = kD; s.r0 = q0; s.r1 = q1; s.N = 3 * kD; s.bias = L.qkv_b; s.lg = L.n1_g; s.lb = L.n1_b; s.lf = fb_prev; }
    else if (kind == 1) { s.W = (const bf16*)L.out_w; s.ldw = kD; s.r0 = o0; s.r1 = o1; s.N = kD; s.bias = L.out_b; s.bias2 = fb_prev; }
    else if (kind == 2) { s.W = (const bf16*)L.ff1_w; s.ldw = kD; s.r0 = f0; s.r1 = f1; s.N = kFF; s.bias = L.ff1_b; s.lg = L.n2_g; s.lb = L.n2_b; }
    else { s.W = (const bf16*)L.ff2_w + (size_t)kc * kD; s.ldw = kFF; s.r0 = g0; s.r1 = g1; s.N = kD; }
    return s;
  };
  constexpr int CYC = 4 * kL + 1;
  if (tid == 0) {
    mbar_init(&rbar[0], 1); mbar_init(&rbar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  int g = 0;  // GEMV phases consumed so far (phase g lives in ring slot g & 1, its mbarrier parity is (g >> 1) & 1)
  if (tid == 0) {
    ring_issue(ring, &rbar[0], slice_of(0));
    ring_issue(ring + RSTAGE_BYTES, &rbar[1], slice_of(1));
  }
  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  // finish GEMV phase g: after every thread has left the slot, thread 0 refills it with phase g + 2, then the grid barrier
  auto next_phase = [&]() {
    __syncthreads();
    if (tid == 0) ring_issue(ring + (g & 1) * RSTAGE_BYTES, &rbar[g & 1], slice_of((g + 2) % CYC));
    ++g;
    grid_barrier(p.bar, epoch, prof);
  };

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0 ----
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      // ---- P1: LN1 + QKV ----
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        if (tid == 0) {  // K/V ranges of the coming attention units -> L2
          const int split = u % p.nsplit, bh = u / p.nsplit, b = bh / kH, h = bh % kH;
          const int n = bt.S[b] + bt.P[b] + ld_cg(bt.T + b) + 1;
          int chunk = (n + p.nsplit - 1) / p.nsplit;
          chunk = (chunk + 31) & ~31;
          const int j0 = split * chunk, j1 = min(n - 1, j0 + chunk);
          if (j1 > j0) {
            const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * 2;
            l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * 2);
            l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * 2);
          }
        }
      }
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, slot, s.lf != nullptr, scratch);
        PSTAMP(3);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
      // ---- P2: attention ----
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        attn_unit<bf16>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: combine splits + out-proj + residual ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        stage_att<MR>(sx, p, B, scratch);
        PSTAMP(3);
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_finish<MR>(slot, s, sx, B, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                        [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
      // ---- P4: LN2 + FFN1 + ReLU ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_stage_ln<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, slot, false, scratch);
        PSTAMP(3);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
        PSTAMP(4);
        next_phase();
      }
      // ---- P5: FFN2 split-K ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        stage_plain(sx, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
        PSTAMP(3);
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
    }
    // ---- PF: final LN + predict ----
    {
      const RSlice s = slice_of(g % CYC);
      const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
      mbar_wait(&rbar[g & 1], (g >> 1) & 1);
      ring_stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, slot, true, scratch);
      ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                      [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
      next_phase();
    }
  }
  // drain the two bulk-copy groups still in flight before the CTA (and its shared memory) goes away
  mbar_wait(&rbar[g & 1], (g >> 1) & 1);
  mbar_wait(&rbar[(g + 1) & 1], ((g + 1) >> 1) & 1);
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

// =================================================================================================================
// Fused variant (bf16 weights, B <= 8, VX_DECODE_FUSED=1): 3 grid barriers per layer instead of 5 (38 per step).
//   P1  LN1 + QKV rows                                                      | grid barrier
//   P2  attention unit (b, head, key split); the splits of a head meet on a ticket (a 9-CTA rendezvous, not a
//       grid barrier), every one folds the partials and multiplies the head's 64 outputs by ITS rows of the
//       out-proj head slice, adding into a fixed-point accumulator          | grid barrier
//   P3  residual + LN2 + FFN1 rows (hidden units [f0,f1)) + ReLU, then at once those hidden units' share of FFN2
//       (rows of the TRANSPOSED W2) added into a fixed-point accumulator    | grid barrier
// Cross-CTA sums go through 64-bit integer atomics on 2^-30 fixed point: integer addition is associative, so the
// result is independent of arrival order (deterministic run to run) and carries more precision than an fp32 sum.
// The consumer of an accumulator adds it to the residual stream while staging; accumulators are double-buffered by
// layer parity and re-zeroed one phase after their consumption (always a barrier before their next use).
// =================================================================================================================
constexpr float kFixScale = 1073741824.f;  // 2^30
constexpr int kHS = 64;                    // max hidden units per CTA (grid >= 64)
__device__ __forceinline__ void fix_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * kFixScale));
}
__device__ __forceinline__ float fix_get(const long long* p) { return __ll2float_rn(__ldcg(p)) * (1.0f / kFixScale); }

// stage_ln + write-back: the staged residual value (before normalisation) of columns [c0,c1) goes to xout
template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_x(float* sx, int B, Val value, const LnPre& lp, float* red, float* xout, int c0, int c1) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = MR < 8 ? MR : 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v0[R], v1[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      v0[m] = (m < nb) ? value(mb + m, tid) + lp.f0 : 0.f;
      v1[m] = (m < nb) ? value(mb + m, tid + PNT) + lp.f1 : 0.f;
    }
    if (xout) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m < nb) {
          if (tid >= c0 && tid < c1) xout[(size_t)(mb + m) * kD + tid] = v0[m];
          if (tid + PNT >= c0 && tid + PNT < c1) xout[(size_t)(mb + m) * kD + tid + PNT] = v1[m];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float s = warp_sum(v0[m] + v1[m]);
      if (lane == 0) red[warp * 8 + m] = s;
    }
    __syncthreads();
    float mean[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PNW; ++w) s += red[w * 8 + m];
      mean[m] = s / (float)kD;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float d0 = v0[m] - mean[m], d1 = v1[m] - mean[m];
      const float s = warp_sum(d0 * d0 + d1 * d1);
      if (lane == 0) red[PNW * 8 + warp * 8 + m] = s;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) s += red[PNW * 8 + w * 8 + m];
        const float rstd = 1.0f / sqrtf(s / (float)kD + 1e-5f);
        sx[(mb + m) * kD + tid] = __fadd_rn(__fmul_rn((v0[m] - mean[m]) * rstd, lp.g0), lp.b0);
        sx[(mb + m) * kD + tid + PNT] = __fadd_rn(__fmul_rn((v1[m] - mean[m]) * rstd, lp.g1), lp.b1);
      }
    }
    __syncthreads();
  }
}

__device__ __forceinline__ void zero_slice(long long* acc, int B, int c0, int c1) {
  const int w = c1 - c0;
  for (int i = threadIdx.x; i < B * w; i += PNT) acc[(size_t)(i / w) * kD + c0 + (i % w)] = 0;
}

struct WoPre { uint4 r[2]; };  // this thread's 16 bytes of the first two 64-row passes over the unit's out-proj rows
// rows [r0,r1) x columns [h*64, h*64+64) of Wo: 4 rows per warp (8 lanes x 8 elements each), 64 rows per pass
__device__ __forceinline__ void wo_issue(WoPre& wp, const bf16* __restrict__ Wo, int h, int r0, int r1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int n = r0 + pass * 64 + warp * 4 + (lane >> 3);
    wp.r[pass] = (n < r1) ? ldw_raw(Wo + (size_t)n * kD + h * kDH + (lane & 7) * 8) : make_uint4(0, 0, 0, 0);
  }
}

// attention unit + head rendezvous + out-proj share.  `att_epoch`: tickets this unit's head has seen so far.
__device__ void attn_unit_fused(const PParams& p, int layer, int b, int h, int split, float* smem, int& att_epoch,
                                const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();
  if (tid < kDH) {
    sq[tid] = __ldcg(row + tid) * 0.125f;
    sknew[tid] = __ldcg(row + kD + tid);
    svnew[tid] = __ldcg(row + 2 * kD + tid);
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<bf16>(sknew[d]), s);
    } else {
      const bf16* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  __syncthreads();
  // ---- rendezvous of the head's key splits (all co-resident: one unit per CTA) ---------------------------------
  att_epoch += p.nsplit;
  if (tid == 0) {
    int32_t* tk = p.tickets + b * kH + h;
    __threadfence();
    atomicAdd(tk, 1);
    if (p.nsplit > 1) {
      unsigned int spins = 0;
      int v;
      do {
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(tk) : "memory");
        if (++spins > (1u << 24)) __trap();
      } while (v - att_epoch < 0);
    }
  }
  __syncthreads();
  // ---- fold the splits (fixed order) -> normalised head output in shared memory ----------------------------------
  if (tid < kDH) sq[tid] = att_combined(p, b, h * kDH + tid);
  __syncthreads();
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns --------------------------------------------------------
  const int grp = lane >> 3, sub = lane & 7;
  float xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
  for (int pass = 0; pass * 64 < r1 - r0; ++pass) {
    const int nrow = r0 + pass * 64 + warp * 4 + grp;
    const bool ok = nrow < r1;
    uint4 raw;
    if (pass < 2) raw = wp.r[pass];
    else raw = ok ? ldw_raw(Wo + (size_t)nrow * kD + h * kDH + sub * 8) : make_uint4(0, 0, 0, 0);
    float w[8];
    raw_to_f32<bf16>(&raw, 0, w);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && sub == 0) fix_add(accO + (size_t)b * kD + nrow, s);
  }
}

// hidden units [f0,f1) of this CTA (values in hs[m*kHS + j-f0]) times their rows of W2^T [4096][1024] -> accF[m][0..1023]
template <int MR>
__device__ __forceinline__ void ffn2_share(const bf16* __restrict__ W2T, int f0, int f1, const float* __restrict__ hs, int B,
                                           long long* accF) {
  const int tid = threadIdx.x;
  float acc[MR][2];
#pragma unroll
  for (int m = 0; m < MR; ++m) { acc[m][0] = 0.f; acc[m][1] = 0.f; }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + tid;  // columns 2*tid, 2*tid+1
  for (int j0 = 0; j0 < f1 - f0; j0 += 16) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      w[i] = 0;
      if (j0 + i < f1 - f0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[i]) : "l"(base + (size_t)(j0 + i) * (kD / 2)));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float wa = __uint_as_float(w[i] << 16), wb = __uint_as_float(w[i] & 0xffff0000u);
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float hv = hs[m * kHS + j0 + i];   // zero beyond f1-f0 (cleared by the caller)
        acc[m][0] = fmaf(wa, hv, acc[m][0]);
        acc[m][1] = fmaf(wb, hv, acc[m][1]);
      }
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    if (m < B) {
      fix_add(accF + (size_t)m * kD + 2 * tid, acc[m][0]);
      fix_add(accF + (size_t)m * kD + 2 * tid + 1, acc[m][1]);
    }
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_fused_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* hs = dsm + MR * kD;             // [MR][kHS] FFN hidden units of this CTA
  float* scratch = hs + MR * kHS;        // attention / sampling / LN-reduction scratch
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);            // residual-stream columns this CTA writes back / accumulator columns it re-zeroes
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // this CTA's attention unit, fixed for the whole launch
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;
  int att_epoch = 0;
  long long* const accO2 = p.acc_out;   // [2][B][1024]
  long long* const accF2 = p.acc_ffn;   // [2][B][1024]
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed (x_0 -> xa) ---------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      // ---- P1: x_l = xmid_{l-1} + FFN2_{l-1} (+bias); LN1; QKV rows --------------------------------------------------------
      if (has_unit && tid == 0) {
        const int n = bt.S[u_b] + bt.P[u_b] + ld_cg(bt.T + u_b) + 1;
        int chunk = (n + ns - 1) / ns;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)u_b * kH + u_h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
          l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
          l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
        zero_slice(accO_prev, B, o0, o1);   // consumed by P3 of layer l-1
      }
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows ride through the attention phase
      if (has_unit) wo_issue(wp, Wo, u_h, wr0, wr1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention unit + head rendezvous + out-proj share ------------------------------------------------------
      if (has_unit && !ld_cg(bt.finished + u_b))
        attn_unit_fused(p, l, u_b, u_h, u_split, scratch, att_epoch, wp, Wo, wr0, wr1, accO);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by P1 of this layer
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: xmid_l = x_l + out-proj (+bias); LN2; FFN1 rows + ReLU; FFN2 share --------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1, p.layer[l + 1].qkv_b, nullptr);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: x_12 = xmid_11 + FFN2_11 (+bias); final LayerNorm; predict rows -> logits -----------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by P3 of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t fused_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;             // attention
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;     // sampling
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_fused_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_fused_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_fused_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

// =================================================================================================================
// Trimmed 3-barrier variant (VX_DECODE_TRIM=<bits>, experimental A/B switches on top of ar_decode_fused_kernel):
//   1  V chunks of the attention unit requested together with the K rows (one memory round trip less)
//   2  the CTA's W2^T share prefetched into registers at the start of the FFN phase
//   4  attention partials exchanged through flag-carrying stores instead of fence + ticket + spin + reload
// =================================================================================================================
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, uint32_t epoch);
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p);
__device__ void attn_unit_fused2(const PParams& p, int layer, int b, int h, int split, float* smem, int& att_epoch, uint32_t ll_epoch,
                                const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();
  if (tid < kDH) {
    sq[tid] = __ldcg(row + tid) * 0.125f;
    sknew[tid] = __ldcg(row + kD + tid);
    svnew[tid] = __ldcg(row + 2 * kD + tid);
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  // trim bit 1: this thread's V chunks (keys j0+g, +64, ...) are requested now, so they travel with the K rows
  constexpr int VP = 4;
  uint4 vraw[VP];
  {
    const int g_ = tid >> 3, dc_ = (tid & 7) * 8;
#pragma unroll
    for (int t = 0; t < VP; ++t) {
      const int j = j0 + g_ + t * (PNT / 8);
      vraw[t] = ((p.trim & 1) && j < j1 && j != pos) ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + dc_)) : make_uint4(0, 0, 0, 0);
    }
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<bf16>(sknew[d]), s);
    } else {
      const bf16* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int vt = 0;
  for (int j = j0 + g; j < j1; j += NG, ++vt) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else if ((p.trim & 1) && vt < VP) {
      uint4 r = vraw[0];
#pragma unroll
      for (int t = 1; t < VP; ++t)
        if (vt == t) r = vraw[t];
      raw_to_f32<bf16>(&r, 0, w);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
  if (p.trim & 4) {
    // trim bit 4: partials carry their own flag (value + epoch in one 8-byte store); the fold polls the data itself
    unsigned long long* pg = p.part_ll + ((size_t)b * kH + h) * p.nsplit * (kDH + 2);
    if (tid < kDH) ll_store(pg + (size_t)split * (kDH + 2) + 2 + tid, od, ll_epoch);
    if (tid == 0) {
      ll_store(pg + (size_t)split * (kDH + 2), m, ll_epoch);
      ll_store(pg + (size_t)split * (kDH + 2) + 1, l, ll_epoch);
    }
    if (tid < kDH) {
      constexpr int U = 12;
      const int ns = p.nsplit;
      float mv[U], lv[U], ov[U];
      bool ok;
      unsigned int spins = 0;
      do {
        ok = true;
#pragma unroll
        for (int s = 0; s < U; ++s) {
          if (s < ns) {
            const unsigned long long a = ll_load(pg + (size_t)s * (kDH + 2));
            const unsigned long long c = ll_load(pg + (size_t)s * (kDH + 2) + 1);
            const unsigned long long d = ll_load(pg + (size_t)s * (kDH + 2) + 2 + tid);
            ok = ok && (uint32_t)(a >> 32) == ll_epoch && (uint32_t)(c >> 32) == ll_epoch && (uint32_t)(d >> 32) == ll_epoch;
            mv[s] = __uint_as_float((uint32_t)a); lv[s] = __uint_as_float((uint32_t)c); ov[s] = __uint_as_float((uint32_t)d);
          } else {
            mv[s] = -INFINITY; lv[s] = 0.f; ov[s] = 0.f;
          }
        }
        if (++spins > (1u << 22)) __trap();
      } while (!ok);
      float M = -INFINITY;
#pragma unroll
      for (int s = 0; s < U; ++s) M = fmaxf(M, mv[s]);
      float L = 0.f, O = 0.f;
#pragma unroll
      for (int s = 0; s < U; ++s) {
        const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
        L += w * lv[s];
        O += w * ov[s];
      }
      sq[tid] = O / L;
    }
    __syncthreads();
  } else {
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  __syncthreads();
  // ---- rendezvous of the head's key splits (all co-resident: one unit per CTA) ---------------------------------
  att_epoch += p.nsplit;
  if (tid == 0) {
    int32_t* tk = p.tickets + b * kH + h;
    __threadfence();
    atomicAdd(tk, 1);
    if (p.nsplit > 1) {
      unsigned int spins = 0;
      int v;
      do {
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(tk) : "memory");
        if (++spins > (1u << 24)) __trap();
      } while (v - att_epoch < 0);
    }
  }
  __syncthreads();
  // ---- fold the splits (fixed order) -> normalised head output in shared memory ----------------------------------
  if (tid < kDH) sq[tid] = att_combined(p, b, h * kDH + tid);
  __syncthreads();
  }
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns --------------------------------------------------------
  const int grp = lane >> 3, sub = lane & 7;
  float xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
  for (int pass = 0; pass * 64 < r1 - r0; ++pass) {
    const int nrow = r0 + pass * 64 + warp * 4 + grp;
    const bool ok = nrow < r1;
    uint4 raw;
    if (pass < 2) raw = wp.r[pass];
    else raw = ok ? ldw_raw(Wo + (size_t)nrow * kD + h * kDH + sub * 8) : make_uint4(0, 0, 0, 0);
    float w[8];
    raw_to_f32<bf16>(&raw, 0, w);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && sub == 0) fix_add(accO + (size_t)b * kD + nrow, s);
  }
}

struct W2Pre { uint32_t w[32]; };   // columns (2*tid, 2*tid+1) of the first 32 hidden units' rows of W2^T
__device__ __forceinline__ void ffn2_issue(W2Pre& wp2, const bf16* __restrict__ W2T, int f0, int f1) {
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    wp2.w[i] = 0;
    if (i < f1 - f0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(wp2.w[i]) : "l"(base + (size_t)i * (kD / 2)));
  }
}
template <int MR>
__device__ __forceinline__ void ffn2_share_pre(const W2Pre& wp2, const bf16* __restrict__ W2T, int f0, int f1, const float* __restrict__ hs,
                                               int B, long long* accF) {
  const int tid = threadIdx.x;
  float acc[MR][2];
#pragma unroll
  for (int m = 0; m < MR; ++m) { acc[m][0] = 0.f; acc[m][1] = 0.f; }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float wa = __uint_as_float(wp2.w[i] << 16), wb = __uint_as_float(wp2.w[i] & 0xffff0000u);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float hv = hs[m * kHS + i];   // zero beyond f1-f0
      acc[m][0] = fmaf(wa, hv, acc[m][0]);
      acc[m][1] = fmaf(wb, hv, acc[m][1]);
    }
  }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + tid;
  for (int j = 32; j < f1 - f0; ++j) {   // grids smaller than 128 CTAs only
    uint32_t w;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w) : "l"(base + (size_t)j * (kD / 2)));
    const float wa = __uint_as_float(w << 16), wb = __uint_as_float(w & 0xffff0000u);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float hv = hs[m * kHS + j];
      acc[m][0] = fmaf(wa, hv, acc[m][0]);
      acc[m][1] = fmaf(wb, hv, acc[m][1]);
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    if (m < B) {
      fix_add(accF + (size_t)m * kD + 2 * tid, acc[m][0]);
      fix_add(accF + (size_t)m * kD + 2 * tid + 1, acc[m][1]);
    }
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_fused2_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* hs = dsm + MR * kD;             // [MR][kHS] FFN hidden units of this CTA
  float* scratch = hs + MR * kHS;        // attention / sampling / LN-reduction scratch
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);            // residual-stream columns this CTA writes back / accumulator columns it re-zeroes
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // this CTA's attention unit, fixed for the whole launch
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;
  int att_epoch = 0;
  long long* const accO2 = p.acc_out;   // [2][B][1024]
  long long* const accF2 = p.acc_ffn;   // [2][B][1024]
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  W2Pre wp2;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed (x_0 -> xa) ---------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      // ---- P1: x_l = xmid_{l-1} + FFN2_{l-1} (+bias); LN1; QKV rows --------------------------------------------------------
      if (has_unit && tid == 0) {
        const int n = bt.S[u_b] + bt.P[u_b] + ld_cg(bt.T + u_b) + 1;
        int chunk = (n + ns - 1) / ns;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)u_b * kH + u_h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
          l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
          l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
        zero_slice(accO_prev, B, o0, o1);   // consumed by P3 of layer l-1
      }
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows ride through the attention phase
      if (has_unit) wo_issue(wp, Wo, u_h, wr0, wr1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention unit + head rendezvous + out-proj share ------------------------------------------------------
      if (has_unit && !ld_cg(bt.finished + u_b))
        attn_unit_fused2(p, l, u_b, u_h, u_split, scratch, att_epoch, (uint32_t)(step * kL + l + 1), wp, Wo, wr0, wr1, accO);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by P1 of this layer
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: xmid_l = x_l + out-proj (+bias); LN2; FFN1 rows + ReLU; FFN2 share --------------------------------------
      if (p.trim & 2) ffn2_issue(wp2, W2T, f0, f1);   // trim bit 2: the W2^T share rides under staging + LN2 + FFN1
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      if (p.trim & 2) ffn2_share_pre<MR>(wp2, W2T, f0, f1, hs, B, accF);
      else ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1, p.layer[l + 1].qkv_b, nullptr);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: x_12 = xmid_11 + FFN2_11 (+bias); final LayerNorm; predict rows -> logits -----------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by P3 of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t fused2_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_fused2_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_fused2_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_fused2_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

// =================================================================================================================
// Two-barrier variant (bf16 weights, VX_DECODE_V4=1, experimental): 2 grid barriers per layer (26 per step).
//   PA  stage x_l, LN1, the unit's share of ITS HEAD's q/k/v rows (192 rows of Wqkv split over the head's key-split
//       CTAs) -> exchanged inside the head group through flag-carrying 8-byte stores (value + epoch in one store,
//       readers poll the data itself: one L2 round trip, no fence, no ticket) -> attention on the unit's key split ->
//       partials exchanged the same way -> fold -> out-proj share -> fixed-point accumulator      | grid barrier
//   PB  residual + LN2 + FFN1 rows + ReLU + FFN2 share -> fixed-point accumulator                  | grid barrier
// Only the two cross-head sums (out-proj, FFN2) need the whole grid; everything else is local to the <= 9 CTAs of a head.
// =================================================================================================================
__device__ __forceinline__ void ll_store(unsigned long long* p, float v, uint32_t epoch) {
  const unsigned long long w = ((unsigned long long)epoch << 32) | (unsigned long long)__float_as_uint(v);
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long ll_load(const unsigned long long* p) {
  unsigned long long w;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
  return w;
}
// poll one element until its flag carries `epoch`
__device__ __forceinline__ float ll_wait(const unsigned long long* p, uint32_t epoch) {
  unsigned long long w;
  unsigned int spins = 0;
  do {
    w = ll_load(p);
    if (++spins > (1u << 24)) __trap();
  } while ((uint32_t)(w >> 32) != epoch);
  return __uint_as_float((uint32_t)w);
}

// row i (0..191) of head h's q|k|v block -> row of Wqkv [3072][1024]
__device__ __forceinline__ int qkv_head_row(int h, int i) { return (i >> 6) * kD + h * kDH + (i & 63); }

// attention unit of the two-barrier variant: q/k/v arrive through qkv_ll, partials leave through part_ll
__device__ void attn_unit_v4(const PParams& p, int layer, int b, int h, int split, float* smem, uint32_t epoch,
                             const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ns = p.nsplit;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + ns - 1) / ns;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  // ---- q/k/v of this head from the group's flag-carrying stores -------------------------------------------------
  if (tid < 3 * kDH) {
    const float v = ll_wait(p.qkv_ll + ((size_t)b * kH + h) * (3 * kDH) + tid, epoch);
    if (tid < kDH) sq[tid] = v * 0.125f;
    else smem[tid] = v;              // sknew | svnew are contiguous behind sq
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<bf16>(sknew[d]), s);
    } else {
      const bf16* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  // ---- this split's partial (m, l, o[64]) to the group ---------------------------------------------------------------
  unsigned long long* pg = p.part_ll + ((size_t)b * kH + h) * ns * (kDH + 2);
  if (tid < kDH) {
    float od = 0.f;
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
    ll_store(pg + (size_t)split * (kDH + 2) + 2 + tid, od, epoch);
  }
  if (tid == 0) {
    ll_store(pg + (size_t)split * (kDH + 2), m, epoch);
    ll_store(pg + (size_t)split * (kDH + 2) + 1, l, epoch);
  }
  // ---- fold all splits (fixed order); every load of a round is in flight at once, rounds repeat until all flags match --
  if (tid < kDH) {
    constexpr int U = 12;
    float mv[U], lv[U], ov[U];
    bool ok;
    unsigned int spins = 0;
    do {
      ok = true;
#pragma unroll
      for (int s = 0; s < U; ++s) {
        if (s < ns) {
          const unsigned long long a = ll_load(pg + (size_t)s * (kDH + 2));
          const unsigned long long c = ll_load(pg + (size_t)s * (kDH + 2) + 1);
          const unsigned long long d = ll_load(pg + (size_t)s * (kDH + 2) + 2 + tid);
          ok = ok && (uint32_t)(a >> 32) == epoch && (uint32_t)(c >> 32) == epoch && (uint32_t)(d >> 32) == epoch;
          mv[s] = __uint_as_float((uint32_t)a); lv[s] = __uint_as_float((uint32_t)c); ov[s] = __uint_as_float((uint32_t)d);
        } else {
          mv[s] = -INFINITY; lv[s] = 0.f; ov[s] = 0.f;
        }
      }
      if (++spins > (1u << 22)) __trap();
    } while (!ok);
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < U; ++s) M = fmaxf(M, mv[s]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int s = 0; s < U; ++s) {
      const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
      L += w * lv[s];
      O += w * ov[s];
    }
    sq[tid] = O / L;
  }
  __syncthreads();
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns -------------------------------------------------------
  const int grp = lane >> 3, sub = lane & 7;
  float xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
  for (int pass = 0; pass * 64 < r1 - r0; ++pass) {
    const int nrow = r0 + pass * 64 + warp * 4 + grp;
    const bool ok = nrow < r1;
    uint4 raw;
    if (pass < 2) raw = wp.r[pass];
    else raw = ok ? ldw_raw(Wo + (size_t)nrow * kD + h * kDH + sub * 8) : make_uint4(0, 0, 0, 0);
    float w[8];
    raw_to_f32<bf16>(&raw, 0, w);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && sub == 0) fix_add(accO + (size_t)b * kD + nrow, s);
  }
}

// request rows map(i), i = warp + j*16 < cnt, of W [.][1024]
template <int MR, typename Map>
__device__ __forceinline__ void gemv_issue_map(WPre<bf16, MR>& pre, const bf16* __restrict__ W, int cnt, const float* __restrict__ bias, Map map) {
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = warp + j * PNW;
    pre.bias[j] = 0.f;
    if (i < cnt) {
      const int n = map(i);
      load_row_raw<bf16, MR>(W, kD, n, pre.r[j]);
      if (bias) pre.bias[j] = bias[n];
    }
  }
}
// consume them (and any rows beyond 2 per warp); emit(i, m, value) by lane m
template <int MR, typename Map, typename Emit>
__device__ __forceinline__ void gemv_finish_map(WPre<bf16, MR>& pre, const bf16* __restrict__ W, int cnt, const float* __restrict__ sx,
                                                int B, const float* __restrict__ bias, Map map, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = warp, i1 = warp + PNW;
  if (i0 < cnt) {  // warp-uniform
    float a0[MR], a1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      float w0[8], w1[8];
      raw_to_f32<bf16>(pre.r[0], it, w0);
      raw_to_f32<bf16>(pre.r[1], it, w1);
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[m] = fmaf(w0[i], xv[i], a0[m]); a1[m] = fmaf(w1[i], xv[i], a1[m]); }
      }
    }
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float s0 = warp_sum(a0[m]), s1 = warp_sum(a1[m]);
      if (lane == m) { m0 = s0; m1 = s1; }
    }
    if (lane < B) {
      emit(i0, lane, m0 + pre.bias[0]);
      if (i1 < cnt) emit(i1, lane, m1 + pre.bias[1]);
    }
  }
  for (int i = warp + 2 * PNW; i < cnt; i += PNW) {   // more than 32 rows per CTA (fewer key splits)
    const int n = map(i);
    uint4 raw[WPre<bf16, MR>::kVec];
    load_row_raw<bf16, MR>(W, kD, n, raw);
    float acc[MR];
    row_dot<bf16, MR>(raw, sx, acc);
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m)
      if (lane == m) mine = acc[m];
    if (lane < B) emit(i, lane, mine + (bias ? bias[n] : 0.f));
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_v4_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* hs = dsm + MR * kD;             // [MR][kHS]
  float* scratch = hs + MR * kHS;
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int o0, o1, f0, f1, pr0, pr1;
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;            // out-proj rows of this unit
  const int qi0 = (u_split * 3 * kDH) / ns, qi1 = ((u_split + 1) * 3 * kDH) / ns;   // its share of the head's q|k|v rows
  const int qcnt = has_unit ? qi1 - qi0 : 0;
  auto qmap = [&](int i) { return qkv_head_row(u_h, qi0 + i); };
  long long* const accO2 = p.acc_out;
  long long* const accF2 = p.acc_ffn;
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), qcnt, p.layer[0].qkv_b, qmap);
  if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed (x_0 -> xa) ---------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      const uint32_t ll_epoch = (uint32_t)(step * kL + l + 1);
      const bool unit_live = has_unit && !ld_cg(bt.finished + u_b);
      // ---- PA ------------------------------------------------------------------------------------------------------
      if (has_unit && tid == 0) {   // pull this unit's K/V range towards L2 while the q/k/v rows are computed
        const int n = bt.S[u_b] + bt.P[u_b] + ld_cg(bt.T + u_b) + 1;
        int chunk = (n + ns - 1) / ns;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)u_b * kH + u_h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
          l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
          l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
      }
      zero_slice(accO_prev, B, o0, o1);   // consumed by PB of the previous layer (layer 0: by PB of layer 11, re-zeroed twice: harmless)
      PSTAMP(3);
      // only this utterance's row of sx matters to the unit; all B rows are computed and row u_b is emitted
      gemv_finish_map<MR>(pre, Wq, qcnt, sx, B, L.qkv_b, qmap, [&](int i, int m, float v) {
        if (m == u_b && unit_live) ll_store(p.qkv_ll + ((size_t)u_b * kH + u_h) * (3 * kDH) + qi0 + i, v, ll_epoch);
      });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows ride through the attention part
      if (unit_live) attn_unit_v4(p, l, u_b, u_h, u_split, scratch, ll_epoch, wp, Wo, wr0, wr1, accO);
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- PB ------------------------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by PA of this layer
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), qcnt, p.layer[l + 1].qkv_b, qmap);
        if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[l + 1].out_w), u_h, wr0, wr1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF --------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by PB of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), qcnt, p.layer[0].qkv_b, qmap);
    if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

// =================================================================================================================
// Cluster variant (bf16 weights, B = 1, VX_DECODE_CLUSTER=1, experimental): the two-barrier layout of the v4 kernel with
// the head group mapped onto a thread-block CLUSTER: 16 clusters x 8 CTAs = 128 CTAs (8 GPCs x two 8-CTA clusters is what
// fits co-resident on the 148 SMs), cluster c = head c, rank = key split.  q/k/v of the head and the attention partials
// are written to the producer's own shared memory and read by the other ranks through distributed shared memory behind
// barrier.cluster (~0.2 us) instead of through L2 flags (~1.5-2 us per exchange).  Grid barriers as before.
// =================================================================================================================
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the same shared-memory offset in cluster rank `rank`
__device__ __forceinline__ float dsmem_ld(const float* local, uint32_t rank) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

constexpr int kCS = 8;                 // cluster size = key splits per head
constexpr int kQShare = 3 * kDH / kCS; // 24 of the head's 192 q|k|v rows per CTA

// xchg: [0,24) this CTA's q|k|v values, [32, 32+66) its attention partial (m, l, o[64]); read by the other ranks via DSMEM
__device__ void attn_unit_cluster(const PParams& p, int layer, int h, int split, float* smem, float* xchg,
                                  const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int b = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + kCS - 1) / kCS;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  // ---- q/k/v of this head: element i lives in rank i / 24 at xchg[i % 24] -----------------------------------------------
  cluster_sync_all();                // every rank's q|k|v share is in its xchg[0,24)
  if (tid < 3 * kDH) {
    const float v = dsmem_ld(xchg + (tid % kQShare), (uint32_t)(tid / kQShare));
    if (tid < kDH) sq[tid] = v * 0.125f;
    else smem[tid] = v;              // sknew | svnew are contiguous behind sq
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<bf16>(sknew[d]), s);
    } else {
      const bf16* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  // ---- this split's partial into its own xchg[32..98) ------------------------------------------------------------------
  float* mypart = xchg + 32;
  if (tid < kDH) {
    float od = 0.f;
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
    mypart[2 + tid] = od;
  }
  if (tid == 0) { mypart[0] = m; mypart[1] = l; }
  cluster_sync_all();                // every rank's partial is visible cluster-wide
  // ---- fold the 8 splits (fixed order) ---------------------------------------------------------------------------------------
  if (tid < kDH) {
    float mv[kCS], lv[kCS], ov[kCS];
#pragma unroll
    for (int s = 0; s < kCS; ++s) {
      mv[s] = dsmem_ld(mypart, (uint32_t)s);
      lv[s] = dsmem_ld(mypart + 1, (uint32_t)s);
      ov[s] = dsmem_ld(mypart + 2 + tid, (uint32_t)s);
    }
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < kCS; ++s) M = fmaxf(M, mv[s]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int s = 0; s < kCS; ++s) {
      const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
      L += w * lv[s];
      O += w * ov[s];
    }
    sq[tid] = O / L;
  }
  __syncthreads();
  // ---- out-proj share: rows [r0,r1) (128 rows) x this head's 64 columns -------------------------------------------------------
  const int grp = lane >> 3, sub = lane & 7;
  float xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
  for (int pass = 0; pass * 64 < r1 - r0; ++pass) {
    const int nrow = r0 + pass * 64 + warp * 4 + grp;
    const bool ok = nrow < r1;
    uint4 raw;
    if (pass < 2) raw = wp.r[pass];
    else raw = ok ? ldw_raw(Wo + (size_t)nrow * kD + h * kDH + sub * 8) : make_uint4(0, 0, 0, 0);
    float w[8];
    raw_to_f32<bf16>(&raw, 0, w);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && sub == 0) fix_add(accO + (size_t)b * kD + nrow, s);
  }
  // the partial / q|k|v slots are rewritten two grid barriers from now, long after every rank has read them
}

__global__ void __launch_bounds__(PNT, 1) ar_decode_cluster_kernel(const PParams p) {
  using WT = bf16;
  constexpr int MR = 1;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [1024]
  float* hs = dsm + MR * kD;             // [kHS]
  float* xchg = hs + MR * kHS;           // [128] cluster exchange slots (same offset in every CTA of the cluster)
  float* scratch = xchg + 128;
  const Batch& bt = p.bt;
  const int B = 1;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int o0, o1, f0, f1, pr0, pr1;
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int u_h = blockIdx.x / kCS, u_split = (int)cluster_rank();   // cluster c = head c (blocks 8c .. 8c+7), rank = key split
  const int wr0 = (u_split * kD) / kCS, wr1 = ((u_split + 1) * kD) / kCS;
  const int qi0 = u_split * kQShare;
  auto qmap = [&](int i) { return qkv_head_row(u_h, qi0 + i); };
  long long* const accO2 = p.acc_out;
  long long* const accF2 = p.acc_ffn;
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), kQShare, p.layer[0].qkv_b, qmap);
  wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  bool live = true;
  for (int step = 0; step < p.max_steps && live; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    if (blockIdx.x == 0 && !ld_cg(bt.finished)) sample_utterance(p, 0, scratch);
    zero_slice(accF2 + accN, B, o0, o1);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      // ---- PA ------------------------------------------------------------------------------------------------------
      if (tid == 0) {
        const int n = bt.S[0] + bt.P[0] + ld_cg(bt.T) + 1;
        int chunk = (n + kCS - 1) / kCS;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + (size_t)u_h * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
          l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
          l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
      }
      zero_slice(accO_prev, B, o0, o1);
      PSTAMP(3);
      gemv_finish_map<MR>(pre, Wq, kQShare, sx, B, L.qkv_b, qmap, [&](int i, int, float v) { xchg[i] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      attn_unit_cluster(p, l, u_h, u_split, scratch, xchg, wp, Wo, wr0, wr1, accO);
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- PB ------------------------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kQShare, p.layer[l + 1].qkv_b, qmap);
        wo_issue(wp, reinterpret_cast<const WT*>(p.layer[l + 1].out_w), u_h, wr0, wr1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF --------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), kQShare, p.layer[0].qkv_b, qmap);
    wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
  cluster_sync_all();   // no CTA of the cluster leaves while a peer may still read its shared memory
}

static cudaError_t cluster_launch(const PParams& p, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)kD + kHS + 128 + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kH * kCS);
  cfg.blockDim = dim3(PNT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, ar_decode_cluster_kernel, &cfg);
  if (e != cudaSuccess) return e;
  if (nclusters < kH) return cudaErrorLaunchOutOfResources;   // all 16 clusters must be co-resident (grid barrier)
  return cudaLaunchKernelEx(&cfg, ar_decode_cluster_kernel, p);
}

template <int MR>
static cudaError_t v4_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_v4_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_v4_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_v4_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

// =================================================================================================================
// Batched variant (bf16 weights, 8 <= B <= 32, VX_DECODE_MMA=1, experimental): the five GEMV phases of the 62-barrier
// kernel on tensor cores.  Weights are the M operand of mma.sync.m16n8k16 (a CTA's <= 32-row slice = two m-tiles),
// the batch is N, activations are split into bf16 hi + lo halves (two MMAs, fp32 accumulate: fp32-grade activations).
// The k index inside a 32-element chunk is permuted identically for A and B, so every lane loads 16 contiguous bytes
// of a weight row (64-byte runs per row) and its B fragments are 16-byte shared-memory loads from a fragment-ordered
// image written by the staging step.  16 warps = 2 m-tiles x 8 k-groups; the partial tiles are reduced through shared
// memory in fixed order.  (The CUDA-core path re-reads B x 4 KB of activations per weight row and spills at MR >= 8.)
// =================================================================================================================
struct MPre { uint4 ra[4], rb[4]; };   // weight rows (g, g+8) of this warp's m-tile: 4 chunks of 32 k of its k-group

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Fragment image of the activations (hi and lo separately): [k-group 8][chunk 4][n-tile MR/8][lane 32] x 16 B, the 16 B of
// lane (g, q) = batch row nt*8+g, columns kg*128 + ch*32 + q*8 .. +7 in order (words 0,1 = (b0,b1) of the chunk's first
// m16n8k16, words 2,3 = of its second).
template <int MR>
__device__ __forceinline__ void frag_store(uint16_t* fh, uint16_t* fl, int m, int c, float v) {
  constexpr int NT = MR / 8;
  const int kg = c >> 7, ch = (c >> 5) & 3, q = (c >> 3) & 3, e = c & 7;
  const size_t idx = (((((size_t)kg * 4 + ch) * NT + (m >> 3)) * 32) + (m & 7) * 4 + q) * 8 + e;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  fh[idx] = __bfloat16_as_ushort(h);
  fl[idx] = __bfloat16_as_ushort(l);
}

template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_frag(uint16_t* fh, uint16_t* fl, int B, Val value, const LnPre& lp, float* red) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v0[R], v1[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      v0[m] = (m < nb) ? value(mb + m, tid) + lp.f0 : 0.f;
      v1[m] = (m < nb) ? value(mb + m, tid + PNT) + lp.f1 : 0.f;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float s = warp_sum(v0[m] + v1[m]);
      if (lane == 0) red[warp * 8 + m] = s;
    }
    __syncthreads();
    float mean[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PNW; ++w) s += red[w * 8 + m];
      mean[m] = s / (float)kD;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float d0 = v0[m] - mean[m], d1 = v1[m] - mean[m];
      const float s = warp_sum(d0 * d0 + d1 * d1);
      if (lane == 0) red[PNW * 8 + warp * 8 + m] = s;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) s += red[PNW * 8 + w * 8 + m];
        const float rstd = 1.0f / sqrtf(s / (float)kD + 1e-5f);
        frag_store<MR>(fh, fl, mb + m, tid, __fadd_rn(__fmul_rn((v0[m] - mean[m]) * rstd, lp.g0), lp.b0));
        frag_store<MR>(fh, fl, mb + m, tid + PNT, __fadd_rn(__fmul_rn((v1[m] - mean[m]) * rstd, lp.g1), lp.b1));
      }
    }
    __syncthreads();
  }
}
template <int MR, typename Val>
__device__ __forceinline__ void stage_plain_frag(uint16_t* fh, uint16_t* fl, int B, Val value) {
  for (int i = threadIdx.x; i < B * kD; i += PNT) frag_store<MR>(fh, fl, i / kD, i % kD, value(i / kD, i % kD));
  __syncthreads();
}

// request this warp's share of rows [r0,r1) (<= 32 rows) of W [.][ldw], k columns [0,1024)
__device__ __forceinline__ void gemv_issue_mma(MPre& pre, const bf16* __restrict__ W, size_t ldw, int r0, int r1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 1, kg = warp >> 1, g = lane >> 2, q = lane & 3;
  const int ra = r0 + mt * 16 + g, rb = ra + 8;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int k = kg * 128 + ch * 32 + q * 8;
    pre.ra[ch] = (ra < r1) ? ldw_raw(W + (size_t)ra * ldw + k) : make_uint4(0, 0, 0, 0);
    pre.rb[ch] = (rb < r1) ? ldw_raw(W + (size_t)rb * ldw + k) : make_uint4(0, 0, 0, 0);
  }
}

// rows [r0,r1) x batch: tensor-core partial tiles, fixed-order reduction over the 8 k-groups, then
// emit(n, m, dot + addend(n, m) + bias[n] + bias2[n]).  red: [8][32][MR] floats.
template <int MR, typename Add, typename Emit>
__device__ __forceinline__ void gemv_finish_mma(const MPre& pre, const uint16_t* fh, const uint16_t* fl, int r0, int r1, int B,
                                                const float* __restrict__ bias, const float* __restrict__ bias2, float* red,
                                                Add addend, Emit emit) {
  constexpr int NT = MR / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = warp & 1, kg = warp >> 1, g = lane >> 2, q = lane & 3;
  float c[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
  const uint4* H = reinterpret_cast<const uint4*>(fh);
  const uint4* Lo = reinterpret_cast<const uint4*>(fl);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int idx = ((kg * 4 + ch) * NT + nt) * 32 + lane;
      const uint4 xh = H[idx], xl = Lo[idx];
      mma_16816(c[nt], pre.ra[ch].x, pre.rb[ch].x, pre.ra[ch].y, pre.rb[ch].y, xh.x, xh.y);
      mma_16816(c[nt], pre.ra[ch].x, pre.rb[ch].x, pre.ra[ch].y, pre.rb[ch].y, xl.x, xl.y);
      mma_16816(c[nt], pre.ra[ch].z, pre.rb[ch].z, pre.ra[ch].w, pre.rb[ch].w, xh.z, xh.w);
      mma_16816(c[nt], pre.ra[ch].z, pre.rb[ch].z, pre.ra[ch].w, pre.rb[ch].w, xl.z, xl.w);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int row = mt * 16 + g, n = nt * 8 + q * 2;
    float* d0 = red + ((size_t)kg * 32 + row) * MR + n;
    float* d1 = red + ((size_t)kg * 32 + row + 8) * MR + n;
    d0[0] = c[nt][0]; d0[1] = c[nt][1];
    d1[0] = c[nt][2]; d1[1] = c[nt][3];
  }
  __syncthreads();
  for (int t = tid; t < 32 * MR; t += PNT) {
    const int row = t / MR, m = t % MR, n = r0 + row;
    if (n < r1 && m < B) {
      float s = 0.f;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) s += red[((size_t)k8 * 32 + row) * MR + m];
      emit(n, m, s + (addend(n, m) + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f)));
    }
  }
  __syncthreads();
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_mma_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  uint16_t* fh = reinterpret_cast<uint16_t*>(dsm);        // [MR*1024] hi halves, fragment order
  uint16_t* fl = fh + MR * kD;                             // [MR*1024] lo halves
  float* scratch = dsm + MR * kD;                          // attention / sampling / LN scratch
  float* red = scratch;                                    // [8][32][MR] partial tiles (GEMV phases only)
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;
  const size_t wsz = sizeof(WT);

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int per = gridDim.x / 4;
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  MPre pre;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;
  auto zero = [](int, int) { return 0.f; };

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      (void)Wq;
      // ---- P1: LN1 + QKV rows ----------------------------------------------------------------------------------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        if (tid == 0) {
          const int split = u % p.nsplit, bh = u / p.nsplit, b = bh / kH, h = bh % kH;
          const int n = bt.S[b] + bt.P[b] + ld_cg(bt.T + b) + 1;
          int chunk = (n + p.nsplit - 1) / p.nsplit;
          chunk = (chunk + 31) & ~31;
          const int j0 = split * chunk, j1 = min(n - 1, j0 + chunk);
          if (j1 > j0) {
            const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * wsz;
            l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * wsz);
            l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * wsz);
          }
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, lp, scratch);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, q0, q1, B, L.qkv_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue_mma(pre, Wo, kD, o0, o1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention (ticket fold) -----------------------------------------------------------------------------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        attn_unit<bf16>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: out-proj + residual -> xa ---------------------------------------------------------------------------------
      stage_plain_frag<MR>(fh, fl, B, [&](int m, int c) {
        if (ld_cg(bt.finished + m)) return 0.f;
        return (p.dbg & 2) ? att_combined(p, m, c) : __ldcg(p.att + (size_t)m * kD + c);
      });
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, o0, o1, B, L.out_b, fb_prev, red, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                          [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue_mma(pre, W1, kD, f0, f1);
      ln_issue(lp, L.n2_g, L.n2_b, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU -----------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, nullptr);
      stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, f0, f1, B, L.ff1_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue_mma(pre, W2, kFF, g0, g1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2, split-K -----------------------------------------------------------------------------------------------
      stage_plain_frag<MR>(fh, fl, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, g0, g1, B, nullptr, nullptr, red, zero,
                          [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
      PSTAMP(4);
      if (l + 1 < kL) {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF ------------------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, lp, scratch);
    gemv_finish_mma<MR>(pre, fh, fl, pr0, pr1, B, nullptr, nullptr, red, zero,
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t mma_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t red_floats = (size_t)8 * 32 * MR;
  const size_t smem = ((size_t)MR * kD + std::max(std::max(scratch_floats, sample_floats), red_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_mma_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_mma_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_mma_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

template <int MR>
static cudaError_t ring_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t smem = 2 * (size_t)RSTAGE_BYTES + (size_t)MR * kD * 4 + 32768 + 64;
  cudaError_t e = cudaFuncSetAttribute(ar_decode_ring_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_ring_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

template <typename WT, int MR>
static cudaError_t persist_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;             // attention
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;     // sampling
  const size_t smem = ((size_t)MR * kD + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_persistent_kernel<WT, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_persistent_kernel<WT, MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_persistent_kernel<WT, MR>, dim3(grid), dim3(PNT), args, smem, st);
}

// wt: 0 fp32 / 1 bf16.  B <= 32.  Returns a cudaError_t as int.
int launch_decode_persistent(const void* params, int wt, int B, int grid, cudaStream_t st) {
  const PParams& p = *reinterpret_cast<const PParams*>(params);
  const int mr = B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : B <= 8 ? 8 : B <= 16 ? 16 : 32;
  const char* mma_env = getenv("VX_DECODE_MMA");   // tensor-core batched variant (<= 32 rows per CTA slice); VX_DECODE_MMA=0 -> CUDA-core GEMV
  if (wt == 1 && mr >= 8 && grid >= 128 && !(mma_env && mma_env[0] == '0')) {
    switch (mr) {
      case 8: return (int)mma_launch<8>(p, grid, st);
      case 16: return (int)mma_launch<16>(p, grid, st);
      default: return (int)mma_launch<32>(p, grid, st);
    }
  }
  if (wt == 1 && p.fused == 4) {  // experimental trimmed 3-barrier variant (VX_DECODE_TRIM=<bits>)
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kHS || p.nsplit > 12) return (int)cudaErrorInvalidConfiguration;
    switch (mr) {
      case 1: return (int)fused2_launch<1>(p, grid, st);
      case 2: return (int)fused2_launch<2>(p, grid, st);
      case 4: return (int)fused2_launch<4>(p, grid, st);
      default: return (int)fused2_launch<8>(p, grid, st);
    }
  }
  if (wt == 1 && p.fused == 3) {  // experimental cluster variant (VX_DECODE_CLUSTER=1): B = 1, 16 clusters x 8 CTAs
    if (B != 1 || p.nsplit != kCS || (kFF + kH * kCS - 1) / (kH * kCS) > kHS) return (int)cudaErrorInvalidConfiguration;
    return (int)cluster_launch(p, st);
  }
  if (wt == 1 && p.fused == 2) {  // experimental two-barrier variant (VX_DECODE_V4=1): same preconditions as the 3-barrier one
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kHS) return (int)cudaErrorInvalidConfiguration;
    switch (mr) {
      case 1: return (int)v4_launch<1>(p, grid, st);
      case 2: return (int)v4_launch<2>(p, grid, st);
      case 4: return (int)v4_launch<4>(p, grid, st);
      default: return (int)v4_launch<8>(p, grid, st);
    }
  }
  if (wt == 1 && p.fused) {  // 3-barrier variant: requires one attention unit per CTA and <= kHS hidden units per CTA
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kHS) return (int)cudaErrorInvalidConfiguration;
    switch (mr) {
      case 1: return (int)fused_launch<1>(p, grid, st);
      case 2: return (int)fused_launch<2>(p, grid, st);
      case 4: return (int)fused_launch<4>(p, grid, st);
      default: return (int)fused_launch<8>(p, grid, st);
    }
  }
  if (wt == 1 && mr <= 8 && grid >= 128 && getenv("VX_DECODE_RING")) {  // experimental TMA-ring variant (slower: smem-bound dots)  // 28 rows/CTA max fits the 32-row ring slot
    switch (mr) {
      case 1: return (int)ring_launch<1>(p, grid, st);
      case 2: return (int)ring_launch<2>(p, grid, st);
      case 4: return (int)ring_launch<4>(p, grid, st);
      default: return (int)ring_launch<8>(p, grid, st);
    }
  }
#define VX_P(WT_, MR_) return (int)persist_launch<WT_, MR_>(p, grid, st)
  if (wt == 0) {
    switch (mr) { case 1: VX_P(float, 1); case 2: VX_P(float, 2); case 4: VX_P(float, 4); case 8: VX_P(float, 8); case 16: VX_P(float, 16); default: VX_P(float, 32); }
  } else {
    switch (mr) { case 1: VX_P(bf16, 1); case 2: VX_P(bf16, 2); case 4: VX_P(bf16, 4); case 8: VX_P(bf16, 8); case 16: VX_P(bf16, 16); default: VX_P(bf16, 32); }
  }
#undef VX_P
}

}  // namespace vx
