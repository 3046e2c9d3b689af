// Device-side building blocks shared by the persistent decode kernels (decode_persist.cu, decode_fused.cu,
// decode_exchange.cu, decode_mma.cu): grid barrier, weight-streaming GEMV with cross-barrier register prefetch,
// LayerNorm staging, on-device sampling, the attention unit of one (utterance, head, key split) and its fold.
#pragma once
#include <cooperative_groups.h>

#include <algorithm>
#include <cstdlib>

#include "common.cuh"
#include "kernels.h"
#include "decode_persist.h"
#include "tc_common.cuh"

namespace vx {
namespace {   // internal linkage: every translation unit gets its own copy of the non-template device functions

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
    // No acquire fence here, on measurement: `fence.acq_rel.gpu` after the poll (the formally complete acquire side)
    // costs +0.6 us per barrier = +24 us per batch-1 step (260 -> 284 us; it invalidates L1, so the LayerNorm vectors and
    // biases are re-fetched from L2 after every barrier).  What makes the relaxed form safe in practice: every cross-CTA
    // mutable value is read with ld.global.cg / ld.relaxed.gpu (served by L2, never a stale L1 line), the producers' stores are
    // released by the red.release above, and a thread's later loads cannot issue before the branch on the polled counter resolves.
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


}  // namespace
}  // namespace vx
