// Persistent AR decode kernel: the whole token loop of models/vallex.py:528-598 in ONE cooperative launch.
//
// One CTA per SM (256 threads), all resident; every decode step walks
//     S0  sample + append + embed (CTA b serves utterance b)              models/vallex.py:529-531,569-598
//     12 x [ P1 LN1 + QKV rows | P2 attention (KV append, keys split over CTAs) | P3 split-combine + out-proj + residual
//            | P4 LN2 + FFN1 + ReLU | P5 FFN2 (split-K x4, partial sums) ]  modules/transformer.py:337-347
//     PF  final LN + predict rows -> logits                               modules/transformer.py:470-471
// separated by grid-wide barriers (release/relaxed counter in L2), 62 per step.  Nothing returns to the host between
// tokens: sampling, EOS/budget bookkeeping and the all-finished early exit live on the device.
//
// Data path (HBM / latency bound):
//  * Weights.  Each CTA owns a fixed, contiguous row slice of every matrix (row-major bf16/fp32).  A slice is at
//    most 4 rows per warp, so it is requested into REGISTERS with 128-bit streaming loads BEFORE the grid barrier that
//    precedes its phase (weights do not depend on the barrier) together with its biases and LayerNorm vectors; the
//    phase itself then only stages the activations (fp32, B x 1024, shared memory), runs the FMAs and a warp-shuffle
//    reduction.  FFN2 is split over K so every GEMV phase has the same shape; its four partial sums, the bias and
//    the residual are folded (fixed order -> deterministic) by the consumers.
//  * KV cache ([layer][b][head][pos][64], 128-byte rows in bf16).  The CTA's key range of the coming attention phase
//    is contiguous; it is staged into shared memory by the TMA engine (cp.async.bulk + mbarrier) while the QKV phase
//    runs.  Scores use 8 lanes per key (conflict-free 128-bit shared loads), softmax statistics are block
//    reductions, P.V accumulates in registers; this step's key/value are appended in place by split 0.
//
// Algorithmic bytes per step: 304,412,672 + 49,152 * sum_b(L_b + 1) at bf16.
#include <cstdlib>

#include "common.cuh"
#include "decode_persist.h"
#include "kernels.h"
#include "tc_common.cuh"

namespace vx {

constexpr int PNT = 256;            // threads per CTA
constexpr int PNW = PNT / 32;       // warps
constexpr int PCHUNK = 2048;        // max keys per attention unit
constexpr int PSORT = 2048;
constexpr int KVS_KEYS = 256;       // keys per unit that can be staged in shared memory (2 x 32 KB)

// ---- grid barrier -------------------------------------------------------------------------------------
// `prof` (nullable): (tag, clock64) stamps -- 1 = arrival of this CTA, 2 = release of the grid
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
      if (++spins > (1u << 23)) __trap();  // a lost CTA must not hang the box (~seconds)
    } while ((int)(v - epoch) < 0);
    // no acquire fence: every cross-CTA mutable value is read with ld.global.cg (L2), never through L1
    if (prof) { prof[0] = 2; prof[1] = clock64(); prof += 2; }
  }
  __syncthreads();
}
#define PSTAMP(id_) do { if (prof) { prof[0] = (id_); prof[1] = clock64(); prof += 2; } } while (0)

__device__ __forceinline__ int ld_cg(const int32_t* p) { return __ldcg(p); }

// VX_DECODE_DBG & 32: record the first non-finite value seen (code = step * 1000 + layer * 10 + site) in steps_done[2..3]
__device__ __forceinline__ void nan_probe(const PParams& p, int code, float v) {
  if ((p.dbg & 32) && !(fabsf(v) <= 3.0e38f)) {
    atomicMin(p.steps_done + 2, code);
    p.steps_done[3] = __float_as_int(v);
  }
}

__device__ __forceinline__ void slice(int total, int& r0, int& r1) {
  r0 = (int)(((long long)blockIdx.x * total) / gridDim.x);
  r1 = (int)(((long long)(blockIdx.x + 1) * total) / gridDim.x);
}

// ---- weight-streaming GEMV with cross-barrier register prefetch ---------------------------------------------
template <typename WT, int MR>
struct WPre {
  static constexpr int kRows = (sizeof(WT) == 2) ? (MR <= 4 ? 4 : (MR <= 8 ? 2 : 1)) : (MR <= 4 ? 2 : 1);
  static constexpr int kVec = (sizeof(WT) == 2) ? 4 : 8;  // uint4 per row per lane (K = 1024)
  uint4 r[kRows][kVec];
  float bias[kRows], bias2[kRows];  // constant addends of each prefetched row (bias, previous layer's FFN2 bias)
};
struct LnPre {  // LayerNorm parameters (+ previous layer's FFN2 bias) of this thread's four columns
  float g[4], b[4], f[4];
};
__device__ __forceinline__ void ln_issue(LnPre& lp, const float* __restrict__ g, const float* __restrict__ b,
                                         const float* __restrict__ fb) {
  const int t = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    lp.g[i] = g[t + i * PNT];
    lp.b[i] = b[t + i * PNT];
    lp.f[i] = fb ? fb[t + i * PNT] : 0.f;
  }
}

__device__ __forceinline__ uint4 ldw_raw(const void* p) {  // streaming: each weight byte is used once per step
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
template <typename WT, int NV>
__device__ __forceinline__ void load_row_raw(const WT* __restrict__ W, size_t ldw, int n, uint4 (&dst)[NV]) {
  const int lane = threadIdx.x & 31;
  const WT* p = W + (size_t)n * ldw + lane * 8;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    if constexpr (sizeof(WT) == 2) dst[it] = ldw_raw(p + it * 256);
    else { dst[2 * it] = ldw_raw(p + it * 256); dst[2 * it + 1] = ldw_raw(p + it * 256 + 4); }
  }
}

// request this warp's rows of slice [r0,r1): row j of the warp is r0 + warp + j*PNW
template <typename WT, int MR>
__device__ __forceinline__ void gemv_issue(WPre<WT, MR>& pre, const WT* __restrict__ W, size_t ldw, int r0, int r1,
                                           const float* __restrict__ bias, const float* __restrict__ bias2) {
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < WPre<WT, MR>::kRows; ++j) {
    const int n = r0 + warp + j * PNW;
    pre.bias[j] = 0.f;
    pre.bias2[j] = 0.f;
    if (n < r1) {
      load_row_raw<WT, WPre<WT, MR>::kVec>(W, ldw, n, pre.r[j]);
      if (bias) pre.bias[j] = bias[n];      // plain loads: consumed only after the barrier, no wait here
      if (bias2) pre.bias2[j] = bias2[n];
    }
  }
}

// consume the prefetched rows (x is read from shared memory once per k block and reused by all rows of the warp).
// addend(n, m) is evaluated by lane m BEFORE the dot products; emit(n, m, value) is called by lane m < B.
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
    add[j] = ((n < r1 && lane < B) ? addend(n, lane) : 0.f) + (pre.bias[j] + pre.bias2[j]);
  }
  float acc[kRows][MR];
#pragma unroll
  for (int j = 0; j < kRows; ++j)
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[j][m] = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    float w[kRows][8];
#pragma unroll
    for (int j = 0; j < kRows; ++j) raw_to_f32<WT>(pre.r[j], it, w[j]);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
      const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
      for (int j = 0; j < kRows; ++j)
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[j][m] = fmaf(w[j][i], xv[i], acc[j][m]);
    }
  }
#pragma unroll
  for (int j = 0; j < kRows; ++j) {
    const int n = r0 + warp + j * PNW;
    if (n < r1) {  // warp-uniform
      float mine = 0.f;
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float s = warp_sum(acc[j][m]);
        if (lane == m) mine = s;
      }
      if (lane < B) emit(n, lane, mine + add[j]);
    }
  }
  for (int n = r0 + warp + kRows * PNW; n < r1; n += PNW) {  // rows beyond the prefetch window (small grids / large B)
    uint4 raw[WPre<WT, MR>::kVec];
    load_row_raw<WT, WPre<WT, MR>::kVec>(W, ldw, n, raw);
    const float ad = ((lane < B) ? addend(n, lane) : 0.f) + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f);
    float a1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) a1[m] = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      float w[8];
      raw_to_f32<WT>(raw, it, w);
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) a1[m] = fmaf(w[i], xv[i], a1[m]);
      }
    }
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float s = warp_sum(a1[m]);
      if (lane == m) mine = s;
    }
    if (lane < B) emit(n, lane, mine + ad);
  }
}

// Stage B rows into shared memory with LayerNorm (eps 1e-5), two-pass variance, all threads (4 columns each).
// value(m, c) is the dynamic part of the stream; lp.f (previous layer's FFN2 bias) is added here.
template <int MR, typename Val>
__device__ __forceinline__ void stage_ln(float* sx, int B, Val value, const LnPre& lp, float* red /*[2][PNW][8]*/) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = MR < 8 ? MR : 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v[R][4];
#pragma unroll
    for (int m = 0; m < R; ++m)
#pragma unroll
      for (int i = 0; i < 4; ++i) v[m][i] = (m < nb) ? value(mb + m, tid + i * PNT) + lp.f[i] : 0.f;
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float s = warp_sum((v[m][0] + v[m][1]) + (v[m][2] + v[m][3]));
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
      float q = 0.f;
#pragma unroll
      for (int i = 0; i < 4; ++i) { const float d = v[m][i] - mean[m]; q += d * d; }
      q = warp_sum(q);
      if (lane == 0) red[PNW * 8 + warp * 8 + m] = q;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) s += red[PNW * 8 + w * 8 + m];
        const float rstd = 1.0f / sqrtf(s / (float)kD + 1e-5f);
#pragma unroll
        for (int i = 0; i < 4; ++i)
          sx[(mb + m) * kD + tid + i * PNT] = __fadd_rn(__fmul_rn((v[m][i] - mean[m]) * rstd, lp.g[i]), lp.b[i]);
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

// residual stream entering layer l: xa + the previous layer's four FFN2 split-K partials (its bias, a constant, is
// added from prefetched registers by the callers)
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
  float* scan = reinterpret_cast<float*>(sidx + PSORT);    // [PNT]
  float* red = scan + PNT;                  // [32]
  int* s_tok = reinterpret_cast<int*>(red + 32);           // [2]
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
  const int Tb = ld_cg(bt.T + b);
  int g_tok = -1;
  float g_logprob = 0.f, m = -INFINITY, Z = 1.f;
  if (greedy) {
    // top_k == 1 keeps only the maximum (models/vallex.py:803-809): the draw is its index (first one on an exact tie)
    // and log_softmax over the kept entries is -log(#ties).
    for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
    m = block_max(m, red);
    int first = 0x00ffffff, cnt = 0;
    for (int i = tid; i < kARLogits; i += PNT)
      if (v[i] == m) { first = min(first, i); ++cnt; }
    first = -(int)block_max((float)(-first), red);   // indices < 2^24: exact in fp32
    cnt = (int)block_sum((float)cnt, red);
    g_tok = first;
    g_logprob = -logf((float)cnt);
    if (tid == 0) *s_tok = g_tok;
    __syncthreads();
  } else {
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
        const float m0 = sv[0];
        float part = 0.f;
        for (int i = tid; i < kARLogits; i += PNT) part += __expf(sv[i] - m0);
        const float Z0 = block_sum(part, red);
        if (tid == 0) {
          float c = 0.f;
          bool drop_next = false;
          for (int i = 0; i < kARLogits; ++i) {
            const bool drop = drop_next;
            c += __expf(sv[i] - m0) / Z0;
            drop_next = (c > p.top_p);
            if (drop) v[sidx[i]] = -INFINITY;
          }
        }
        __syncthreads();
      }
    }
    for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
    m = block_max(m, red);
    constexpr int CH = (kARLogits + PNT - 1) / PNT;  // 5
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
  }
  if (tid == 0) {
    int tok = *s_tok;
    if (tok < 0)  // rounding at the CDF tail: fall back to the last token with mass
      for (int i = kARLogits - 1; i >= 0; --i)
        if (v[i] > -INFINITY) { tok = i; break; }
    if (!(tok >= 0 && tok < kARLogits)) {  // non-finite logits: never index the embedding table out of range
      if (p.dbg & 32) { atomicMin(p.steps_done + 2, Tb * 1000 + 999); p.steps_done[3] = __float_as_int(m); }
      tok = 0;
    }
    const float logprob = greedy ? g_logprob : (v[tok] - m) - logf(Z);
    const int prev = ld_cg(bt.y + (size_t)b * bt.ymax + bt.P[b] + Tb);
    if (prev != kEOS) bt.sum_logprobs[b] += logprob;  // models/vallex.py:572
    p.out_tokens[(size_t)b * p.out_stride + Tb] = tok;
    int fed = -1;
    if ((tok == kEOS) && !p.ignore_eos) {
      bt.finished[b] = 1;
      atomicSub(bt.n_active, 1);
    } else {
      const int f = p.forced ? p.forced[(size_t)b * p.forced_stride + Tb] : tok;
      bt.y[(size_t)b * bt.ymax + bt.P[b] + 1 + Tb] = f;
      bt.T[b] = Tb + 1;
      if (Tb + 1 >= bt.budget[b]) {
        bt.finished[b] = 1;
        atomicSub(bt.n_active, 1);
      } else {
        fed = f;
      }
    }
    s_tok[0] = fed;
    s_tok[1] = Tb + 1;
  }
  __syncthreads();
  const int fed = s_tok[0];
  if (fed >= 0) {  // next input row: emb[token] + alpha * pe[y position]   (models/vallex.py:529-531)
    const int j = bt.P[b] + s_tok[1];
    const float a = p.alpha[0];
    for (int c = tid; c < kD; c += PNT)
      p.xa[(size_t)b * kD + c] = __fadd_rn(__fmul_rn(p.audio_emb[(size_t)fed * kD + c], 1.0f), __fmul_rn(a, p.pe[(size_t)j * kD + c]));
  }
  __syncthreads();
}

// ---- decode attention unit (b, h, split) ------------------------------------------------------------------------
template <typename KVT> __device__ __forceinline__ float kv_round(float v);
template <> __device__ __forceinline__ float kv_round<float>(float v) { return v; }
template <> __device__ __forceinline__ float kv_round<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

// 8 consecutive cache elements -> fp32.  KV rows are immutable once written and never read before they are
// written, so ordinary (L1-allocating) loads are safe; the shared-memory path uses the same generic load.
__device__ __forceinline__ void kv8(const bf16* p, float (&o)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { o[2 * i] = __uint_as_float(w[i] << 16); o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u); }
}
__device__ __forceinline__ void kv8(const float* p, float (&o)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}

struct AttnRange { int b, h, split, pos, j0, j1; };
__device__ __forceinline__ AttnRange attn_range(const PParams& p, int u) {
  AttnRange r;
  r.split = u % p.nsplit;
  const int bh = u / p.nsplit;
  r.b = bh / kH; r.h = bh % kH;
  r.pos = p.bt.S[r.b] + p.bt.P[r.b] + ld_cg(p.bt.T + r.b);  // cache slot of this step's key
  const int n = r.pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  r.j0 = r.split * chunk;
  r.j1 = min(n, r.j0 + chunk);
  return r;
}

// staged: K rows [j0, j0 + n_staged) at ks, V rows at vs (shared memory, row-major [key][64]); otherwise global.
template <typename KVT>
__device__ void attn_unit(const PParams& p, int layer, const AttnRange& r, float* smem, const KVT* ks, const KVT* vs,
                          int n_staged) {
  float* sq = smem;                  // [64]
  float* sknew = smem + 64;          // [64]
  float* svnew = smem + 128;         // [64]
  float* red = smem + 192;           // [32]
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const int tid = threadIdx.x;
  const int b = r.b, h = r.h, pos = r.pos, j0 = r.j0, j1 = r.j1;
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();  // scratch reuse across units
  if (tid < kDH) {
    sq[tid] = __ldcg(row + tid) * 0.125f;   // 1/sqrt(64), modules/activation.py:161
    sknew[tid] = __ldcg(row + kD + tid);
    svnew[tid] = __ldcg(row + 2 * kD + tid);
  }
  __syncthreads();
  KVT* kb = reinterpret_cast<KVT*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  KVT* vb = reinterpret_cast<KVT*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (r.split == 0 && tid < kDH) {  // append (rounded to the cache type once; later steps read the same bits)
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  // scores: 8 lanes per key, 8 dims per lane
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float qv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) qv[i] = sq[dc + i];
  float mloc = -INFINITY;
  for (int jb = j0; jb < j1; jb += PNT / 8) {  // warp-uniform trip count: the shuffles below need every lane
    const int j = jb + g;
    const bool valid = j < j1;
    float w[8];
    if (!valid) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = 0.f;
    } else if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<KVT>(sknew[dc + i]);
    } else if (j - j0 < n_staged) {
      kv8(ks + (size_t)(j - j0) * kDH + dc, w);
    } else {
      kv8(kb + (size_t)j * kDH + dc, w);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(qv[i], w[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (valid) {
      if ((tid & 7) == 0) sp[j - j0] = s;
      mloc = fmaxf(mloc, s);
    }
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);  // also orders the sp[] writes before the PV reads
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += PNT / 8) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<KVT>(svnew[dc + i]);
    } else if (j - j0 < n_staged) {
      kv8(vs + (size_t)(j - j0) * kDH + dc, w);
    } else {
      kv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  // partial (m, l, o[64]) of this split; the out-proj phase folds the splits while staging (stage_att)
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + r.split) * (kDH + 2);
  if (tid < kDH) {
    float od = 0.f;
#pragma unroll 8
    for (int q = 0; q < PNT / 8; ++q) od += sacc[q * kDH + tid];
    pp[2 + tid] = od;
  }
  if (tid == 0) { pp[0] = m; pp[1] = l; }
}

// out-proj input: softmax-combine the key splits.  The normalised split weights of every (b, head) go to shared
// memory first (one exp per split, not per column), then every column folds its splits with coalesced loads.
__device__ __forceinline__ void stage_att(float* sx, const PParams& p, int B, float* swn /*[B*16*nsplit] + [2*B*16*nsplit]*/) {
  const int tid = threadIdx.x;
  const int ns = p.nsplit;
  const int npairs = B * kH * ns;
  float* sml = swn + npairs;  // (m, l) of every (b, head, split): one L2 round trip for all of them
  for (int i = tid; i < 2 * npairs; i += PNT) sml[i] = __ldcg(p.part + (size_t)(i >> 1) * (kDH + 2) + (i & 1));
  __syncthreads();
  for (int bh = tid; bh < B * kH; bh += PNT) {
    const float* ml = sml + (size_t)bh * ns * 2;
    float M = -INFINITY;
    for (int s = 0; s < ns; ++s) M = fmaxf(M, ml[2 * s]);
    float L = 0.f;
    for (int s = 0; s < ns; ++s) {
      const float w = (ml[2 * s] == -INFINITY) ? 0.f : expf(ml[2 * s] - M);
      swn[bh * ns + s] = w;
      L += w * ml[2 * s + 1];
    }
    const float inv = 1.0f / L;
    for (int s = 0; s < ns; ++s) swn[bh * ns + s] *= inv;
  }
  __syncthreads();
  for (int i = tid; i < B * kD; i += PNT) {
    const int b = i / kD, c = i % kD, h = c >> 6, d = c & 63;
    const float* wn = swn + (size_t)(b * kH + h) * ns;
    const float* po = p.part + ((size_t)(b * kH + h) * ns) * (kDH + 2) + 2 + d;
    float o[16];
#pragma unroll
    for (int s = 0; s < 16; ++s) o[s] = (s < ns) ? __ldcg(po + (size_t)s * (kDH + 2)) : 0.f;  // all splits in flight at once
    float O = 0.f;
#pragma unroll
    for (int s = 0; s < 16; ++s)
      if (s < ns) O = fmaf(wn[s], o[s], O);
    for (int s = 16; s < ns; ++s) O = fmaf(wn[s], __ldcg(po + (size_t)s * (kDH + 2)), O);
    nan_probe(p, 900000000 + h, O);
    sx[i] = ld_cg(p.bt.finished + b) ? 0.f : O;
  }
  __syncthreads();
}

// ---- the kernel -------------------------------------------------------------------------------------------
template <typename WT, int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_persistent_kernel(const PParams p) {
  extern __shared__ __align__(128) uint8_t dsm_raw[];
  // [K stage 32 KB | V stage 32 KB (bf16, B <= 8 only) | sx MR*4 KB | scratch 36 KB | mbarrier]
  constexpr bool kStageKV = (sizeof(WT) == 2) && MR <= 8;  // 128-byte bf16 rows: 256 keys x (K + V) = 64 KB
  constexpr int kKvBytes = kStageKV ? 2 * KVS_KEYS * kDH * 2 : 0;
  WT* ks = reinterpret_cast<WT*>(dsm_raw);
  WT* vs = reinterpret_cast<WT*>(dsm_raw + kKvBytes / 2);
  float* sx = reinterpret_cast<float*>(dsm_raw + kKvBytes);
  float* scratch = sx + MR * kD;
  uint64_t* kvbar = reinterpret_cast<uint64_t*>(scratch + 9216);
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

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

  if (tid == 0) {
    tc::mbar_init(kvbar, 1);
    tc::fence_barrier_init();
  }
  __syncthreads();
  uint32_t kv_parity = 0;

  WPre<WT, MR> pre;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  // everything that does not depend on a barrier is requested BEFORE it: weight rows, their biases, LN parameters
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  // optional timeline: CTAs 0, grid/2 and grid-1 stamp steps [prof_step0, prof_step0 + 4)
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
      if (!ld_cg(bt.finished + b) && !(p.dbg & 16)) sample_utterance(p, b, scratch);
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
      // ---- P1: LN1 + QKV rows.  First: TMA-stage the K/V rows of this CTA's coming attention unit. -------------
      int n_staged = 0;
      const bool have_unit = (int)blockIdx.x < B * kH * p.nsplit;
      if (kStageKV && have_unit) {
        const AttnRange r = attn_range(p, blockIdx.x);
        const int cnt = min(r.j1, r.pos) - r.j0;  // the new key (pos) is not in the cache yet
        if (cnt > 0 && cnt <= KVS_KEYS && !ld_cg(bt.finished + r.b)) {
          n_staged = cnt;
          if (tid == 0) {
            const size_t off = (size_t)l * p.kv_layer_elems + ((size_t)r.b * kH + r.h) * p.ctx_cap * kDH + (size_t)r.j0 * kDH;
            tc::fence_proxy_async();
            tc::mbar_expect_tx(kvbar, (uint32_t)cnt * kDH * 2 * 2);
            tc::bulk_g2s(ks, reinterpret_cast<const WT*>(p.kc) + off, (uint32_t)cnt * kDH * 2, kvbar);
            tc::bulk_g2s(vs, reinterpret_cast<const WT*>(p.vc) + off, (uint32_t)cnt * kDH * 2, kvbar);
          }
        }
      }
      if (!(p.dbg & 8)) stage_ln<MR>(sx, B, [&](int m, int c) { const float v = stream_dyn(p, l, m, c, B); nan_probe(p, step * 1000 + l * 10 + 1, v); return v; }, lp, scratch);
      PSTAMP(3);
      if (!(p.dbg & 4)) gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { nan_probe(p, step * 1000 + l * 10 + 2, v); p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, Wo, kD, o0, o1, L.out_b, fb_prev);   // out-proj rows ride through the attention phase
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention ---------------------------------------------------------------------------------
      if (n_staged > 0) {
        tc::mbar_wait(kvbar, kv_parity);
        kv_parity ^= 1;
      }
      PSTAMP(6);
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const AttnRange r = attn_range(p, u);
        if (ld_cg(bt.finished + r.b) || (p.dbg & 1)) continue;
        attn_unit<WT>(p, l, r, scratch, ks, vs, (u == (int)blockIdx.x) ? n_staged : 0);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: combine key splits + out-proj + residual -> xa ---------------------------------------------------
      if (!(p.dbg & 2)) stage_att(sx, p, B, scratch);
      PSTAMP(3);
      if (!(p.dbg & 4)) gemv_finish<WT, MR>(pre, Wo, kD, o0, o1, sx, B, L.out_b, fb_prev, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                          [&](int n, int m, float v) { nan_probe(p, step * 1000 + l * 10 + 4, v); p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      ln_issue(lp, L.n2_g, L.n2_b, nullptr);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU -------------------------------------------------------------------------------
      if (!(p.dbg & 8)) stage_ln<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch);
      PSTAMP(3);
      if (!(p.dbg & 4)) gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { nan_probe(p, step * 1000 + l * 10 + 5, v); p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W2, kFF, g0, g1, nullptr, nullptr);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2, split-K: rows [g0,g1) x k chunk kc -> partial pf[kc] --------------------------------------
      stage_plain(sx, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
      PSTAMP(3);
      if (!(p.dbg & 4)) gemv_finish<WT, MR>(pre, W2, kFF, g0, g1, sx, B, nullptr, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { nan_probe(p, step * 1000 + l * 10 + 6, v); p.pf[((size_t)kc * B + m) * kD + n] = v; });
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
    stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, lp, scratch);
    if (!(p.dbg & 4)) gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <typename WT, int MR>
static cudaError_t persist_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t kv_bytes = (sizeof(WT) == 2 && MR <= 8) ? 2 * (size_t)KVS_KEYS * kDH * 2 : 0;
  const size_t smem = kv_bytes + (size_t)MR * kD * 4 + 9216 * 4 + 64;
  cudaError_t e = cudaFuncSetAttribute(ar_decode_persistent_kernel<WT, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_persistent_kernel<WT, MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_persistent_kernel<WT, MR>, dim3(grid), dim3(PNT), args, smem, st);
}

// wt: 0 fp32 / 1 bf16 (weights and KV cache).  B <= 32.  Returns a cudaError_t as int.
int launch_decode_persistent(const void* params, int wt, int B, int grid, cudaStream_t st) {
  const PParams& p = *reinterpret_cast<const PParams*>(params);
  const int mr = B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : B <= 8 ? 8 : B <= 16 ? 16 : 32;
#define VX_P(WT_, MR_) return (int)persist_launch<WT_, MR_>(p, grid, st)
  if (wt == 0) {
    switch (mr) { case 1: VX_P(float, 1); case 2: VX_P(float, 2); case 4: VX_P(float, 4); case 8: VX_P(float, 8); case 16: VX_P(float, 16); default: VX_P(float, 32); }
  } else {
    switch (mr) { case 1: VX_P(bf16, 1); case 2: VX_P(bf16, 2); case 4: VX_P(bf16, 4); case 8: VX_P(bf16, 8); case 16: VX_P(bf16, 16); default: VX_P(bf16, 32); }
  }
#undef VX_P
}

}  // namespace vx
