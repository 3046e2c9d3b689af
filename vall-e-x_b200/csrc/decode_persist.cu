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

#include "common.cuh"
#include "kernels.h"
#include "decode_persist.h"

namespace vx {

constexpr int PNT = 512;            // threads per CTA
constexpr int PNW = PNT / 32;       // warps
constexpr int PCHUNK = 2048;        // max keys per attention unit
constexpr int PSORT = 2048;

// ---- grid barrier -------------------------------------------------------------------------------------
__device__ __forceinline__ void grid_barrier(unsigned int* bar, unsigned int& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += gridDim.x;
    asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int v;
    unsigned int spins = 0;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (++spins > (1u << 26)) __trap();  // a lost CTA must not hang the box
    } while ((int)(v - epoch) < 0);
  }
  __syncthreads();
}

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

// rows [r0,r1) of W[*,1024] (row stride ldw elements) against sx[MR][1024] (shared); emit(n, acc[MR]) on lane 0.
template <typename WT, int MR, typename Emit>
__device__ __forceinline__ void gemv_rows(const WT* __restrict__ W, size_t ldw, int r0, int r1,
                                          const float* __restrict__ sx, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr bool kTwo = MR <= 8;  // two rows in flight per warp while the accumulators fit in registers
  for (int n = r0 + warp; n < r1; n += (kTwo ? 2 : 1) * PNW) {
    const int n2 = n + PNW;
    const bool two = kTwo && n2 < r1;
    float w0[4][8], w1[4][8];
    const WT* p0 = W + (size_t)n * ldw + lane * 8;
    const WT* p1 = W + (size_t)(two ? n2 : n) * ldw + lane * 8;
#pragma unroll
    for (int it = 0; it < 4; ++it) ldw8(p0 + it * 256, w0[it]);
    if (kTwo) {
#pragma unroll
      for (int it = 0; it < 4; ++it) ldw8(p1 + it * 256, w1[it]);
    }
    float a0[MR], a1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a0[m] = fmaf(w0[it][i], xv[i], a0[m]);
          if (kTwo) a1[m] = fmaf(w1[it][i], xv[i], a1[m]);
        }
      }
    }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      a0[m] = warp_sum(a0[m]);
      if (kTwo) a1[m] = warp_sum(a1[m]);
    }
    if (lane == 0) {
      emit(n, a0);
      if (two) emit(n2, a1);
    }
  }
}

// LayerNorm of B rows held in shared memory (in place), eps 1e-5.  All threads call.
template <int MR>
__device__ __forceinline__ void ln_rows_smem(float* sx, int B, const float* __restrict__ g, const float* __restrict__ bta,
                                             float* red) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  (void)red;
  for (int m = warp; m < B; m += PNW) {  // one warp per row: 32 values per lane
    float v[32], s = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[i] = sx[m * kD + lane + i * 32]; s += v[i]; }
    const float mean = warp_sum(s) / (float)kD;
    float q = 0.f;
#pragma unroll
    for (int i = 0; i < 32; ++i) { const float d = v[i] - mean; q += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)kD + 1e-5f);
#pragma unroll
    for (int i = 0; i < 32; ++i) {
      const int c = lane + i * 32;
      sx[m * kD + c] = __fadd_rn(__fmul_rn((v[i] - mean) * rstd, g[c]), bta[c]);
    }
  }
  __syncthreads();
}

// residual stream entering a layer: xa (+ previous layer's FFN2 bias and 4 split-K partials, fixed order)
__device__ __forceinline__ float stream_in(const PParams& p, int l, int b, int c, int B) {
  float v = __ldcg(p.xa + (size_t)b * kD + c);
  if (l > 0) {
    v += p.layer[l - 1].ff2_b[c];
#pragma unroll
    for (int s = 0; s < 4; ++s) v += __ldcg(p.pf + ((size_t)s * B + b) * kD + c);
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
  if (p.top_k == 1) {
    float m = -INFINITY;
    for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
    m = block_max(m, red);
    for (int i = tid; i < kARLogits; i += PNT)
      if (v[i] < m) v[i] = -INFINITY;
    __syncthreads();
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
  for (int i = tid; i < kARLogits; i += PNT) m = fmaxf(m, v[i]);
  m = block_max(m, red);
  constexpr int CH = (kARLogits + PNT - 1) / PNT;  // 3
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
  const float Z = red[0];
  const int Tb = ld_cg(bt.T + b);
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
  if (tid == 0) {
    int tok = *s_tok;
    if (tok < 0)
      for (int i = kARLogits - 1; i >= 0; --i)
        if (v[i] > -INFINITY) { tok = i; break; }
    const float logprob = (v[tok] - m) - logf(Z);
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
  float* dst = p.att + (size_t)b * kD + h * kDH;
  if (p.nsplit == 1) {
    if (tid < kDH) dst[tid] = od / l;
    return;
  }
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  __threadfence();
  __syncthreads();
  if (tid == 0) *s_last = (atomicAdd(&p.tickets[b * kH + h], 1) == p.nsplit - 1);
  __syncthreads();
  if (!*s_last) return;
  __threadfence();
  const float* p0 = p.part + ((size_t)b * kH + h) * p.nsplit * (kDH + 2);
  float M = -INFINITY;
  for (int s = 0; s < p.nsplit; ++s) M = fmaxf(M, __ldcg(p0 + s * (kDH + 2)));
  if (tid < kDH) {
    float L = 0.f, O = 0.f;
    for (int s = 0; s < p.nsplit; ++s) {
      const float ms = __ldcg(p0 + s * (kDH + 2));
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
      L += w * __ldcg(p0 + s * (kDH + 2) + 1);
      O += w * __ldcg(p0 + s * (kDH + 2) + 2 + tid);
    }
    dst[tid] = O / L;
  }
  if (tid == 0) p.tickets[b * kH + h] = 0;
}

// ---- the kernel -------------------------------------------------------------------------------------------
template <typename WT, int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_persistent_kernel(const PParams p) {
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* scratch = dsm + MR * kD;        // attention / sampling scratch
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

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    // ---- S0: sample, append, embed -------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    if (tid == 0) l2_prefetch(reinterpret_cast<const char*>(p.layer[0].qkv_w) + (size_t)q0 * kD * wsz, (size_t)(q1 - q0) * kD * wsz);
    grid_barrier(p.bar, epoch);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      // ---- P1: LN1 + QKV rows ---------------------------------------------------------------------------
      if (tid == 0) l2_prefetch(reinterpret_cast<const char*>(L.out_w) + (size_t)o0 * kD * wsz, (size_t)(o1 - o0) * kD * wsz);
      for (int i = tid; i < B * kD; i += PNT) sx[i] = stream_in(p, l, i / kD, i % kD, B);
      for (int i = B * kD + tid; i < MR * kD; i += PNT) sx[i] = 0.f;
      __syncthreads();
      ln_rows_smem<MR>(sx, B, L.n1_g, L.n1_b, nullptr);
      gemv_rows<WT, MR>(reinterpret_cast<const WT*>(L.qkv_w), kD, q0, q1, sx, [&](int n, const float (&a)[MR]) {
        const float bv = L.qkv_b[n];
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < B) p.qkv[(size_t)m * 3 * kD + n] = a[m] + bv;
      });
      grid_barrier(p.bar, epoch);
      // ---- P2: attention ---------------------------------------------------------------------------------
      if (tid == 0) l2_prefetch(reinterpret_cast<const char*>(L.ff1_w) + (size_t)f0 * kD * wsz, (size_t)(f1 - f0) * kD * wsz);
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        if (sizeof(WT) == 2) attn_unit<bf16>(p, l, b, h, split, scratch);
        else attn_unit<float>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch);
      // ---- P3: out-proj + residual -> xa -------------------------------------------------------------------
      for (int i = tid; i < B * kD; i += PNT) sx[i] = __ldcg(p.att + i);
      for (int i = B * kD + tid; i < MR * kD; i += PNT) sx[i] = 0.f;
      __syncthreads();
      gemv_rows<WT, MR>(reinterpret_cast<const WT*>(L.out_w), kD, o0, o1, sx, [&](int n, const float (&a)[MR]) {
        const float bv = L.out_b[n];
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < B) p.xa[(size_t)m * kD + n] = stream_in(p, l, m, n, B) + (a[m] + bv);
      });
      grid_barrier(p.bar, epoch);
      // ---- P4: LN2 + FFN1 + ReLU -------------------------------------------------------------------------------
      if (tid < g1 - g0)  // FFN2 slice: one 1024-wide k chunk of rows [g0,g1): a 2-4 KB segment per row
        l2_prefetch(reinterpret_cast<const char*>(L.ff2_w) + ((size_t)(g0 + tid) * kFF + (size_t)kc * kD) * wsz, (size_t)kD * wsz);
      for (int i = tid; i < B * kD; i += PNT) sx[i] = __ldcg(p.xa + i);
      for (int i = B * kD + tid; i < MR * kD; i += PNT) sx[i] = 0.f;
      __syncthreads();
      ln_rows_smem<MR>(sx, B, L.n2_g, L.n2_b, nullptr);
      gemv_rows<WT, MR>(reinterpret_cast<const WT*>(L.ff1_w), kD, f0, f1, sx, [&](int n, const float (&a)[MR]) {
        const float bv = L.ff1_b[n];
#pragma unroll
        for (int m = 0; m < MR; ++m)
          if (m < B) p.hid[(size_t)m * kFF + n] = fmaxf(a[m] + bv, 0.f);
      });
      grid_barrier(p.bar, epoch);
      // ---- P5: FFN2, split-K: rows [g0,g1) x k chunk kc -> partial pf[kc] --------------------------------------
      if (tid == 0) {
        const void* nxt = (l + 1 < kL) ? p.layer[l + 1].qkv_w : p.predict_w;
        const int a0 = (l + 1 < kL) ? q0 : pr0, a1 = (l + 1 < kL) ? q1 : pr1;
        l2_prefetch(reinterpret_cast<const char*>(nxt) + (size_t)a0 * kD * wsz, (size_t)(a1 - a0) * kD * wsz);
      }
      for (int i = tid; i < B * kD; i += PNT) sx[i] = __ldcg(p.hid + (size_t)(i / kD) * kFF + kc * kD + (i % kD));
      for (int i = B * kD + tid; i < MR * kD; i += PNT) sx[i] = 0.f;
      __syncthreads();
      gemv_rows<WT, MR>(reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD, kFF, g0, g1, sx,
                        [&](int n, const float (&a)[MR]) {
#pragma unroll
                          for (int m = 0; m < MR; ++m)
                            if (m < B) p.pf[((size_t)kc * B + m) * kD + n] = a[m];
                        });
      grid_barrier(p.bar, epoch);
    }
    // ---- PF: final LayerNorm + predict rows -> logits ---------------------------------------------------------
    for (int i = tid; i < B * kD; i += PNT) sx[i] = stream_in(p, kL, i / kD, i % kD, B);
    for (int i = B * kD + tid; i < MR * kD; i += PNT) sx[i] = 0.f;
    __syncthreads();
    ln_rows_smem<MR>(sx, B, p.norm_g, p.norm_b, nullptr);
    gemv_rows<WT, MR>(reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, [&](int n, const float (&a)[MR]) {
#pragma unroll
      for (int m = 0; m < MR; ++m)
        if (m < B) p.logits[(size_t)m * kARLogits + n] = a[m];
    });
    grid_barrier(p.bar, epoch);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
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
#define VX_P(WT_, MR_) return (int)persist_launch<WT_, MR_>(p, grid, st)
  if (wt == 0) {
    switch (mr) { case 1: VX_P(float, 1); case 2: VX_P(float, 2); case 4: VX_P(float, 4); case 8: VX_P(float, 8); case 16: VX_P(float, 16); default: VX_P(float, 32); }
  } else {
    switch (mr) { case 1: VX_P(bf16, 1); case 2: VX_P(bf16, 2); case 4: VX_P(bf16, 4); case 8: VX_P(bf16, 8); case 16: VX_P(bf16, 16); default: VX_P(bf16, 32); }
  }
#undef VX_P
}

}  // namespace vx
