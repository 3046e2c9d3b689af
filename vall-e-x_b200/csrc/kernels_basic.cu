// Embedding / position / LayerNorm / sampling kernels of the VALL-E-X hot path.
// All of these are HBM- or latency-bound element-wise/row-wise work: 128-bit coalesced accesses,
// warp-shuffle reductions, no tensor cores.
#include "common.cuh"
#include "kernels.h"

namespace vx {

// ---------------------------------------------------------------------------------------------
// Text rows: x = (text_emb[id] + lang_emb[lang]) * 1.0 + alpha * pe[s]
// reference: models/vallex.py:497-507 (AR), :622-632 (NAR); modules/embedding.py:93-97.
// __fmul_rn/__fadd_rn keep the reference's op order (no FMA contraction) so rows are bit-identical.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_text_kernel(Batch bt, const float* __restrict__ text_emb,
                                                         const float* __restrict__ lang_emb,
                                                         const float* __restrict__ alpha,
                                                         const float* __restrict__ pe,
                                                         const int32_t* __restrict__ row0, float* __restrict__ x) {
  const int b = blockIdx.y, s = blockIdx.x;
  if (s >= bt.S[b]) return;
  const int id = bt.text_ids[b * bt.Smax + s], lg = bt.lang_ids[b * bt.Smax + s];
  const float a = alpha[0];
  const int c = threadIdx.x * 4;
  const float4 e = *reinterpret_cast<const float4*>(text_emb + (size_t)id * kD + c);
  const float4 l = *reinterpret_cast<const float4*>(lang_emb + (size_t)lg * kD + c);
  const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)s * kD + c);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(__fadd_rn(e.x, l.x), 1.0f), __fmul_rn(a, p.x));
  o.y = __fadd_rn(__fmul_rn(__fadd_rn(e.y, l.y), 1.0f), __fmul_rn(a, p.y));
  o.z = __fadd_rn(__fmul_rn(__fadd_rn(e.z, l.z), 1.0f), __fmul_rn(a, p.z));
  o.w = __fadd_rn(__fmul_rn(__fadd_rn(e.w, l.w), 1.0f), __fmul_rn(a, p.w));
  *reinterpret_cast<float4*>(x + ((size_t)row0[b] + s) * kD + c) = o;
}

void launch_embed_text(const Batch& bt, const float* text_emb, const float* lang_emb, const float* alpha,
                       const float* pe, const int32_t* row0, float* x, cudaStream_t st) {
  embed_text_kernel<<<dim3(bt.Smax, bt.B), 256, 0, st>>>(bt, text_emb, lang_emb, alpha, pe, row0, x);
}

// Audio rows of the AR prefill: y = [BOS, prompt codebook 0]; x = emb[y_j] + alpha * pe[j]
// (positions restart at 0 and include BOS; models/vallex.py:515-517,529-531).
__global__ void __launch_bounds__(256) embed_audio_prefill_kernel(Batch bt, const float* __restrict__ audio_emb,
                                                                  const float* __restrict__ alpha,
                                                                  const float* __restrict__ pe,
                                                                  float* __restrict__ x) {
  const int b = blockIdx.y, j = blockIdx.x;
  if (j > bt.P[b]) return;
  const int tok = bt.y[(size_t)b * bt.ymax + j];
  const float a = alpha[0];
  const int c = threadIdx.x * 4;
  const float4 e = *reinterpret_cast<const float4*>(audio_emb + (size_t)tok * kD + c);
  const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)j * kD + c);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(e.x, 1.0f), __fmul_rn(a, p.x));
  o.y = __fadd_rn(__fmul_rn(e.y, 1.0f), __fmul_rn(a, p.y));
  o.z = __fadd_rn(__fmul_rn(e.z, 1.0f), __fmul_rn(a, p.z));
  o.w = __fadd_rn(__fmul_rn(e.w, 1.0f), __fmul_rn(a, p.w));
  *reinterpret_cast<float4*>(x + ((size_t)bt.ar_row0[b] + bt.S[b] + j) * kD + c) = o;
}

void launch_embed_audio_prefill(const Batch& bt, const float* audio_emb, const float* alpha, const float* pe,
                                float* x, cudaStream_t st) {
  embed_audio_prefill_kernel<<<dim3(bt.Pmax + 1, bt.B), 256, 0, st>>>(bt, audio_emb, alpha, pe, x);
}

// y[b][0] = BOS, y[b][1..P] = prompt[...,0]   (models/vallex.py:515-517)
__global__ void init_y_kernel(Batch bt) {
  const int b = blockIdx.x;
  const int P = bt.P[b];
  for (int j = threadIdx.x; j <= P; j += blockDim.x)
    bt.y[(size_t)b * bt.ymax + j] = (j == 0) ? kBOS : bt.prompt[((size_t)b * bt.Pmax + (j - 1)) * kNQ];
}
void launch_init_y(const Batch& bt, cudaStream_t st) { init_y_kernel<<<bt.B, 256, 0, st>>>(bt); }

// One decode step's input row: the newest token of y at its y-position (models/vallex.py:529-531,552-553).
__global__ void __launch_bounds__(256) embed_audio_step_kernel(Batch bt, const float* __restrict__ audio_emb,
                                                               const float* __restrict__ alpha,
                                                               const float* __restrict__ pe,
                                                               float* __restrict__ xstep) {
  const int b = blockIdx.x;
  const int j = bt.P[b] + bt.T[b];  // index of the last token in y (BOS at 0)
  const int tok = bt.y[(size_t)b * bt.ymax + j];
  const float a = alpha[0];
  const int c = threadIdx.x * 4;
  const float4 e = *reinterpret_cast<const float4*>(audio_emb + (size_t)tok * kD + c);
  const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)j * kD + c);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(e.x, 1.0f), __fmul_rn(a, p.x));
  o.y = __fadd_rn(__fmul_rn(e.y, 1.0f), __fmul_rn(a, p.y));
  o.z = __fadd_rn(__fmul_rn(e.z, 1.0f), __fmul_rn(a, p.z));
  o.w = __fadd_rn(__fmul_rn(e.w, 1.0f), __fmul_rn(a, p.w));
  *reinterpret_cast<float4*>(xstep + (size_t)b * kD + c) = o;
}
void launch_embed_audio_step(const Batch& bt, const float* audio_emb, const float* alpha, const float* pe,
                             float* xstep, cudaStream_t st) {
  embed_audio_step_kernel<<<bt.B, 256, 0, st>>>(bt, audio_emb, alpha, pe, xstep);
}

// ---------------------------------------------------------------------------------------------
// NAR embeddings (models/vallex.py:605-607,659-662,682-683)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 emb_row4(const EmbTables& t, int j, int tok, int c) {
  if (t.bf[j]) {
    float o[4];
    load4(reinterpret_cast<const bf16*>(t.p[j]) + (size_t)tok * kD + c, o);
    return make_float4(o[0], o[1], o[2], o[3]);
  }
  return *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(t.p[j]) + (size_t)tok * kD + c);
}

// yemb[j] = E0[y[1+j]]; prompt rows additionally += E_1[prompt_1] += ... += E_7[prompt_7] in that order.
__global__ void __launch_bounds__(256) nar_embed_init_kernel(Batch bt, EmbTables tab, float* __restrict__ yemb,
                                                             const int32_t* __restrict__ yrow0) {
  const int b = blockIdx.y, j = blockIdx.x;
  const int P = bt.P[b];
  if (j >= P + bt.T[b]) return;
  const int c = threadIdx.x * 4;
  float4 v = emb_row4(tab, 0, bt.y[(size_t)b * bt.ymax + 1 + j], c);
  if (j < P) {
    for (int q = 1; q < kNQ; ++q) {
      const float4 e = emb_row4(tab, q, bt.prompt[((size_t)b * bt.Pmax + j) * kNQ + q], c);
      v.x = __fadd_rn(v.x, e.x); v.y = __fadd_rn(v.y, e.y); v.z = __fadd_rn(v.z, e.z); v.w = __fadd_rn(v.w, e.w);
    }
  }
  *reinterpret_cast<float4*>(yemb + ((size_t)yrow0[b] + j) * kD + c) = v;
}
void launch_nar_embed_init(const Batch& bt, const EmbTables& tab, float* yemb, const int32_t* yrow0, int max_pt,
                           cudaStream_t st) {
  if (max_pt <= 0) return;
  nar_embed_init_kernel<<<dim3(max_pt, bt.B), 256, 0, st>>>(bt, tab, yemb, yrow0);
}

// x[audio rows] = yemb * 1.0 + alpha * pe[j]  (nar_audio_position; models/vallex.py:670-672)
__global__ void __launch_bounds__(256) nar_build_x_kernel(Batch bt, const float* __restrict__ yemb,
                                                          const int32_t* __restrict__ yrow0,
                                                          const float* __restrict__ alpha,
                                                          const float* __restrict__ pe, float* __restrict__ x) {
  const int b = blockIdx.y, j = blockIdx.x;
  if (j >= bt.P[b] + bt.T[b]) return;
  const float a = alpha[0];
  const int c = threadIdx.x * 4;
  const float4 e = *reinterpret_cast<const float4*>(yemb + ((size_t)yrow0[b] + j) * kD + c);
  const float4 p = *reinterpret_cast<const float4*>(pe + (size_t)j * kD + c);
  float4 o;
  o.x = __fadd_rn(__fmul_rn(e.x, 1.0f), __fmul_rn(a, p.x));
  o.y = __fadd_rn(__fmul_rn(e.y, 1.0f), __fmul_rn(a, p.y));
  o.z = __fadd_rn(__fmul_rn(e.z, 1.0f), __fmul_rn(a, p.z));
  o.w = __fadd_rn(__fmul_rn(e.w, 1.0f), __fmul_rn(a, p.w));
  *reinterpret_cast<float4*>(x + ((size_t)bt.nar_row0[b] + bt.S[b] + j) * kD + c) = o;
}
void launch_nar_build_x(const Batch& bt, const float* yemb, const int32_t* yrow0, const float* alpha,
                        const float* pe, float* x, int max_pt, cudaStream_t st) {
  if (max_pt <= 0) return;
  nar_build_x_kernel<<<dim3(max_pt, bt.B), 256, 0, st>>>(bt, yemb, yrow0, alpha, pe, x);
}

// argmax over the 1024 logits of one target frame (first max wins, like torch.argmax on CPU), record the
// code and add the next codebook's embedding row to yemb (models/vallex.py:677-683).
__global__ void __launch_bounds__(256) nar_argmax_update_kernel(Batch bt, const float* __restrict__ logits,
                                                                const int32_t* __restrict__ trow0, int stage,
                                                                EmbTables tab, float* __restrict__ yemb,
                                                                const int32_t* __restrict__ yrow0,
                                                                int32_t* __restrict__ codes, int Tmax) {
  const int b = blockIdx.y, t = blockIdx.x;
  if (t >= bt.T[b]) return;
  __shared__ float sv[8];
  __shared__ int si[8];
  const float* row = logits + ((size_t)trow0[b] + t) * kVocab;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < kVocab; i += 256) {
    const float v = row[i];
    if (v > best || (v == best && i < bi)) { best = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > best || (ov == best && oi < bi)) { best = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = best; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  best = sv[0]; bi = si[0];
#pragma unroll
  for (int w = 1; w < 8; ++w)
    if (sv[w] > best || (sv[w] == best && si[w] < bi)) { best = sv[w]; bi = si[w]; }
  if (threadIdx.x == 0) codes[((size_t)b * Tmax + t) * kNQ + stage + 1] = bi;
  if (stage < kNQ - 2) {
    const int c = threadIdx.x * 4;
    const float4 e = emb_row4(tab, stage + 1, bi, c);
    float4* dst = reinterpret_cast<float4*>(yemb + ((size_t)yrow0[b] + bt.P[b] + t) * kD + c);
    float4 v = *dst;
    v.x = __fadd_rn(v.x, e.x); v.y = __fadd_rn(v.y, e.y); v.z = __fadd_rn(v.z, e.z); v.w = __fadd_rn(v.w, e.w);
    *dst = v;
  }
}
void launch_nar_argmax_update(const Batch& bt, const float* logits, const int32_t* trow0, int stage,
                              const EmbTables& tab, float* yemb, const int32_t* yrow0, int32_t* codes, int Tmax,
                              cudaStream_t st) {
  if (Tmax <= 0) return;
  nar_argmax_update_kernel<<<dim3(Tmax, bt.B), 256, 0, st>>>(bt, logits, trow0, stage, tab, yemb, yrow0, codes, Tmax);
}

// codes[b][t][0] = generated codebook-0 token (models/vallex.py:600)
__global__ void nar_copy_cb0_kernel(Batch bt, int32_t* __restrict__ codes, int Tmax) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < bt.T[b]) codes[((size_t)b * Tmax + t) * kNQ] = bt.y[(size_t)b * bt.ymax + 1 + bt.P[b] + t];
}
void launch_nar_copy_cb0(const Batch& bt, int32_t* codes, int Tmax, cudaStream_t st) {
  if (Tmax <= 0) return;
  nar_copy_cb0_kernel<<<dim3((Tmax + 255) / 256, bt.B), 256, 0, st>>>(bt, codes, Tmax);
}

// ---------------------------------------------------------------------------------------------
// LayerNorm / AdaptiveLayerNorm rows (modules/transformer.py:57-74, 93-108; eps 1e-5; Vocos 1e-6).
// out[r] = ada_w * (LN(x[src(r)]) * gamma + beta) + ada_b ; gamma/beta/ada_* nullable.
// One 128-thread block per row, values held in registers (dim <= 1024), two-pass variance.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) layernorm_kernel(const float* __restrict__ x, const int32_t* __restrict__ row_map,
                                                        int dim, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta,
                                                        const float* __restrict__ ada_w,
                                                        const float* __restrict__ ada_b, float eps,
                                                        float* __restrict__ out) {
  __shared__ float red[32];
  const int64_t r = blockIdx.x;
  const int64_t src = row_map ? row_map[r] : r;
  const float* xr = x + src * dim;
  float v[8];
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    v[i] = (c < dim) ? xr[c] : 0.f;
    s += v[i];
  }
  const float mean = block_sum(s, red) / (float)dim;
  float q = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    const float d = (c < dim) ? v[i] - mean : 0.f;
    q += d * d;
  }
  const float var = block_sum(q, red) / (float)dim;
  const float rstd = 1.0f / sqrtf(var + eps);
  float* o = out + r * dim;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = threadIdx.x + i * 128;
    if (c < dim) {
      float y = (v[i] - mean) * rstd;
      if (gamma) y = __fadd_rn(__fmul_rn(y, gamma[c]), beta[c]);
      if (ada_w) y = __fadd_rn(__fmul_rn(ada_w[c], y), ada_b[c]);
      o[c] = y;
    }
  }
}
void launch_layernorm(const float* x, const int32_t* row_map, int64_t rows, int dim, const float* gamma,
                      const float* beta, const float* ada_w, const float* ada_b, float eps, float* out,
                      cudaStream_t st) {
  if (rows <= 0) return;
  layernorm_kernel<<<(unsigned)rows, 128, 0, st>>>(x, row_map, dim, gamma, beta, ada_w, ada_b, eps, out);
}

// ---------------------------------------------------------------------------------------------
// Sampling + AR bookkeeping: one block per utterance over its 1025 logits.
// reference: topk_sampling / top_k_top_p_filtering (models/vallex.py:791-853) and the loop tail
// (models/vallex.py:569-598).  Philox4x32-10 supplies the uniform for the multinomial draw.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t mulhi32(uint32_t a, uint32_t b) { return __umulhi(a, b); }
__device__ float philox_uniform(uint64_t seed, uint32_t c0, uint32_t c1) {
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  uint32_t x0 = c0, x1 = c1, x2 = 0x9E3779B9u, x3 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t h0 = mulhi32(0xD2511F53u, x0), l0 = 0xD2511F53u * x0;
    const uint32_t h1 = mulhi32(0xCD9E8D57u, x2), l1 = 0xCD9E8D57u * x2;
    x0 = h1 ^ x1 ^ k0; x1 = l1; x2 = h0 ^ x3 ^ k1; x3 = l0;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  return ((float)(x0 >> 8) + 0.5f) * (1.0f / 16777216.0f);  // (0,1)
}

constexpr int kSortN = 2048;

__global__ void __launch_bounds__(256) sample_kernel(Batch bt, const float* __restrict__ logits, int top_k,
                                                     float top_p, float temperature, uint64_t seed, int step,
                                                     int ignore_eos, const int32_t* __restrict__ forced,
                                                     int forced_stride, int32_t* __restrict__ out_tokens,
                                                     int out_stride) {
  __shared__ float v[kARLogits + 7];
  __shared__ float sv[kSortN];
  __shared__ int16_t sidx[kSortN];
  __shared__ float red[32];
  __shared__ float scan[256];
  __shared__ int s_tok;
  const int b = blockIdx.x, tid = threadIdx.x;
  if (bt.finished[b]) return;  // uniform per block
  const float* lg = logits + (size_t)b * kARLogits;
  for (int i = tid; i < kARLogits; i += 256) v[i] = (temperature != 1.0f) ? lg[i] / temperature : lg[i];
  __syncthreads();

  const bool need_sort = (top_k > 1) || (top_p < 1.0f);
  if (top_k == 1) {
    float m = -INFINITY;
    for (int i = tid; i < kARLogits; i += 256) m = fmaxf(m, v[i]);
    m = block_max(m, red);
    for (int i = tid; i < kARLogits; i += 256)
      if (v[i] < m) v[i] = -INFINITY;
    __syncthreads();
  }
  if (need_sort) {
    // bitonic sort, descending, of (value, index) padded to 2048 with -inf
    for (int i = tid; i < kSortN; i += 256) {
      sv[i] = (i < kARLogits) ? v[i] : -INFINITY;
      sidx[i] = (int16_t)i;
    }
    __syncthreads();
    for (int k = 2; k <= kSortN; k <<= 1)
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < kSortN; i += 256) {
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
    if (top_k > 1) {
      const int k = min(top_k, kARLogits);
      const float thr = sv[k - 1];
      for (int i = tid; i < kARLogits; i += 256)
        if (v[i] < thr) v[i] = -INFINITY;
      for (int i = tid; i < kSortN; i += 256)
        if (sv[i] < thr) sv[i] = -INFINITY;
      __syncthreads();
    }
    if (top_p < 1.0f) {
      // softmax over the sorted values, inclusive cumsum, drop entries whose PREDECESSOR's cum > top_p
      const float m = sv[0];
      float part = 0.f;
      for (int i = tid; i < kARLogits; i += 256) part += __expf(sv[i] - m);
      const float Z = block_sum(part, red);
      if (tid == 0) {  // 1025 serial adds: negligible next to a decode step, and exactly ordered
        float c = 0.f;
        bool drop_next = false;
        for (int i = 0; i < kARLogits; ++i) {
          const bool drop = drop_next;
          c += __expf(sv[i] - m) / Z;
          drop_next = (c > top_p);
          if (drop) v[sidx[i]] = -INFINITY;
        }
      }
      __syncthreads();
    }
  }

  // softmax statistics of the filtered logits
  float m = -INFINITY;
  for (int i = tid; i < kARLogits; i += 256) m = fmaxf(m, v[i]);
  m = block_max(m, red);
  // multinomial by inverse CDF in index order; each thread owns a contiguous chunk of 5
  float loc[5];
  float s = 0.f;
#pragma unroll
  for (int q = 0; q < 5; ++q) {
    const int i = tid * 5 + q;
    loc[q] = (i < kARLogits) ? expf(v[i] - m) : 0.f;
    s += loc[q];
  }
  scan[tid] = s;
  __syncthreads();
  if (tid < 32) {  // exclusive scan of 256 partial sums by one warp (8 per lane)
    float run = 0.f, mine[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) { mine[q] = scan[tid * 8 + q]; run += mine[q]; }
    float incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += n;
    }
    float ex = incl - run;
#pragma unroll
    for (int q = 0; q < 8; ++q) { scan[tid * 8 + q] = ex; ex += mine[q]; }
    if (tid == 31) red[0] = incl;
  }
  if (tid == 0) s_tok = -1;
  __syncthreads();
  const float Z = red[0];
  const float u = philox_uniform(seed, (uint32_t)step, (uint32_t)b) * Z;
  {
    float c = scan[tid];
    int cand = -1;
#pragma unroll
    for (int q = 0; q < 5; ++q) {
      const float n = c + loc[q];
      if (cand < 0 && loc[q] > 0.f && u <= n && u > c) cand = tid * 5 + q;
      c = n;
    }
    if (cand >= 0) atomicMax(&s_tok, cand);
  }
  __syncthreads();
  if (tid == 0) {
    int tok = s_tok;
    if (tok < 0) {  // rounding at the CDF tail: fall back to the last token with mass
      for (int i = kARLogits - 1; i >= 0; --i)
        if (v[i] > -INFINITY) { tok = i; break; }
    }
    const float logprob = (v[tok] - m) - logf(Z);
    const int Tb = bt.T[b];
    const int prev = bt.y[(size_t)b * bt.ymax + bt.P[b] + Tb];
    if (prev != kEOS) bt.sum_logprobs[b] += logprob;  // models/vallex.py:572
    out_tokens[(size_t)b * out_stride + Tb] = tok;
    const bool eos = (tok == kEOS) && !ignore_eos;
    if (eos) {
      bt.finished[b] = 1;
      atomicSub(bt.n_active, 1);
    } else {
      const int fed = forced ? forced[(size_t)b * forced_stride + Tb] : tok;
      bt.y[(size_t)b * bt.ymax + bt.P[b] + 1 + Tb] = fed;
      bt.T[b] = Tb + 1;
      if (Tb + 1 >= bt.budget[b]) {
        bt.finished[b] = 1;
        atomicSub(bt.n_active, 1);
      }
    }
  }
}
void launch_sample(const Batch& bt, const float* logits, int top_k, float top_p, float temperature, uint64_t seed,
                   int step, int ignore_eos, const int32_t* forced, int forced_stride, int32_t* out_tokens,
                   int out_stride, cudaStream_t st) {
  sample_kernel<<<bt.B, 256, 0, st>>>(bt, logits, top_k, top_p, temperature, seed, step, ignore_eos, forced,
                                      forced_stride, out_tokens, out_stride);
}

}  // namespace vx
