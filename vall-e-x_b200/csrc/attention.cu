// Attention kernels of the VALL-E-X hot path (head dim 64, 16 heads).
//
//  attn_full_kernel   : full-sequence attention over packed rows, fp32, flash-style online softmax.
//                       mode 0 = bidirectional (NAR; modules/activation.py:566-585),
//                       mode 1 = AR prefill mask (models/vallex.py:534-549): text rows see text only,
//                       audio rows see all text + causal audio  <=>  key j visible to query i iff
//                       j < S or j <= i.
//  kv_scatter_kernel  : prefill K/V -> cache [b][h][pos][64] (replaces the torch.cat growth of
//                       modules/activation.py:148-152 with an in-place append).
//  attn_decode_kernel : one query per (utterance, head) against the cache, split over the keys across
//                       CTAs with an in-kernel combine (last-arriver ticket).  HBM-bound: K/V rows are
//                       128 B (bf16) / 256 B (fp32) contiguous, read with 128-bit loads.
#include "common.cuh"
#include "kernels.h"

namespace vx {

constexpr int AQ = 64, AK = 64, APAD = 4;

__global__ void __launch_bounds__(256) attn_full_kernel(Batch bt, const float* __restrict__ qkv,
                                                        const int32_t* __restrict__ row0, int mode,
                                                        float* __restrict__ out) {
  extern __shared__ __align__(16) float attn_smem[];
  float(*sQt)[AQ + APAD] = reinterpret_cast<float(*)[AQ + APAD]>(attn_smem);                          // [d][query]
  float(*sKt)[AK + APAD] = reinterpret_cast<float(*)[AK + APAD]>(attn_smem + kDH * (AQ + APAD));      // [d][key]
  float(*sV)[kDH + APAD] = reinterpret_cast<float(*)[kDH + APAD]>(attn_smem + 2 * kDH * (AQ + APAD)); // [key][d]
  float(*sP)[AK + APAD] = reinterpret_cast<float(*)[AK + APAD]>(attn_smem + 3 * kDH * (AQ + APAD));   // [query][key]
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * AQ;
  const int S = bt.S[b];
  const int len = (mode == 1) ? S + bt.P[b] + 1 : S + bt.P[b] + bt.T[b];
  if (q0 >= len) return;
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const float* base = qkv + (size_t)row0[b] * (3 * kD) + h * kDH;

  // load the Q tile transposed, pre-scaled by 1/sqrt(64) (modules/activation.py:161)
  for (int i = tid; i < AQ * (kDH / 4); i += 256) {
    const int r = i / (kDH / 4), d4 = (i % (kDH / 4)) * 4;
    float4 v = make_float4(0, 0, 0, 0);
    if (q0 + r < len) v = *reinterpret_cast<const float4*>(base + (size_t)(q0 + r) * (3 * kD) + d4);
    sQt[d4 + 0][r] = v.x; sQt[d4 + 1][r] = v.y; sQt[d4 + 2][r] = v.z; sQt[d4 + 3][r] = v.w;
  }
  float m_i[4], l_i[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_i[i] = -INFINITY; l_i[i] = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
  }
  // keys needed by this query tile: bidirectional -> all; AR mask -> max(S, last query + 1)
  int kend = len;
  if (mode == 1) kend = min(len, max(S, q0 + AQ));
  for (int k0 = 0; k0 < kend; k0 += AK) {
    __syncthreads();
    for (int i = tid; i < AK * (kDH / 4); i += 256) {
      const int r = i / (kDH / 4), d4 = (i % (kDH / 4)) * 4;
      float4 kv = make_float4(0, 0, 0, 0), vv = make_float4(0, 0, 0, 0);
      if (k0 + r < len) {
        const float* p = base + (size_t)(k0 + r) * (3 * kD) + d4;
        kv = *reinterpret_cast<const float4*>(p + kD);
        vv = *reinterpret_cast<const float4*>(p + 2 * kD);
      }
      sKt[d4 + 0][r] = kv.x; sKt[d4 + 1][r] = kv.y; sKt[d4 + 2][r] = kv.z; sKt[d4 + 3][r] = kv.w;
      *reinterpret_cast<float4*>(&sV[r][d4]) = vv;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 16
    for (int d = 0; d < kDH; ++d) {
      const float4 a = *reinterpret_cast<const float4*>(&sQt[d][ty * 4]);
      const float4 c = *reinterpret_cast<const float4*>(&sKt[d][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w}, cv[4] = {c.x, c.y, c.z, c.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) s[i][j] = fmaf(av[i], cv[j], s[i][j]);
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int qi = q0 + ty * 4 + i;
      float mx = -INFINITY;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int kj = k0 + tx * 4 + j;
        bool vis = kj < len;
        if (mode == 1) vis = vis && (kj < S || kj <= qi);
        s[i][j] = vis ? s[i][j] * 0.125f : -INFINITY;
        mx = fmaxf(mx, s[i][j]);
      }
#pragma unroll
      for (int ofs = 8; ofs > 0; ofs >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, ofs));
      const float m_new = fmaxf(m_i[i], mx);
      const float corr = (m_new == -INFINITY) ? 1.f : expf(m_i[i] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float p = (s[i][j] == -INFINITY) ? 0.f : expf(s[i][j] - m_new);
        sP[ty * 4 + i][tx * 4 + j] = p;
        rs += p;
      }
#pragma unroll
      for (int ofs = 8; ofs > 0; ofs >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, ofs);
      l_i[i] = l_i[i] * corr + rs;
      m_i[i] = m_new;
#pragma unroll
      for (int j = 0; j < 4; ++j) o[i][j] *= corr;
    }
    __syncwarp();  // sP rows of this thread's ty are produced by the 16 lanes of the same half-warp
#pragma unroll 8
    for (int c = 0; c < AK; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(&sV[c][tx * 4]);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float p = sP[ty * 4 + i][c];
        o[i][0] = fmaf(p, v.x, o[i][0]); o[i][1] = fmaf(p, v.y, o[i][1]);
        o[i][2] = fmaf(p, v.z, o[i][2]); o[i][3] = fmaf(p, v.w, o[i][3]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int qi = q0 + ty * 4 + i;
    if (qi < len) {
      const float inv = 1.0f / l_i[i];
      *reinterpret_cast<float4*>(out + ((size_t)row0[b] + qi) * kD + h * kDH + tx * 4) =
          make_float4(o[i][0] * inv, o[i][1] * inv, o[i][2] * inv, o[i][3] * inv);
    }
  }
}

void launch_attn_full(const Batch& bt, const float* qkv, const int32_t* row0, int mode, int max_len, float* out,
                      cudaStream_t st) {
  if (max_len <= 0) return;
  const size_t smem = 4 * kDH * (AQ + APAD) * sizeof(float);  // 69,632 B
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaFuncSetAttribute(attn_full_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    attr_done[dev & 63] = true;
  }
  attn_full_kernel<<<dim3((max_len + AQ - 1) / AQ, kH, bt.B), 256, smem, st>>>(bt, qkv, row0, mode, out);
}

// ---------------------------------------------------------------------------------------------
template <typename KVT>
__global__ void __launch_bounds__(256) kv_scatter_kernel(Batch bt, const float* __restrict__ qkv,
                                                         KVT* __restrict__ kc, KVT* __restrict__ vc, int ctx_cap) {
  const int b = blockIdx.y, pos = blockIdx.x;
  if (pos >= bt.S[b] + bt.P[b] + 1) return;
  const int c = threadIdx.x * 4, h = c / kDH, d = c % kDH;
  const float* p = qkv + ((size_t)bt.ar_row0[b] + pos) * (3 * kD) + c;
  const float4 k = *reinterpret_cast<const float4*>(p + kD);
  const float4 v = *reinterpret_cast<const float4*>(p + 2 * kD);
  const size_t o = (((size_t)b * kH + h) * ctx_cap + pos) * kDH + d;
  store_as(kc + o, k.x); store_as(kc + o + 1, k.y); store_as(kc + o + 2, k.z); store_as(kc + o + 3, k.w);
  store_as(vc + o, v.x); store_as(vc + o + 1, v.y); store_as(vc + o + 2, v.z); store_as(vc + o + 3, v.w);
}
void launch_kv_scatter(const Batch& bt, const float* qkv, void* kcache, void* vcache, int kvt, int max_len,
                       int ctx_cap, cudaStream_t st) {
  if (max_len <= 0) return;
  dim3 g(max_len, bt.B);
  if (kvt == 0) kv_scatter_kernel<float><<<g, 256, 0, st>>>(bt, qkv, (float*)kcache, (float*)vcache, ctx_cap);
  else kv_scatter_kernel<bf16><<<g, 256, 0, st>>>(bt, qkv, (bf16*)kcache, (bf16*)vcache, ctx_cap);
}

// ---------------------------------------------------------------------------------------------
// Decode attention.  grid (nsplit, H, B), 128 threads.  The new key/value (this step's row) lives at
// position pos = S+P+T in the cache; split 0 appends it, every split whose range covers it reads it
// from qkv_step so there is no inter-CTA ordering on the append.
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ float round_to(float v);
template <> __device__ __forceinline__ float round_to<float>(float v) { return v; }
template <> __device__ __forceinline__ float round_to<bf16>(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

constexpr int DEC_THREADS = 128;
constexpr int DEC_MAX_CHUNK = 2048;

template <typename KVT>
__global__ void __launch_bounds__(DEC_THREADS) attn_decode_kernel(Batch bt, const float* __restrict__ qkv_step,
                                                                  KVT* __restrict__ kc, KVT* __restrict__ vc,
                                                                  int ctx_cap, float* __restrict__ out,
                                                                  float* __restrict__ partial,
                                                                  int32_t* __restrict__ tickets, int nsplit) {
  __shared__ __align__(16) float sq[kDH];
  __shared__ __align__(16) float sknew[kDH];
  __shared__ __align__(16) float svnew[kDH];
  __shared__ float sp[DEC_MAX_CHUNK];
  __shared__ float red[32];
  __shared__ __align__(16) float sacc[16][kDH];
  __shared__ int s_last;
  const int split = blockIdx.x, h = blockIdx.y, b = blockIdx.z, tid = threadIdx.x;
  if (bt.finished[b]) return;
  const int pos = bt.S[b] + bt.P[b] + bt.T[b];  // cache slot of this step's key
  const int n = pos + 1;
  int chunk = (n + nsplit - 1) / nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = qkv_step + (size_t)b * (3 * kD) + h * kDH;
  if (tid < kDH) {
    sq[tid] = row[tid] * 0.125f;
    sknew[tid] = row[kD + tid];
    svnew[tid] = row[2 * kD + tid];
  }
  __syncthreads();
  KVT* kb = kc + ((size_t)b * kH + h) * ctx_cap * kDH;
  KVT* vb = vc + ((size_t)b * kH + h) * ctx_cap * kDH;
  if (split == 0 && tid < kDH) {  // append (round to the cache type once; later steps read the same bits)
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  // scores: one key per thread
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += DEC_THREADS) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], round_to<KVT>(sknew[d]), s);
    } else {
      const KVT* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        load8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += DEC_THREADS) {
    const float p = expf(sp[j - j0] - m);
    sp[j - j0] = p;
    lloc += p;
  }
  const float l = block_sum(lloc, red);  // also orders the sp[] writes before the PV reads
  // PV: 16 key groups x 8 dim chunks
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = j0 + g; j < j1; j += 16) {
    const float p = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = round_to<KVT>(svnew[dc + i]);
    } else {
      load8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(p, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g][dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
#pragma unroll
    for (int q = 0; q < 16; ++q) od += sacc[q][tid];
  }
  float* dst = out + (size_t)b * kD + h * kDH;
  if (nsplit == 1) {
    if (tid < kDH) dst[tid] = od / l;
    return;
  }
  float* pp = partial + (((size_t)b * kH + h) * nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  __threadfence();
  __syncthreads();
  if (tid == 0) s_last = (atomicAdd(&tickets[b * kH + h], 1) == nsplit - 1);
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  const float* p0 = partial + ((size_t)b * kH + h) * nsplit * (kDH + 2);
  float M = -INFINITY;
  for (int s = 0; s < nsplit; ++s) M = fmaxf(M, __ldcg(p0 + s * (kDH + 2)));  // empty splits carry m = -inf, l = 0
  if (tid < kDH) {
    float L = 0.f, O = 0.f;
    for (int s = 0; s < nsplit; ++s) {
      const float ms = __ldcg(p0 + s * (kDH + 2));
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
      L += w * __ldcg(p0 + s * (kDH + 2) + 1);
      O += w * __ldcg(p0 + s * (kDH + 2) + 2 + tid);
    }
    dst[tid] = O / L;
  }
  if (tid == 0) tickets[b * kH + h] = 0;
}

void launch_attn_decode(const Batch& bt, const float* qkv_step, void* kcache, void* vcache, int kvt, int ctx_cap,
                        float* out, float* partial, int32_t* tickets, int nsplit, cudaStream_t st) {
  dim3 g(nsplit, kH, bt.B);
  if (kvt == 0)
    attn_decode_kernel<float><<<g, DEC_THREADS, 0, st>>>(bt, qkv_step, (float*)kcache, (float*)vcache, ctx_cap, out,
                                                         partial, tickets, nsplit);
  else
    attn_decode_kernel<bf16><<<g, DEC_THREADS, 0, st>>>(bt, qkv_step, (bf16*)kcache, (bf16*)vcache, ctx_cap, out,
                                                        partial, tickets, nsplit);
}

}  // namespace vx
