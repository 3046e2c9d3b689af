// bf16 tensor-core attention for the full-sequence passes (bf16 mode): flash-style online softmax,
// S = Q K^T and O = P V on mma.sync.m16n8k16 (bf16 in, fp32 accumulate), K/V tiles double-buffered with
// cp.async, shared memory XOR-swizzled for conflict-free ldmatrix.  Head dim 64.
//   mode 0: bidirectional (NAR, modules/activation.py:566-585)
//   mode 1: AR prefill mask (models/vallex.py:534-549): key j visible to query i iff j < S or j <= i.
// Input : qkv bf16 row-major [rows, 3072] (q | k | v, head h at columns h*64).
// Output: bf16 tile image [rows_pad, 1024] = the A operand of the out-proj tcgen05 GEMM.
// 12*4*L^2*1024 FLOP per pass is ~23 % of a NAR pass; the tcgen05/TMEM formulation of this
// kernel is the next step -- this one already keeps the pass tensor-bound rather than SIMT-bound.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vx {

constexpr int FQ = 64, FK = 64;

__device__ __forceinline__ uint32_t swz(int row, int chunk) { return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4)); }

__device__ __forceinline__ void cp_async16(void* smem, const void* g, bool pred) {
  const uint32_t s = tc::smem_u32(smem);
  const int sz = pred ? 16 : 0;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(g), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void ldsm_x4(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void ldsm_x4_t(uint32_t addr, uint32_t& r0, uint32_t& r1, uint32_t& r2, uint32_t& r3) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {  // one cvt.rn.bf16x2.f32 (low half = a)
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ float ex2(float x) {  // single MUFU.EX2; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

__global__ void __launch_bounds__(128, 4) attn_tc_kernel(Batch bt, const bf16* __restrict__ qkv,
                                                      const int32_t* __restrict__ row0, int mode,
                                                      uint8_t* __restrict__ out_img) {
  __shared__ __align__(1024) uint8_t sQ[FQ * 128];
  __shared__ __align__(1024) uint8_t sK[2][FK * 128];
  __shared__ __align__(1024) uint8_t sV[2][FK * 128];
  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * FQ;
  const int S = bt.S[b];
  const int len = (mode == 1) ? S + bt.P[b] + 1 : S + bt.P[b] + bt.T[b];
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const bf16* base = qkv + (size_t)row0[b] * (3 * kD) + h * kDH;
  int kend = len;
  if (mode == 1) kend = min(len, max(S, q0 + FQ));
  const int ntiles = (kend + FK - 1) / FK;

  auto load_kv = [&](int tile, int buf) {
    const int k0 = tile * FK;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int idx = tid + i * 128;  // 512 chunks per matrix
      const int r = idx >> 3, c = idx & 7;
      const bool ok = (k0 + r) < len;
      const bf16* src = base + (size_t)(ok ? (k0 + r) : 0) * (3 * kD) + c * 8;
      cp_async16(sK[buf] + swz(r, c), src + kD, ok);
      cp_async16(sV[buf] + swz(r, c), src + 2 * kD, ok);
    }
  };
  // Q tile (64 x 64 bf16)
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = tid + i * 128;
    const int r = idx >> 3, c = idx & 7;
    const bool ok = (q0 + r) < len;
    cp_async16(sQ + swz(r, c), base + (size_t)(ok ? (q0 + r) : 0) * (3 * kD) + c * 8, ok);
  }
  load_kv(0, 0);
  cp_async_commit();

  float o[8][4];
#pragma unroll
  for (int j = 0; j < 8; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) o[j][i] = 0.f;
  float m_r[2] = {-INFINITY, -INFINITY}, l_r[2] = {0.f, 0.f};
  uint32_t qf[4][4];
  const float sl2 = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
  const int qrow_lo = q0 + warp * 16 + (lane >> 2);

  for (int t = 0; t < ntiles; ++t) {
    const int buf = t & 1;
    if (t + 1 < ntiles) load_kv(t + 1, buf ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncthreads();
    if (t == 0) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {  // A fragments of this warp's 16 query rows
        const int r = warp * 16 + (lane & 15), c = kk * 2 + (lane >> 4);
        ldsm_x4(tc::smem_u32(sQ + swz(r, c)), qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3]);
      }
    }
    float s[8][4];
#pragma unroll
    for (int j = 0; j < 8; ++j)
#pragma unroll
      for (int i = 0; i < 4; ++i) s[j][i] = 0.f;
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
#pragma unroll
      for (int jp = 0; jp < 4; ++jp) {  // two key n-tiles per ldmatrix.x4
        uint32_t b0, b1, b2, b3;
        const int r = jp * 16 + (lane & 7) + ((lane >> 4) << 3), c = kk * 2 + ((lane >> 3) & 1);
        ldsm_x4(tc::smem_u32(sK[buf] + swz(r, c)), b0, b1, b2, b3);
        mma_bf16(s[2 * jp], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b0, b1);
        mma_bf16(s[2 * jp + 1], qf[kk][0], qf[kk][1], qf[kk][2], qf[kk][3], b2, b3);
      }
    }
    // mask + online softmax (rows qrow_lo and qrow_lo + 8).  Interior key tiles (every key visible to every row of this
    // warp) skip the per-element visibility logic: raw scores stay unscaled and the exp argument is one FFMA.
    const int k0 = t * FK;
    const bool full = (k0 + FK <= len) && (mode == 0 || k0 + FK <= S || k0 + FK - 1 <= q0 + warp * 16);
    float mx[2] = {-INFINITY, -INFINITY};
    if (full) {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) mx[i >> 1] = fmaxf(mx[i >> 1], s[j][i]);
      mx[0] *= sl2; mx[1] *= sl2;   // sl2 > 0: max commutes with the scale
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int kj = k0 + j * 8 + (lane & 3) * 2 + (i & 1);
          const int qi = qrow_lo + ((i >> 1) << 3);
          bool vis = kj < len;
          if (mode == 1) vis = vis && (kj < S || kj <= qi);
          s[j][i] = vis ? s[j][i] : -INFINITY;
          mx[i >> 1] = fmaxf(mx[i >> 1], s[j][i] * sl2);
        }
    }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_r[r], mx[r]);
      corr[r] = (m_new == -INFINITY) ? 1.f : ex2(m_r[r] - m_new);
      m_r[r] = m_new;
    }
    float rs[2] = {0.f, 0.f};
    uint32_t pf[4][4];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float p[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // masked scores are -inf: fmaf(-inf, sl2, -m) = -inf -> exp2f = 0 (m is finite once any key of the row was seen)
        p[i] = ex2(fmaf(s[j][i], sl2, -m_r[i >> 1]));
        rs[i >> 1] += p[i];
      }
      pf[j >> 1][(j & 1) * 2 + 0] = pack_bf16(p[0], p[1]);
      pf[j >> 1][(j & 1) * 2 + 1] = pack_bf16(p[2], p[3]);
    }
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 1);
      rs[r] += __shfl_xor_sync(0xffffffffu, rs[r], 2);
      l_r[r] = l_r[r] * corr[r] + rs[r];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      o[j][0] *= corr[0]; o[j][1] *= corr[0]; o[j][2] *= corr[1]; o[j][3] *= corr[1];
    }
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {      // keys 16kk .. 16kk+15
#pragma unroll
      for (int dp = 0; dp < 4; ++dp) {    // two d n-tiles per ldmatrix.x4.trans
        uint32_t b0, b1, b2, b3;
        const int r = kk * 16 + (lane & 7) + (((lane >> 3) & 1) << 3), c = dp * 2 + (lane >> 4);
        ldsm_x4_t(tc::smem_u32(sV[buf] + swz(r, c)), b0, b1, b2, b3);
        mma_bf16(o[2 * dp], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b0, b1);
        mma_bf16(o[2 * dp + 1], pf[kk][0], pf[kk][1], pf[kk][2], pf[kk][3], b2, b3);
      }
    }
    __syncthreads();  // everyone done with buf before it is refilled two iterations later
  }
  cp_async_wait<0>();
  // normalise, stage through sQ (this warp's 16 rows only), then 16-byte stores into the tile image
  const float inv0 = 1.0f / l_r[0], inv1 = 1.0f / l_r[1];
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r_lo = warp * 16 + (lane >> 2), c = j;       // chunk j holds d 8j..8j+7
    const int within = (lane & 3) * 4;                      // byte offset of this thread's pair inside the chunk
    *reinterpret_cast<uint32_t*>(sQ + swz(r_lo, c) + within) = pack_bf16(o[j][0] * inv0, o[j][1] * inv0);
    *reinterpret_cast<uint32_t*>(sQ + swz(r_lo + 8, c) + within) = pack_bf16(o[j][2] * inv1, o[j][3] * inv1);
  }
  __syncwarp();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int idx = lane + i * 32;  // 16 rows x 8 chunks
    const int r = warp * 16 + (idx >> 3), c = idx & 7;
    const int qi = q0 + r;
    if (qi < len) {
      const uint4 v = *reinterpret_cast<const uint4*>(sQ + swz(r, c));
      *reinterpret_cast<uint4*>(out_img + tc::tile_image_offset((int64_t)row0[b] + qi, h * kDH + c * 8, kD)) = v;
    }
  }
}

void launch_attn_tc(const Batch& bt, const void* qkv_bf16, const int32_t* row0, int mode, int max_len, void* out_img,
                    cudaStream_t st) {
  if (max_len <= 0) return;
  attn_tc_kernel<<<dim3((max_len + FQ - 1) / FQ, kH, bt.B), 128, 0, st>>>(bt, (const bf16*)qkv_bf16, row0, mode,
                                                                          (uint8_t*)out_img);
}

// bf16 qkv rows -> KV cache (bf16) for the AR prefill in bf16 mode
__global__ void __launch_bounds__(128) kv_scatter_bf16_kernel(Batch bt, const bf16* __restrict__ qkv,
                                                              bf16* __restrict__ kc, bf16* __restrict__ vc, int ctx_cap) {
  const int b = blockIdx.y, pos = blockIdx.x;
  if (pos >= bt.S[b] + bt.P[b] + 1) return;
  const int c = threadIdx.x * 8, h = c / kDH, d = c % kDH;
  const bf16* p = qkv + ((size_t)bt.ar_row0[b] + pos) * (3 * kD) + c;
  const size_t o = (((size_t)b * kH + h) * ctx_cap + pos) * kDH + d;
  *reinterpret_cast<uint4*>(kc + o) = *reinterpret_cast<const uint4*>(p + kD);
  *reinterpret_cast<uint4*>(vc + o) = *reinterpret_cast<const uint4*>(p + 2 * kD);
}
void launch_kv_scatter_bf16(const Batch& bt, const void* qkv_bf16, void* kcache, void* vcache, int max_len, int ctx_cap,
                            cudaStream_t st) {
  if (max_len <= 0) return;
  kv_scatter_bf16_kernel<<<dim3(max_len, bt.B), 128, 0, st>>>(bt, (const bf16*)qkv_bf16, (bf16*)kcache, (bf16*)vcache, ctx_cap);
}

}  // namespace vx
