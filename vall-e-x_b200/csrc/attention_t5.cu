// tcgen05 / TMEM attention for the full-sequence passes (bf16 mode): S = Q K^T and O = P V on the 5th-generation
// tensor cores, accumulators in TMEM, operands staged by the TMA engine (cp.async.bulk) as tile images.
//
// One CTA = 128 query rows of one (utterance, head); 128 threads, thread t owns query row t (= TMEM lane t), so the
// softmax statistics are per-thread registers -- no shuffles, no shared-memory reductions.  Two passes over the key
// tiles (128 keys each):
//   pass 1: S_j = Q K_j^T (4 x tcgen05.mma M128 N128 K16) -> tcgen05.ld -> running row max m and row sum l;
//   pass 2: S_j again, P_j = exp2(S_j - m) / l written as bf16 into shared memory in the K-major 128B-swizzled image a
//           tcgen05 A operand wants, then O += P_j V_j (8 x tcgen05.mma M128 N64 K16).  Because m and l are final,
//           O never needs the online-softmax rescale (no TMEM read-modify-write); the price is a second QK^T, which
//           the tensor pipe does in ~260 cycles per tile while the exponentials (SFU) are the real bound.
// Two CTAs are resident per SM (112 KB shared memory, 256 TMEM columns each): while one runs its softmax the other
// one's MMAs and copies proceed, so the synchronous in-CTA flow needs no warp specialisation.
//   mode 0: bidirectional (NAR, modules/activation.py:566-585)
//   mode 1: AR prefill mask (models/vallex.py:534-549): key j visible to query i iff j < S or j <= i.
// Operands (built by pack_qkv_heads_kernel from the QKV GEMM's row-major bf16 output; utterances start on 128-row
// boundaries so a key tile is exactly one image block):
//   q_img, k_img : per head, tile image of [rows, 64]     (block = 128 rows x 64 d, 16 KB)
//   vt_img       : per head, per 64-key block, [64 d rows x 128 B] K-major swizzled (V transposed: B operand of P V)
// Output: bf16 tile image [rows, 1024] = the A operand of the out-proj GEMM.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

#include <cstdlib>

namespace vx {
using namespace tc;


// row-major bf16 qkv [R,3072] -> per-head operand images (R % 128 == 0)
__global__ void __launch_bounds__(256) pack_qkv_heads_kernel(const bf16* __restrict__ qkv, int64_t R,
                                                             uint8_t* __restrict__ q_img, uint8_t* __restrict__ k_img,
                                                             uint8_t* __restrict__ vt_img) {
  const int64_t r = (int64_t)blockIdx.x * 2 + (threadIdx.x >> 7);  // 2 rows per CTA, 128 threads per row
  if (r >= R) return;
  const int t = threadIdx.x & 127;   // 128 chunks of 8 elements = 1024 columns
  const int h = t >> 3, c = t & 7;   // head, 16-byte chunk inside the head's 64 dims
  const bf16* row = qkv + r * (3 * kD);
  const uint4 q = *reinterpret_cast<const uint4*>(row + t * 8);
  const uint4 k = *reinterpret_cast<const uint4*>(row + kD + t * 8);
  const uint4 v = *reinterpret_cast<const uint4*>(row + 2 * kD + t * 8);
  const size_t head_bytes = (size_t)R * 128;  // [R, 64] bf16
  const size_t off = (size_t)h * head_bytes + tile_image_offset(r, c * 8, 64);
  *reinterpret_cast<uint4*>(q_img + off) = q;
  *reinterpret_cast<uint4*>(k_img + off) = k;
  // V^T: element (d, key r) -> block (h, r/64), row d, chunk ((r%64)/8) ^ (d&7), lane r%8
  const uint16_t* ve = reinterpret_cast<const uint16_t*>(&v);
  uint8_t* vb = vt_img + ((size_t)h * (R / 64) + (r >> 6)) * 8192;
  const int kc = (int)((r & 63) >> 3), kl = (int)(r & 7);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int d = c * 8 + i;
    *reinterpret_cast<uint16_t*>(vb + d * 128 + ((kc ^ (d & 7)) << 4) + kl * 2) = ve[i];
  }
}
void launch_pack_qkv_heads(const void* qkv_bf16, int64_t R, void* q_img, void* k_img, void* vt_img, cudaStream_t st) {
  if (R <= 0) return;
  pack_qkv_heads_kernel<<<(unsigned)((R + 1) / 2), 256, 0, st>>>((const bf16*)qkv_bf16, R, (uint8_t*)q_img, (uint8_t*)k_img,
                                                               (uint8_t*)vt_img);
}

// shared memory: Q 16 KB | K ring 3 x 8 KB | V^T ring 4 x 8 KB | P 2 x 16 KB | row stats 2 KB | barriers
constexpr int T5_Q = 0, T5_K = 16384, T5_V = T5_K + 3 * 8192, T5_P = T5_V + 4 * 8192, T5_ST = T5_P + 2 * 16384,
              T5_BAR = T5_ST + 2048, T5_BYTES = T5_BAR + 128;

__device__ __forceinline__ float t5_ex2(float x) {  // single MUFU.EX2; ex2(-inf) = 0
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t t5_pack(float a, float b) {  // cvt.rn.bf16x2.f32, low half = a
  uint32_t r;
  asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}

// 256 threads: warps 0-3 own rows 0..127 x key columns [0,32) of every 64-key tile, warps 4-7 the same rows x [32,64).
// S is double-buffered in TMEM so the MMA of tile j+1 runs under the softmax of tile j; P is double-buffered in shared
// memory so O += P_j V_j runs under the softmax of tile j+1.
// MAXONLY (VX_ATTN_T5=2, experimental): pass 1 tracks only the row maximum (no exponentials), pass 2 accumulates the row sum
// while it forms P = exp2(S - m) and the epilogue scales O by 1/l -- half the MUFU work of the statistics pass.
template <bool MAXONLY>
__global__ void __launch_bounds__(256, 2) attn_t5_kernel(Batch bt, const uint8_t* __restrict__ q_img,
                                                         const uint8_t* __restrict__ k_img,
                                                         const uint8_t* __restrict__ vt_img, int64_t R,
                                                         const int32_t* __restrict__ row0, int mode,
                                                         uint8_t* __restrict__ out_img) {
  extern __shared__ __align__(1024) uint8_t t5_raw[];
  uint8_t* smem = t5_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128-byte swizzle needs 1024-byte aligned operand tiles
  uint8_t* sQ = smem + T5_Q;
  uint8_t* sK = smem + T5_K;
  uint8_t* sV = smem + T5_V;
  uint8_t* sP = smem + T5_P;
  float2* sstat = reinterpret_cast<float2*>(smem + T5_ST);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T5_BAR);
  uint64_t* qbar = bars;         // Q landed
  uint64_t* kfull = bars + 1;    // [3]
  uint64_t* vfull = bars + 4;    // [4]
  uint64_t* sfull = bars + 8;    // [2]  S_j in TMEM
  uint64_t* pvdone = bars + 10;  // [2]  O += P_j V_j finished (P buffer and V slot reusable)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int S = bt.S[b];
  const int len = (mode == 1) ? S + bt.P[b] + 1 : S + bt.P[b] + bt.T[b];
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;                 // which 32 of the tile's 64 key columns
  const int rloc = (warp & 3) * 32 + lane;    // query row inside the tile = TMEM lane
  const int64_t r0 = row0[b];                 // multiple of 128
  int kend = len;
  if (mode == 1) kend = min(len, max(S, q0 + 128));
  const int nkt = (kend + 63) >> 6;           // 64-key tiles
  const size_t head_bytes = (size_t)R * 128;
  const uint8_t* qsrc = q_img + (size_t)h * head_bytes + (size_t)((r0 + q0) >> 7) * 16384;
  const uint8_t* ksrc = k_img + (size_t)h * head_bytes + (size_t)(r0 >> 7) * 16384;   // + j * 8192 (64 rows x 128 B)
  const uint8_t* vsrc = vt_img + ((size_t)h * (R / 64) + (size_t)(r0 >> 6)) * 8192;    // + j * 8192

  if (tid == 0) {
    mbar_init(qbar, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&kfull[i], 1);
    for (int i = 0; i < 4; ++i) mbar_init(&vfull[i], 1);
    for (int i = 0; i < 2; ++i) { mbar_init(&sfull[i], 1); mbar_init(&pvdone[i], 1); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tO = tmem + 128;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;

  // barrier phase counters (every thread performs every wait, so these stay uniform)
  uint32_t kph = 0, vph = 0, sph = 0, pph = 0;  // bit i = parity of barrier i of the group
  auto wait_k = [&](int slot) { mbar_wait(&kfull[slot], (kph >> slot) & 1); kph ^= 1u << slot; };
  auto wait_s = [&](int buf) { mbar_wait(&sfull[buf], (sph >> buf) & 1); sph ^= 1u << buf; };
  auto wait_pv = [&](int buf) { mbar_wait(&pvdone[buf], (pph >> buf) & 1); pph ^= 1u << buf; };

  constexpr uint32_t idesc = umma_idesc_bf16(128, 64);
  const float sl2 = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
  const int qi = q0 + rloc;

  auto issue_k = [&](int j) {  // thread 0
    mbar_expect_tx(&kfull[j % 3], 8192);
    bulk_g2s(sK + (j % 3) * 8192, ksrc + (size_t)j * 8192, 8192, &kfull[j % 3]);
  };
  auto issue_v = [&](int j) {
    mbar_expect_tx(&vfull[j & 3], 8192);
    bulk_g2s(sV + (j & 3) * 8192, vsrc + (size_t)j * 8192, 8192, &vfull[j & 3]);
  };
  auto mma_s = [&](int j) {  // thread 0: S_j = Q K_j^T -> TMEM buffer j & 1
    tc_fence_after();
    const uint64_t ad = umma_desc_sw128(smem_u32(sQ)), bd = umma_desc_sw128(smem_u32(sK + (j % 3) * 8192));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem + (j & 1) * 64, ad + 2 * k, bd + 2 * k, idesc, k ? 1u : 0u);
    umma_commit(&sfull[j & 1]);
  };
  auto tile_full = [&](int j) {  // every key of tile j visible to every row of this CTA
    const int k0 = j * 64;
    return (k0 + 64 <= len) && (mode == 0 || k0 + 64 <= S || k0 + 63 <= q0);
  };
  auto visible = [&](int kj) {
    bool vis = kj < len;
    if (mode == 1) vis = vis && (kj < S || kj <= qi);
    return vis;
  };

  if (tid == 0) {
    mbar_expect_tx(qbar, 16384);
    bulk_g2s(sQ, qsrc, 16384, qbar);
    for (int j = 0; j < 3 && j < nkt; ++j) issue_k(j);
  }
  // ---------------- pass 1: row max and row sum of this thread's column half -----------------------------------
  float m_run = -INFINITY, l_run = 0.f;
  mbar_wait(qbar, 0);
  wait_k(0);
  if (tid == 0) mma_s(0);
  for (int j = 0; j < nkt; ++j) {
    if (j + 1 < nkt) {
      wait_k((j + 1) % 3);
      if (tid == 0) mma_s(j + 1);       // runs under the softmax of tile j
    }
    wait_s(j & 1);
    if (tid == 0 && j + 3 < nkt) issue_k(j + 3);  // S_j is complete: its K slot is free
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem + (j & 1) * 64 + lane_addr + half * 32, v);
    const int kb = j * 64 + half * 32;
    float mx = -INFINITY;
    if (tile_full(j)) {
#pragma unroll
      for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
    } else {
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = visible(kb + i) ? v[i] : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
    }
    const float m_new = fmaxf(m_run, mx * sl2);
    if constexpr (MAXONLY) {
      m_run = m_new;
    } else if (m_new != -INFINITY) {
      float s = 0.f;
#pragma unroll
      for (int i = 0; i < 32; ++i) s += t5_ex2(fmaf(v[i], sl2, -m_new));
      l_run = l_run * t5_ex2(m_run - m_new) + s;   // ex2(-inf) = 0 covers the first tile
      m_run = m_new;
    }
    tc_fence_before();
    __syncthreads();  // both halves have read S_j: its TMEM buffer may take S_{j+2}
  }
  // combine the two column halves of every row
  sstat[tid] = make_float2(m_run, l_run);
  __syncthreads();
  {
    const float2 o = sstat[tid ^ 128];
    const float m = fmaxf(m_run, o.x);
    const float l = (m == -INFINITY) ? 1.f : l_run * t5_ex2(m_run - m) + o.y * t5_ex2(o.x - m);
    m_run = m;
    l_run = l;
  }
  const float inv_l = MAXONLY ? 1.0f : 1.0f / l_run;
  float l_acc = 0.f;   // MAXONLY: row sum of this thread's column half, accumulated in pass 2
  // ---------------- pass 2: P = exp2(S - m) / l, O += P V ---------------------------------------------------------
  if (tid == 0) {
    for (int j = 0; j < 3 && j < nkt; ++j) issue_k(j);
    for (int j = 0; j < 2 && j < nkt; ++j) issue_v(j);
  }
  wait_k(0);
  if (tid == 0) mma_s(0);
  for (int j = 0; j < nkt; ++j) {
    if (j + 1 < nkt) {
      wait_k((j + 1) % 3);
      if (tid == 0) mma_s(j + 1);
    }
    wait_s(j & 1);
    if (tid == 0 && j + 3 < nkt) issue_k(j + 3);
    tc_fence_after();
    float v[32];
    tmem_ld32(tmem + (j & 1) * 64 + lane_addr + half * 32, v);
    const int kb = j * 64 + half * 32;
    uint32_t pk[16];
    if (tile_full(j)) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = t5_ex2(fmaf(v[i], sl2, -m_run)) * inv_l, p1 = t5_ex2(fmaf(v[i + 1], sl2, -m_run)) * inv_l;
        if constexpr (MAXONLY) l_acc += p0 + p1;
        pk[i >> 1] = t5_pack(p0, p1);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = visible(kb + i) ? t5_ex2(fmaf(v[i], sl2, -m_run)) * inv_l : 0.f;
        const float p1 = visible(kb + i + 1) ? t5_ex2(fmaf(v[i + 1], sl2, -m_run)) * inv_l : 0.f;
        if constexpr (MAXONLY) l_acc += p0 + p1;
        pk[i >> 1] = t5_pack(p0, p1);
      }
    }
    if (j >= 2) wait_pv(j & 1);                    // O += P_{j-2} V_{j-2} done: P buffer j&1 and V slot (j-2)&3 are free
    if (tid == 0 && j + 2 < nkt) issue_v(j + 2);
    // row rloc, keys half*32 .. +31 of the tile -> chunks half*4 .. +3 of the 128-byte row, swizzled by (row & 7)
    uint8_t* prow = sP + (j & 1) * 16384 + rloc * 128;
#pragma unroll
    for (int q = 0; q < 4; ++q)
      *reinterpret_cast<uint4*>(prow + (((half * 4 + q) ^ (rloc & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    fence_proxy_async();   // P was written through the generic proxy, the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();       // also: both halves have read S_j
    if (tid == 0) {
      mbar_wait(&vfull[j & 3], (vph >> (j & 3)) & 1);
      tc_fence_after();
      const uint64_t ad = umma_desc_sw128(smem_u32(sP + (j & 1) * 16384)), bd = umma_desc_sw128(smem_u32(sV + (j & 3) * 8192));
#pragma unroll
      for (int ks = 0; ks < 4; ++ks) umma_bf16(tO, ad + 2 * ks, bd + 2 * ks, idesc, (j | ks) ? 1u : 0u);
      umma_commit(&pvdone[j & 1]);
    }
    vph ^= 1u << (j & 3);  // every thread tracks the V phases (only thread 0 waits)
  }
  for (int j = max(nkt - 2, 0); j < nkt; ++j) wait_pv(j & 1);  // the last two P V products
  tc_fence_after();
  // ---------------- epilogue: O (already normalised unless MAXONLY) -> bf16 tile image ------------------------------------
  float o_scale = 1.0f;
  if constexpr (MAXONLY) {   // row sum = the two column halves together
    __syncthreads();
    sstat[tid] = make_float2(l_acc, 0.f);
    __syncthreads();
    const float l = l_acc + sstat[tid ^ 128].x;
    o_scale = (l > 0.f) ? 1.0f / l : 0.f;
  }
  {
    float v[32];
    tmem_ld32(tO + lane_addr + half * 32, v);
    if constexpr (MAXONLY) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= o_scale;
    }
    if (qi < len) {
      const int64_t grow = r0 + qi;
      uint8_t* slot = out_img + tile_image_offset(grow, h * kDH, kD) - (((int)(grow & 7)) << 4);  // start of the 128-byte row slot
      const int swz = (int)(grow & 7);
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 pv4;
        pv4.x = t5_pack(v[8 * q], v[8 * q + 1]); pv4.y = t5_pack(v[8 * q + 2], v[8 * q + 3]);
        pv4.z = t5_pack(v[8 * q + 4], v[8 * q + 5]); pv4.w = t5_pack(v[8 * q + 6], v[8 * q + 7]);
        *reinterpret_cast<uint4*>(slot + (((half * 4 + q) ^ swz) << 4)) = pv4;
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<256>(tmem);
}

void launch_attn_t5(const Batch& bt, const void* q_img, const void* k_img, const void* vt_img, int64_t R,
                    const int32_t* row0, int mode, int max_len, void* out_img, cudaStream_t st) {
  if (max_len <= 0) return;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaFuncSetAttribute(attn_t5_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_BYTES);
    cudaFuncSetAttribute(attn_t5_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_BYTES);
    attr_done[dev & 63] = true;
  }
  const char* env = getenv("VX_ATTN_T5");
  if (env && (env[0] == '2' || env[0] == '3')) {
    attn_t5_kernel<true><<<dim3((max_len + 127) / 128, kH, bt.B), 256, T5_BYTES, st>>>(bt, (const uint8_t*)q_img, (const uint8_t*)k_img,
                                                                             (const uint8_t*)vt_img, R, row0, mode,
                                                                             (uint8_t*)out_img);
  } else {
    attn_t5_kernel<false><<<dim3((max_len + 127) / 128, kH, bt.B), 256, T5_BYTES, st>>>(bt, (const uint8_t*)q_img, (const uint8_t*)k_img,
                                                                             (const uint8_t*)vt_img, R, row0, mode,
                                                                             (uint8_t*)out_img);
  }
}

}  // namespace vx
