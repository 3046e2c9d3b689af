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

namespace vx {
using namespace tc;

constexpr int T5_SMEM = 16384 /*Q*/ + 2 * 16384 /*K*/ + 2 * 16384 /*V^T*/ + 32768 /*P*/ + 128 /*barriers + tmem slot*/;  // x2 CTAs fit one SM

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

__global__ void __launch_bounds__(128, 2) attn_t5_kernel(Batch bt, const uint8_t* __restrict__ q_img,
                                                         const uint8_t* __restrict__ k_img,
                                                         const uint8_t* __restrict__ vt_img, int64_t R,
                                                         const int32_t* __restrict__ row0, int mode,
                                                         uint8_t* __restrict__ out_img) {
  extern __shared__ __align__(1024) uint8_t t5_raw[];
  uint8_t* smem = t5_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the 128-byte swizzle needs 1024-byte aligned operand tiles
  uint8_t* sQ = smem;
  uint8_t* sK = smem + 16384;           // [2][16 KB]
  uint8_t* sV = smem + 3 * 16384;       // [2][16 KB]  (two 8 KB k-blocks each)
  uint8_t* sP = smem + 5 * 16384;       // [2 k-blocks][128 rows x 128 B]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 7 * 16384);
  uint64_t* qbar = bars;
  uint64_t* kbar = bars + 1;            // [2]
  uint64_t* vbar = bars + 3;            // [2]
  uint64_t* mbar = bars + 5;            // MMA completion
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int S = bt.S[b];
  const int len = (mode == 1) ? S + bt.P[b] + 1 : S + bt.P[b] + bt.T[b];
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int64_t r0 = row0[b];           // multiple of 128
  int kend = len;
  if (mode == 1) kend = min(len, max(S, q0 + 128));
  const int nkt = (kend + 127) >> 7;
  const size_t head_bytes = (size_t)R * 128;
  const uint8_t* qsrc = q_img + (size_t)h * head_bytes + (size_t)((r0 + q0) >> 7) * 16384;
  const uint8_t* ksrc = k_img + (size_t)h * head_bytes + (size_t)(r0 >> 7) * 16384;           // + j * 16384
  const uint8_t* vsrc = vt_img + ((size_t)h * (R / 64) + (size_t)(r0 >> 6)) * 8192;            // + j * 16384

  if (tid == 0) {
    mbar_init(qbar, 1);
    mbar_init(&kbar[0], 1); mbar_init(&kbar[1], 1);
    mbar_init(&vbar[0], 1); mbar_init(&vbar[1], 1);
    mbar_init(mbar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t tS = tmem, tO = tmem + 128;
  const uint32_t lane_addr = ((uint32_t)(warp * 32)) << 16;  // this warp's TMEM lane quarter

  uint32_t kuse[2] = {0, 0}, vuse[2] = {0, 0}, muse = 0;  // per-barrier use counts -> wait parity (uniform across threads)
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128);
  constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64);
  const float sl2 = 0.125f * 1.44269504088896340736f;  // 1/sqrt(64) * log2(e)
  const int qi = q0 + tid;                               // this thread's query (sequence-local index)

  auto issue_k = [&](int j) {  // thread 0
    mbar_expect_tx(&kbar[j & 1], 16384);
    bulk_g2s(sK + (j & 1) * 16384, ksrc + (size_t)j * 16384, 16384, &kbar[j & 1]);
  };
  auto issue_v = [&](int j) {
    mbar_expect_tx(&vbar[j & 1], 16384);
    bulk_g2s(sV + (j & 1) * 16384, vsrc + (size_t)j * 16384, 16384, &vbar[j & 1]);
  };
  auto mma_s = [&](int j) {  // thread 0: S = Q K_j^T
    const uint64_t ad = umma_desc_sw128(smem_u32(sQ)), bd = umma_desc_sw128(smem_u32(sK + (j & 1) * 16384));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tS, ad + 2 * k, bd + 2 * k, idesc_s, k ? 1u : 0u);
    umma_commit(mbar);
  };
  auto visible = [&](int kj) {
    bool vis = kj < len;
    if (mode == 1) vis = vis && (kj < S || kj <= qi);
    return vis;
  };

  if (tid == 0) {
    mbar_expect_tx(qbar, 16384);
    bulk_g2s(sQ, qsrc, 16384, qbar);
    issue_k(0);
  }
  // ---------------- pass 1: row max and row sum -------------------------------------------------------------
  float m_run = -INFINITY, l_run = 0.f;
  mbar_wait(qbar, 0);
  for (int j = 0; j < nkt; ++j) {
    if (tid == 0 && j + 1 < nkt) issue_k(j + 1);  // the other slot was released by the MMA of tile j-1 (waited below)
    mbar_wait(&kbar[j & 1], kuse[j & 1] & 1);
    kuse[j & 1]++;
    if (tid == 0) {
      tc_fence_after();
      mma_s(j);
    }
    mbar_wait(mbar, muse & 1);
    muse++;
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(tS + lane_addr + c0, v);
      float mx = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; ++i) {
        v[i] = visible(j * 128 + c0 + i) ? v[i] * sl2 : -INFINITY;
        mx = fmaxf(mx, v[i]);
      }
      const float m_new = fmaxf(m_run, mx);
      if (m_new != -INFINITY) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 32; ++i) s += (v[i] == -INFINITY) ? 0.f : exp2f(v[i] - m_new);
        l_run = l_run * ((m_run == -INFINITY) ? 0.f : exp2f(m_run - m_new)) + s;
        m_run = m_new;
      }
    }
    tc_fence_before();
    __syncthreads();  // every row has read S: TMEM S and this K slot may be overwritten
  }
  const float inv_l = 1.0f / l_run;
  // ---------------- pass 2: P = exp2(S - m) / l, O += P V ---------------------------------------------------
  if (tid == 0) { issue_k(0); issue_v(0); }
  for (int j = 0; j < nkt; ++j) {
    if (tid == 0 && j + 1 < nkt) { issue_k(j + 1); issue_v(j + 1); }
    mbar_wait(&kbar[j & 1], kuse[j & 1] & 1);
    kuse[j & 1]++;
    if (tid == 0) {
      tc_fence_after();
      mma_s(j);
    }
    mbar_wait(mbar, muse & 1);
    muse++;
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < 128; c0 += 32) {
      float v[32];
      tmem_ld32(tS + lane_addr + c0, v);
      uint32_t pk[16];
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        const float p0 = visible(j * 128 + c0 + i) ? exp2f(v[i] * sl2 - m_run) * inv_l : 0.f;
        const float p1 = visible(j * 128 + c0 + i + 1) ? exp2f(v[i + 1] * sl2 - m_run) * inv_l : 0.f;
        pk[i >> 1] = (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(p0)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(p1)) << 16);
      }
      // row tid, keys c0..c0+31 -> k-block c0/64, chunks (c0%64)/8 .. +3, swizzled by (row & 7)
      uint8_t* prow = sP + (c0 >> 6) * 16384 + tid * 128;
      const int cb = (c0 & 63) >> 3;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        *reinterpret_cast<uint4*>(prow + (((cb + q) ^ (tid & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
    }
    fence_proxy_async();   // P was written through the generic proxy, the MMA reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      mbar_wait(&vbar[j & 1], vuse[j & 1] & 1);
      tc_fence_after();
      const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (j & 1) * 16384);
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t ad = umma_desc_sw128(pa + (ks >> 2) * 16384) + 2 * (ks & 3);
        const uint64_t bd = umma_desc_sw128(va + (ks >> 2) * 8192) + 2 * (ks & 3);
        umma_bf16(tO, ad, bd, idesc_o, (j | ks) ? 1u : 0u);
      }
      umma_commit(mbar);
    }
    vuse[j & 1]++;
    mbar_wait(mbar, muse & 1);  // P V done: sP, this V slot and TMEM S are free again, O is up to date
    muse++;
    tc_fence_after();
  }
  // ---------------- epilogue: O (already normalised) -> bf16 tile image -------------------------------------
  {
    uint8_t* orow = out_img + tile_image_offset(r0 + qi, h * kDH, kD);  // 64 dims of head h = one 128-byte row slot
    const int swz = (int)((r0 + qi) & 7);
#pragma unroll 1
    for (int c0 = 0; c0 < 64; c0 += 32) {
      float v[32];
      tmem_ld32(tO + lane_addr + c0, v);
      if (qi < len) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 p;
          auto pk2 = [](float a, float b2) {
            return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b2)) << 16);
          };
          p.x = pk2(v[8 * q], v[8 * q + 1]); p.y = pk2(v[8 * q + 2], v[8 * q + 3]);
          p.z = pk2(v[8 * q + 4], v[8 * q + 5]); p.w = pk2(v[8 * q + 6], v[8 * q + 7]);
          // tile_image_offset(row, h*64, 1024) points at chunk 0 ^ swz of the slot; rebuild the slot base and place chunk c
          const int chunk = (c0 >> 3) + q;
          uint8_t* slot = orow - ((0 ^ swz) << 4);
          *reinterpret_cast<uint4*>(slot + ((chunk ^ swz) << 4)) = p;
        }
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
    cudaFuncSetAttribute(attn_t5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_SMEM);
    attr_done[dev & 63] = true;
  }
  attn_t5_kernel<<<dim3((max_len + 127) / 128, kH, bt.B), 128, T5_SMEM, st>>>(bt, (const uint8_t*)q_img, (const uint8_t*)k_img,
                                                                             (const uint8_t*)vt_img, R, row0, mode,
                                                                             (uint8_t*)out_img);
}

}  // namespace vx
