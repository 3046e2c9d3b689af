// tcgen05 / TMEM attention for the full-sequence passes (bf16 mode; default, VX_ATTN_T5=0 selects the mma.sync kernel of
// attention_tc.cu): S = Q K^T and O = P V on the 5th-generation tensor cores, accumulators in TMEM, operands staged by the
// TMA engine (cp.async.bulk) as tile images.
//
// One CTA = 128 query rows of one (utterance, head); 288 threads = 8 softmax warps + 1 controller warp (one lane issues every bulk
// copy and MMA): warps 0-3 own query rows 0..127 (= TMEM lanes) x key columns
// [0,32) of every 64-key tile, warps 4-7 the same rows x [32,64), so the softmax statistics are per-thread registers -- no
// shuffles, no shared-memory reductions.  ONE sweep over the key tiles (single-pass online softmax, see the kernel comment);
// S is double-buffered in TMEM so Q K^T of tile j+1 runs under the softmax of tile j, P is double-buffered in shared memory
// (written as the K-major 128B-swizzled image a tcgen05 A operand wants) so O += P_j V_j runs under the softmax of tile j+1.
// Two CTAs are resident per SM (108 KB shared memory, 256 TMEM columns each).
//   mode 0: bidirectional (NAR, modules/activation.py:566-585)
//   mode 1: AR prefill mask (models/vallex.py:534-549): key j visible to query i iff j < S or j <= i.
// Operands (written by the QKV GEMM's TC_EPI_QKV_HEADS epilogue for the NAR layers, by pack_qkv_heads_kernel from the row-major
// bf16 QKV for the prefill layers, whose KV scatter wants the row-major form; utterances start on 128-row boundaries so a key
// tile is exactly one image block):
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

// shared memory: Q 16 KB | K ring 2 x 16 KB | V^T ring 2 x 16 KB | P 32 KB | barriers  (2 CTAs per SM)
constexpr int T5_Q = 0, T5_K = 16384, T5_V = T5_K + 2 * 16384, T5_P = T5_V + 2 * 16384, T5_ST = T5_P /* the epilogue's row statistics reuse the P buffer */,
              T5_BAR = T5_P + 32768, T5_BYTES = T5_BAR + 256;   // 114,944 B: two CTAs per SM

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

// packed fp32 pairs (sm_100: FFMA2 / FADD2 take two fp32 lanes of a 64-bit register per instruction)
__device__ __forceinline__ uint64_t t5_pack2(float lo, float hi) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void t5_unpack2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t t5_fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ uint64_t t5_add2(uint64_t a, uint64_t b) {
  uint64_t r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}

// 32 lanes x 32 consecutive fp32 columns <- 32 registers per thread (the inverse of tmem_ld32)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
      "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
      "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
      "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
      "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
      "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
      "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
      "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
      : "memory");
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}

// =================================================================================================================
// Single pass over 128-KEY tiles: one sweep with an online softmax whose running row maximum is only raised LAZILY -- when a
// tile's row maximum exceeds the maximum in use by more than 2^8 -- so the rescale of the O accumulator (TMEM read-modify-write by
// the warp that owns the rows, after the P V products in flight have drained) is rare after the first tiles; P = exp2(S - m_used)
// stays <= 256 and the row sum l carries the same scale, so the final O / l is exact softmax.  The two column halves of a row
// (keys 0-63 / 64-127 of every tile) are INDEPENDENT online softmaxes with their own accumulators O_a / O_b in TMEM (the P V
// product of a tile is issued as two K-halves), so no per-tile exchange of row maxima is needed; the halves are merged once in
// the epilogue.  A thread reads its 64 scores twice from TMEM (maximum, then 2 x 32 exponentials) instead of holding 64
// registers; S is single-buffered: its TMEM columns are free as soon as every warp has its last 32 scores in registers, so
// Q K^T of tile j+1 runs under the exponentials of tile j.
// History: two-pass kernel (round 1, tied mma.sync) -> single pass, 64-key tiles (-9 % NAR) -> no block barrier, controller warp
// (-12 %) -> 128-key tiles (this version): the per-tile synchronisation / address / polling overhead was ~40 % of the 340
// instructions a warp issued per 64-key tile (profiles/r02_prof_attn_b8.md).
// =================================================================================================================
__global__ void __launch_bounds__(288, 2) attn_t5_kernel(Batch bt, const uint8_t* __restrict__ q_img,
                                                                 const uint8_t* __restrict__ k_img,
                                                                 const uint8_t* __restrict__ vt_img, int64_t R,
                                                                 const int32_t* __restrict__ row0, int mode,
                                                                 uint8_t* __restrict__ out_img) {
  extern __shared__ __align__(1024) uint8_t t5_raw[];
  uint8_t* smem = t5_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  uint8_t* sQ = smem + T5_Q;
  uint8_t* sK = smem + T5_K;
  uint8_t* sV = smem + T5_V;
  uint8_t* sP = smem + T5_P;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + T5_BAR);
  uint64_t* qbar = bars;
  uint64_t* kfull = bars + 1;    // [2]
  uint64_t* vfull = bars + 3;    // [2]
  uint64_t* sfull = bars + 5;    // [2]  S_j complete in TMEM (alternating by tile parity; one TMEM buffer)
  uint64_t* pvdone = bars + 7;   // [2]  O += P_j V_j finished
  uint64_t* pfull = bars + 9;    // [2]  all 8 softmax warps have written their part of P_j
  uint64_t* sread = bars + 11;   // [2]  all 8 softmax warps hold the last of S_j in registers: the TMEM buffer may be overwritten
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 13);

  const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * 128;
  const int S = bt.S[b];
  const int len = (mode == 1) ? S + bt.P[b] + 1 : S + bt.P[b] + bt.T[b];
  if (q0 >= len) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int half = warp >> 2;
  const int rloc = (warp & 3) * 32 + lane;
  const int64_t r0 = row0[b];
  int kend = len;
  if (mode == 1) kend = min(len, max(S, q0 + 128));
  const int nkt = (kend + 127) >> 7;
  const size_t head_bytes = (size_t)R * 128;
  const uint8_t* qsrc = q_img + (size_t)h * head_bytes + (size_t)((r0 + q0) >> 7) * 16384;
  const uint8_t* ksrc = k_img + (size_t)h * head_bytes + (size_t)(r0 >> 7) * 16384;            // 128-key block = 16 KB
  const uint8_t* vsrc = vt_img + ((size_t)h * (R / 64) + (size_t)(r0 >> 6)) * 8192;           // two 64-key blocks = 16 KB

  if (tid == 0) {
    mbar_init(qbar, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kfull[i], 1); mbar_init(&vfull[i], 1); mbar_init(&sfull[i], 1); mbar_init(&pvdone[i], 1);
      mbar_init(&pfull[i], 8); mbar_init(&sread[i], 8);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<256>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;                // S at columns 0..127
  const uint32_t tO = tmem + 128 + half * 64;      // this column half's own accumulator: O_a at columns 128..191, O_b at 192..255
  const uint32_t tOa = tmem + 128, tOb = tmem + 192;
  const uint32_t lane_addr = ((uint32_t)((warp & 3) * 32)) << 16;

  uint32_t kph = 0, vph = 0, sph = 0, pph = 0;
  int pv_waited = -1;   // P V products of tiles <= pv_waited are known complete (every thread waits each completion once, in order)
  auto wait_k = [&](int slot) { mbar_wait(&kfull[slot], (kph >> slot) & 1); kph ^= 1u << slot; };
  auto wait_s = [&](int buf) { mbar_wait(&sfull[buf], (sph >> buf) & 1); sph ^= 1u << buf; };
  auto ensure_pv = [&](int t) {
    while (pv_waited < t) {
      ++pv_waited;
      const int buf = pv_waited & 1;
      mbar_wait(&pvdone[buf], (pph >> buf) & 1);
      pph ^= 1u << buf;
    }
  };
  constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128), idesc_pv = umma_idesc_bf16(128, 64);
  const float sl2 = 0.125f * 1.44269504088896340736f;
  const int qi = q0 + rloc;
  auto issue_k = [&](int j) {
    mbar_expect_tx(&kfull[j & 1], 16384);
    bulk_g2s(sK + (j & 1) * 16384, ksrc + (size_t)j * 16384, 16384, &kfull[j & 1]);
  };
  auto issue_v = [&](int j) {
    mbar_expect_tx(&vfull[j & 1], 16384);
    bulk_g2s(sV + (j & 1) * 16384, vsrc + (size_t)j * 16384, 16384, &vfull[j & 1]);
  };
  auto mma_s = [&](int j) {
    tc_fence_after();
    const uint64_t ad = umma_desc_sw128(smem_u32(sQ)), bd = umma_desc_sw128(smem_u32(sK + (j & 1) * 16384));
#pragma unroll
    for (int k = 0; k < 4; ++k) umma_bf16(tmem, ad + 2 * k, bd + 2 * k, idesc_s, k ? 1u : 0u);
    umma_commit(&sfull[j & 1]);
  };
  auto tile_full = [&](int j) {
    const int k0 = j * 128;
    return (k0 + 128 <= len) && (mode == 0 || k0 + 128 <= S || k0 + 127 <= q0);
  };
  auto visible = [&](int kj) {
    bool vis = kj < len;
    if (mode == 1) vis = vis && (kj < S || kj <= qi);
    return vis;
  };

  float m_used = -INFINITY;   // the row maximum this half's exponentials use
  float l_part = 0.f;         // this half's row sum, in the scale of m_used
  // No block-wide barrier inside the loop: a softmax thread (warps 0-7) owns its (row, column half) and meets the others only
  // through mbarriers -- sfull (S_j in TMEM), sread (S_j consumed), pfull (P_j complete), pvdone (O += P_j V_j finished).  Warp 8
  // is the controller: one lane issues every bulk copy and every MMA, so no softmax warp lags behind doing that work.
  if (warp == 8) {
    if (lane == 0) {
      uint32_t pfph = 0, srph = 0;
      mbar_expect_tx(qbar, 16384);
      bulk_g2s(sQ, qsrc, 16384, qbar);
      for (int j = 0; j < 2 && j < nkt; ++j) { issue_k(j); issue_v(j); }
      mbar_wait(qbar, 0);
      wait_k(0);
      mma_s(0);
      for (int j = 0; j < nkt; ++j) {
        // Q K^T of tile j+1 goes out as soon as every softmax warp holds the last of S_j in registers
        mbar_wait(&sread[j & 1], (srph >> (j & 1)) & 1);
        srph ^= 1u << (j & 1);
        if (j + 2 < nkt) issue_k(j + 2);                 // S_j was read, so its MMA is complete: K slot j & 1 is free
        if (j + 1 < nkt) { wait_k((j + 1) & 1); mma_s(j + 1); }
        if (j >= 1) {                                    // O += P_{j-1} V_{j-1} done: V slot (j+1) & 1 is free
          ensure_pv(j - 1);
          if (j + 1 < nkt) issue_v(j + 1);
        }
        mbar_wait(&pfull[j & 1], (pfph >> (j & 1)) & 1);   // every warp's part of P_j is in place, rescales are done
        pfph ^= 1u << (j & 1);
        mbar_wait(&vfull[j & 1], (vph >> (j & 1)) & 1);
        vph ^= 1u << (j & 1);
        tc_fence_after();
        const uint32_t pa = smem_u32(sP), va = smem_u32(sV + (j & 1) * 16384);
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {   // keys 0-63 of the tile (P block 0, V^T block 0) accumulate into O_a, keys 64-127 into O_b
          const uint64_t ad = umma_desc_sw128(pa + (ks >> 2) * 16384) + 2 * (ks & 3);
          const uint64_t bd = umma_desc_sw128(va + (ks >> 2) * 8192) + 2 * (ks & 3);
          umma_bf16(ks < 4 ? tOa : tOb, ad, bd, idesc_pv, (j | (ks & 3)) ? 1u : 0u);
        }
        umma_commit(&pvdone[j & 1]);
      }
    }
  } else {
    const uint32_t tS = tmem + lane_addr + half * 64;
    for (int j = 0; j < nkt; ++j) {
      wait_s(j & 1);
      tc_fence_after();
      const int kb = j * 128 + half * 64;
      const bool full = tile_full(j);
      // pass 1: the row maximum over this half's 64 scores; the FIRST 32 stay in registers for pass 2 (one TMEM round trip less)
      float mx = -INFINITY;
      float v0[32];
      {
        float v[32];
        tmem_ld32(tS + 32, v);
        if (full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v[i]);
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) mx = fmaxf(mx, visible(kb + 32 + i) ? v[i] : -INFINITY);
        }
        tmem_ld32(tS, v0);
        if (!full) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v0[i] = visible(kb + i) ? v0[i] : -INFINITY;
        }
#pragma unroll
        for (int i = 0; i < 32; ++i) mx = fmaxf(mx, v0[i]);
      }
      const float m_tile = mx * sl2;
      const bool need = m_tile > m_used + 8.0f;     // -inf + 8 = -inf: the first visible key always raises the maximum
      if (__any_sync(0xffffffffu, need)) {
        const float scale = need ? t5_ex2(m_used - m_tile) : 1.0f;   // ex2(-inf) = 0 on the first raise
        if (need) m_used = m_tile;
        l_part *= scale;
        if (j > 0) {                      // this warp owns its rows of its half's accumulator: O_h *= scale once P V products have drained
          ensure_pv(j - 1);
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 64; c += 32) {
            float o[32];
            tmem_ld32(tO + lane_addr + c, o);
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] *= scale;
            tmem_st32(tO + lane_addr + c, o);
          }
          tc_fence_before();
        }
      }
      const float mu = (m_used == -INFINITY) ? 0.f : m_used;   // a half that has seen no visible key yet: every P is ex2(-inf) = 0
      // packed fp32 pairs (FFMA2 / FADD2): half the scale-and-shift and row-sum instructions; two independent sum chains
      const uint64_t sl2x2 = t5_pack2(sl2, sl2), nmux2 = t5_pack2(-mu, -mu);
      uint64_t ls_a = 0ull, ls_b = 0ull;    // (+0, +0)
      uint8_t* prow = sP + half * 16384 + rloc * 128;      // this half's 64 keys = P block `half`, row rloc
      // pass 2: exponentials, 32 scores at a time (the first 32 are already in registers)
#pragma unroll
      for (int sub = 0; sub < 2; ++sub) {
        float v[32];
        if (sub == 0) {
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = v0[i];
        } else {
          tmem_ld32(tS + 32, v);
          tc_fence_before();                // S_j is in registers: let the controller overwrite the TMEM buffer with S_{j+1}
          __syncwarp();
          if (lane == 0) mbar_arrive(&sread[j & 1]);
          if (!full) {
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = visible(kb + 32 + i) ? v[i] : -INFINITY;
          }
        }
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
          float a0, a1, b0, b1;
          t5_unpack2(t5_fma2(t5_pack2(v[i], v[i + 1]), sl2x2, nmux2), a0, a1);
          t5_unpack2(t5_fma2(t5_pack2(v[i + 2], v[i + 3]), sl2x2, nmux2), b0, b1);
          const float p0 = t5_ex2(a0), p1 = t5_ex2(a1), p2 = t5_ex2(b0), p3 = t5_ex2(b1);   // masked: ex2(-inf) = 0
          ls_a = t5_add2(ls_a, t5_pack2(p0, p1));
          ls_b = t5_add2(ls_b, t5_pack2(p2, p3));
          pk[i >> 1] = t5_pack(p0, p1);
          pk[(i >> 1) + 1] = t5_pack(p2, p3);
        }
        if (sub == 0) ensure_pv(j - 1);     // O += P_{j-1} V_{j-1} done: the P buffer is free
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(prow + (((sub * 4 + q) ^ (rloc & 7)) << 4)) = make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
      }
      {
        float s0, s1, s2, s3;
        t5_unpack2(ls_a, s0, s1);
        t5_unpack2(ls_b, s2, s3);
        l_part += (s0 + s1) + (s2 + s3);
      }
      fence_proxy_async();                // P was written through the generic proxy, the MMA reads it through the async proxy
      tc_fence_before();                  // ... and this thread's TMEM writes of O_h (rescale) are ordered before the arrival
      __syncwarp();
      if (lane == 0) mbar_arrive(&pfull[j & 1]);
    }
    ensure_pv(nkt - 1);
  }
  tc_fence_after();
  // ---------------- epilogue: O / l -> bf16 tile image ---------------------------------------------------------------------
  float2* sml = reinterpret_cast<float2*>(smem + T5_ST);    // [256] (m_used, l_part) of every (row, half)
  if (warp < 8) sml[tid] = make_float2(m_used, l_part);
  __syncthreads();
  if (warp < 8) {
  const float2 other = sml[tid ^ 128];
  const float M = fmaxf(m_used, other.x);
  const float s_self = (m_used == -INFINITY) ? 0.f : t5_ex2(m_used - M), s_other = (other.x == -INFINITY) ? 0.f : t5_ex2(other.x - M);
  const float l = l_part * s_self + other.y * s_other;
  const float inv = (l > 0.f) ? 1.0f / l : 0.f;
  const float sa = (half == 0 ? s_self : s_other) * inv, sb = (half == 0 ? s_other : s_self) * inv;
  {
    float v[32], w[32];
    tmem_ld32(tOa + lane_addr + half * 32, v);   // output columns half*32 .. +31 of this row from BOTH accumulators
    tmem_ld32(tOb + lane_addr + half * 32, w);
    // (a half that never saw a visible key has scale 0 and an accumulator that was never written: select, do not multiply)
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (sa != 0.f ? v[i] * sa : 0.f) + (sb != 0.f ? w[i] * sb : 0.f);
    if (qi < len) {
      const int64_t grow = r0 + qi;
      uint8_t* slot = out_img + tile_image_offset(grow, h * kDH, kD) - (((int)(grow & 7)) << 4);
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
    cudaFuncSetAttribute(attn_t5_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, T5_BYTES);
    attr_done[dev & 63] = true;
  }
  attn_t5_kernel<<<dim3((max_len + 127) / 128, kH, bt.B), 288, T5_BYTES, st>>>(bt, (const uint8_t*)q_img, (const uint8_t*)k_img,
                                                                    (const uint8_t*)vt_img, R, row0, mode, (uint8_t*)out_img);
}

}  // namespace vx
