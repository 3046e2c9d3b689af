// tcgen05 GEMM for the full-sequence passes (AR prefill, the 7 NAR passes) in bf16 mode.
//
//   C[M,N] = A[M,K] (bf16 tile image) x W[N,K]^T (bf16 tile image), fp32 accumulate in TMEM, fused epilogue.
//
// Persistent, warp-specialised (see tc_common.cuh for the operand layout):
//   warp 0   : producer -- one elected lane streams 16 KB operand blocks HBM -> shared with cp.async.bulk
//              (TMA engine) through a kStages-deep mbarrier ring;
//   warp 1   : allocates TMEM; one elected lane issues tcgen05.mma (M=128, N=BN, K=16; 4 per 64-wide k block),
//              tcgen05.commit frees each ring slot and publishes a finished accumulator;
//   warps 2-9: epilogue (two per TMEM lane quarter, half the columns each) -- tcgen05.ld the 128 x BN fp32 accumulator (double-buffered in TMEM, so the next tile's
//              MMAs overlap), add bias / ReLU / residual, store bf16 row-major, bf16 tile image or fp32.
// Bound: tensor pipe (dense bf16).  Tiles are walked m-fastest so CTAs running concurrently share the same
// weight blocks through L2.  modules/activation.py:144,166; modules/transformer.py:371-373; models/vallex.py:677.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vx {
using namespace tc;

enum TcEpi { TC_BF16_ROW = 0, TC_F32_RESIDUAL = 1, TC_BF16_IMAGE_RELU = 2, TC_F32 = 3, TC_QKV_HEADS = 4, TC_F32_SCALE_RESIDUAL = 5,
             TC_GELU_SPLIT3 = 6 };
__device__ __forceinline__ float tc_gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <int BN>
struct TcCfg {
  static constexpr int kStages = (BN == 256) ? 4 : 6;
  static constexpr int kABytes = kBlockBytes;
  static constexpr int kBBytes = kBlockBytes * (BN / 128);
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kTmemCols = 2 * BN;
  static constexpr int kSmemBytes = kStages * kStageBytes + 256 /*barriers*/ + 8 * 4096 /*epilogue staging, 4 KB per warp*/;
};

// 8 epilogue warps: two per TMEM lane quarter, each owning half of the tile's columns.  With 4 warps (round 1/2) a K = 1024 tile's
// epilogue (a 128 x 256 fp32 read-modify-write through 128 threads, one dependent L2/HBM round trip per 32-column chunk) took ~16 us
// against 4.3 us of MMAs -- the out-proj GEMM ran at 26 % tensor-pipe activity (profiles/r02_prof_gemm_b8.md).
constexpr int kEpiWarps = 8;
constexpr int kTcThreads = 64 + 32 * kEpiWarps;

template <int BN, int EPI>
__global__ void __launch_bounds__(kTcThreads, 1) gemm_tc_kernel(const bf16* __restrict__ A, const bf16* __restrict__ W,
                                                         const float* __restrict__ bias, void* __restrict__ out,
                                                         int64_t ldc, int64_t M, int N, int K, const float* __restrict__ aux) {
  using Cfg = TcCfg<BN>;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if ((smem_u32(smem) & 1023u) != 0) __trap();   // the swizzled operand stages need 1024-byte alignment
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* empty = full + Cfg::kStages;
  uint64_t* tfull = empty + Cfg::kStages;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (int)((M + 127) >> 7);
  const int n_tiles = N / BN;
  const int k_tiles = K >> 6;
  const int tiles = m_tiles * n_tiles;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(&tfull[a], 1); mbar_init(&tempty[a], kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x) {
        const int mt = t % m_tiles, nt = t / m_tiles;
        const uint8_t* a_src = reinterpret_cast<const uint8_t*>(A) + (size_t)mt * k_tiles * kBlockBytes;
        for (int kt = 0; kt < k_tiles; ++kt) {
          mbar_wait(&empty[stage], phase ^ 1);
          uint8_t* sa = smem + stage * Cfg::kStageBytes;
          mbar_expect_tx(&full[stage], Cfg::kStageBytes);
          bulk_g2s(sa, a_src + (size_t)kt * kBlockBytes, kBlockBytes, &full[stage]);
#pragma unroll
          for (int j = 0; j < BN / 128; ++j) {
            const uint8_t* w_src = reinterpret_cast<const uint8_t*>(W) +
                                   ((size_t)(nt * (BN / 128) + j) * k_tiles + kt) * kBlockBytes;
            bulk_g2s(sa + Cfg::kABytes + j * kBlockBytes, w_src, kBlockBytes, &full[stage]);
          }
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(128, BN);
      int stage = 0;
      uint32_t phase = 0, acc_phase[2] = {0, 0};
      int it = 0;
      for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
        const int acc = it & 1;
        mbar_wait(&tempty[acc], acc_phase[acc] ^ 1);
        acc_phase[acc] ^= 1;
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kt = 0; kt < k_tiles; ++kt) {
          mbar_wait(&full[stage], phase);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + stage * Cfg::kStageBytes);
          const uint64_t adesc = umma_desc_sw128(sa);
          const uint64_t bdesc = umma_desc_sw128(sa + Cfg::kABytes);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // advance 16 bf16 = 32 bytes inside the 128-byte swizzle row: +2 in 16-byte units
            umma_bf16(d_tmem, adesc + 2 * k, bdesc + 2 * k, idesc, (kt | k) ? 1u : 0u);
          umma_commit(&empty[stage]);
          if (++stage == Cfg::kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tfull[acc]);
      }
    }
  } else {
    // ---- epilogue: TMEM -> registers -> (shared-memory transpose) -> global -------------------------------------------------
    // tcgen05.ld hands a thread ONE accumulator row (32 consecutive columns per chunk), so direct stores put every lane of a
    // warp on a different 128-byte line: 32 L1 wavefronts per instruction.  Measured (profiles/r02_prof_gemm_b8.md): the fp32
    // read-modify-write of a 128 x 256 tile cost ~16 us against 4.3 us of MMAs (out-proj at 26 % tensor activity, no gain from
    // doubling the epilogue warps: the LSU, not latency, was the limit) and the 2-byte V^T scatter ~17 us per tile.  Each warp
    // therefore stages its 32 x 32 chunk in a swizzled 4 KB buffer and re-reads it so that 8 (fp32) or 4 (bf16) lanes cover one
    // row: every global instruction touches 4-8 full lines.
    const int q = warp & 3;                        // TMEM lane quarter this warp may access
    const int chalf = (warp - 2) >> 2;             // which half of the tile's columns this warp owns
    constexpr int kChunks = BN / 2 / 32;           // 32-column chunks per warp
    float* stg = reinterpret_cast<float*>(smem + Cfg::kStages * Cfg::kStageBytes + 256) + (warp - 2) * 1024;
    constexpr bool kRes = (EPI == TC_F32_RESIDUAL || EPI == TC_F32_SCALE_RESIDUAL);
    constexpr bool kF32Out = kRes || EPI == TC_F32;
    const int c4 = lane & 7, r4 = lane >> 3;       // fp32 outputs: lane = (row r4 + 4 i, columns 4 c4 .. +3)
    const int c8 = lane & 3, r8 = lane >> 2;       // bf16 outputs: lane = (row r8 + 8 i, columns 8 c8 .. +7)
    auto pk = [](float a, float b) {
      return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
    };
    uint32_t acc_phase[2] = {0, 0};
    int it = 0;
    for (int t = blockIdx.x; t < tiles; t += gridDim.x, ++it) {
      const int acc = it & 1;
      const int mt = t % m_tiles, nt = t / m_tiles;
      const int64_t wrow0 = (int64_t)mt * 128 + q * 32;    // first row of this warp's 32
      const int cbase = chalf * (BN / 2);
      // residual epilogues: the first chunks of the warp's fp32 residual slice are requested BEFORE waiting for the accumulator,
      // the rest two chunks ahead (profiles/r01_prof_gemm.md: 54 % of the epilogue warps' stall samples were these loads)
      float4 rA[8], rB[8];     // residual of the chunk being folded / of the next one; chunk cc + 2 is requested when cc is done
      if constexpr (kRes) {
        const float* o = reinterpret_cast<const float*>(out) + nt * BN + cbase + c4 * 4;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (wrow0 + r4 + 4 * i < M) {
            rA[i] = *reinterpret_cast<const float4*>(o + (wrow0 + r4 + 4 * i) * ldc);
            if (kChunks > 1) rB[i] = *reinterpret_cast<const float4*>(o + (wrow0 + r4 + 4 * i) * ldc + 32);
          }
      }
      mbar_wait(&tfull[acc], acc_phase[acc]);
      acc_phase[acc] ^= 1;
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BN + ((uint32_t)(q * 32) << 16);
#pragma unroll 1
      for (int cc = 0; cc < kChunks; ++cc) {
        const int c0 = cbase + cc * 32;
        const int n0 = nt * BN + c0;
        float v[32];
        tmem_ld32(taddr + c0, v);
        // stage: row `lane`, 16-byte chunk c at position c ^ (lane & 7) (conflict-free for the row-wise write and both read maps)
#pragma unroll
        for (int c = 0; c < 8; ++c)
          *reinterpret_cast<float4*>(stg + lane * 32 + ((c ^ (lane & 7)) << 2)) = make_float4(v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]);
        __syncwarp();
        if constexpr (kF32Out) {
          float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
          if (bias) b4 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c4 * 4));
          float4 g4 = make_float4(1.f, 1.f, 1.f, 1.f);
          if constexpr (EPI == TC_F32_SCALE_RESIDUAL) g4 = __ldg(reinterpret_cast<const float4*>(aux + n0 + c4 * 4));
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const int rr = r4 + 4 * i;
            const int64_t row = wrow0 + rr;
            float4 a = *reinterpret_cast<const float4*>(stg + rr * 32 + ((c4 ^ (rr & 7)) << 2));
            a.x += b4.x; a.y += b4.y; a.z += b4.z; a.w += b4.w;
            if (row < M) {
              float* o = reinterpret_cast<float*>(out) + row * ldc + n0 + c4 * 4;
              if constexpr (EPI == TC_F32_RESIDUAL) {
                float4 r = rA[i];
                r.x += a.x; r.y += a.y; r.z += a.z; r.w += a.w;
                *reinterpret_cast<float4*>(o) = r;
              } else if constexpr (EPI == TC_F32_SCALE_RESIDUAL) {   // Vocos pwconv2: x += gamma[n] * (acc + bias[n])
                float4 r = rA[i];
                r.x = fmaf(g4.x, a.x, r.x); r.y = fmaf(g4.y, a.y, r.y); r.z = fmaf(g4.z, a.z, r.z); r.w = fmaf(g4.w, a.w, r.w);
                *reinterpret_cast<float4*>(o) = r;
              } else {
                *reinterpret_cast<float4*>(o) = a;
              }
            }
          }
          if constexpr (kRes) {   // rotate: the next chunk's residual becomes current, chunk cc + 2 is requested
            const float* o = reinterpret_cast<const float*>(out) + nt * BN + cbase + (cc + 2) * 32 + c4 * 4;
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              rA[i] = rB[i];
              if (cc + 2 < kChunks && wrow0 + r4 + 4 * i < M) rB[i] = *reinterpret_cast<const float4*>(o + (wrow0 + r4 + 4 * i) * ldc);
            }
          }
        } else if (EPI == TC_QKV_HEADS && n0 >= 2 * kD) {
          // V columns: transposed per 64-key block, [64 d rows x 64 keys] K-major swizzled (the B operand of P V).  Lane = head
          // dim d; 8 consecutive keys (rows of the chunk) form one 16-byte chunk of the block's row d.  ldc = R (multiple of 128).
          const int hc = n0 % kD, h = hc >> 6, d = (hc & 63) + lane;
          const float bd = bias ? __ldg(bias + n0 + lane) : 0.f;
          uint8_t* base = reinterpret_cast<uint8_t*>(out) + (size_t)2 * ((size_t)ldc * kD * 2);
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float e[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int rr = 8 * g + j;
              e[j] = stg[rr * 32 + ((((lane >> 2) ^ (rr & 7))) << 2) + (lane & 3)] + bd;
            }
            const int64_t row = wrow0 + 8 * g;
            if (row < M) {
              uint8_t* vb = base + ((size_t)h * (size_t)(ldc / 64) + (size_t)(row >> 6)) * 8192;
              const int kc = (int)((row & 63) >> 3);
              *reinterpret_cast<uint4*>(vb + d * 128 + ((kc ^ (d & 7)) << 4)) =
                  make_uint4(pk(e[0], e[1]), pk(e[2], e[3]), pk(e[4], e[5]), pk(e[6], e[7]));
            }
          }
        } else {
          // bf16 outputs: lane = (row r8 + 8 i, columns 8 c8 .. +7): 16-byte stores, 4 lanes per row
          float bb[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          if (bias) {
            const float4 b0 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c8 * 8)), b1 = __ldg(reinterpret_cast<const float4*>(bias + n0 + c8 * 8 + 4));
            bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int rr = r8 + 8 * i;
            const int64_t row = wrow0 + rr;
            const float4 a0 = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * c8) ^ (rr & 7)) << 2));
            const float4 a1 = *reinterpret_cast<const float4*>(stg + rr * 32 + (((2 * c8 + 1) ^ (rr & 7)) << 2));
            float y[8] = {a0.x + bb[0], a0.y + bb[1], a0.z + bb[2], a0.w + bb[3], a1.x + bb[4], a1.y + bb[5], a1.z + bb[6], a1.w + bb[7]};
            if (row >= M) continue;
            const int nc = n0 + c8 * 8;
            if constexpr (EPI == TC_BF16_ROW) {
              *reinterpret_cast<uint4*>(reinterpret_cast<bf16*>(out) + row * ldc + nc) = make_uint4(pk(y[0], y[1]), pk(y[2], y[3]), pk(y[4], y[5]), pk(y[6], y[7]));
            } else if constexpr (EPI == TC_BF16_IMAGE_RELU) {
#pragma unroll
              for (int e = 0; e < 8; ++e) y[e] = fmaxf(y[e], 0.f);
              *reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(out) + tile_image_offset(row, nc, N)) =
                  make_uint4(pk(y[0], y[1]), pk(y[2], y[3]), pk(y[4], y[5]), pk(y[6], y[7]));
            } else if constexpr (EPI == TC_QKV_HEADS) {
              // q and k columns: per-head K-major tile images [R,64] (same layouts as pack_qkv_heads_kernel, attention_t5.cu)
              const int which = nc / kD, hc = nc % kD, h = hc >> 6, d0 = hc & 63;
              uint8_t* img = reinterpret_cast<uint8_t*>(out) + (size_t)which * ((size_t)ldc * kD * 2) + (size_t)h * ((size_t)ldc * 128);
              *reinterpret_cast<uint4*>(img + tile_image_offset(row, d0, 64)) = make_uint4(pk(y[0], y[1]), pk(y[2], y[3]), pk(y[4], y[5]), pk(y[6], y[7]));
            } else if constexpr (EPI == TC_GELU_SPLIT3) {
              // Vocos pwconv1: GELU(acc + bias) as the NEXT contraction's split-3 operand image [rows, 3N] = [hi | hi | lo]
              // (bf16 x 3: a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi against the weight image [hi | lo | hi])
              uint32_t hw[4], lw[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const float x0 = tc_gelu_exact(y[2 * e]), x1 = tc_gelu_exact(y[2 * e + 1]);
                const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
                const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
                hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
                lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
              }
              uint8_t* img = reinterpret_cast<uint8_t*>(out);
              const uint4 hp = make_uint4(hw[0], hw[1], hw[2], hw[3]), lp = make_uint4(lw[0], lw[1], lw[2], lw[3]);
              *reinterpret_cast<uint4*>(img + tile_image_offset(row, nc, 3 * N)) = hp;
              *reinterpret_cast<uint4*>(img + tile_image_offset(row, N + nc, 3 * N)) = hp;
              *reinterpret_cast<uint4*>(img + tile_image_offset(row, 2 * N + nc, 3 * N)) = lp;
            }
          }
        }
        __syncwarp();   // the staging buffer is rewritten by the next chunk
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc<Cfg::kTmemCols>(tmem_base);
}

template <int BN, int EPI>
static void tc_launch(const bf16* A, const bf16* W, const float* bias, void* out, int64_t ldc, int64_t M, int N, int K,
                      const float* aux, cudaStream_t st) {
  using Cfg = TcCfg<BN>;
  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaFuncSetAttribute(gemm_tc_kernel<BN, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes);
    attr_done[dev & 63] = true;
  }
  const int tiles = (int)((M + 127) / 128) * (N / BN);
  const int grid = tiles < kSMs ? tiles : kSMs;
  gemm_tc_kernel<BN, EPI><<<grid, kTcThreads, Cfg::kSmemBytes, st>>>(A, W, bias, out, ldc, M, N, K, aux);
}

// epi: TcEpi.  N % 256 == 0 uses 128x256 tiles (shared-memory read rate of the A/B operands stays under
// 128 B/clk/SM), otherwise 128x128.  K % 64 == 0.  A and W are tile images; M is padded to 128 by the caller's
// allocation (rows >= M are never stored).
void launch_gemm_tc(const void* A_img, const void* W_img, const float* bias, void* out, int64_t ldc, int64_t M, int N,
                    int K, int epi, cudaStream_t st, const float* aux) {
  if (M <= 0) return;
  const bf16* A = (const bf16*)A_img;
  const bf16* W = (const bf16*)W_img;
#define VX_TC(BN_)                                                                                      \
  switch (epi) {                                                                                        \
    case TC_BF16_ROW: tc_launch<BN_, TC_BF16_ROW>(A, W, bias, out, ldc, M, N, K, aux, st); break;            \
    case TC_F32_RESIDUAL: tc_launch<BN_, TC_F32_RESIDUAL>(A, W, bias, out, ldc, M, N, K, aux, st); break;    \
    case TC_BF16_IMAGE_RELU: tc_launch<BN_, TC_BF16_IMAGE_RELU>(A, W, bias, out, ldc, M, N, K, aux, st); break; \
    case TC_QKV_HEADS: tc_launch<BN_, TC_QKV_HEADS>(A, W, bias, out, ldc, M, N, K, aux, st); break;          \
    case TC_F32_SCALE_RESIDUAL: tc_launch<BN_, TC_F32_SCALE_RESIDUAL>(A, W, bias, out, ldc, M, N, K, aux, st); break; \
    case TC_GELU_SPLIT3: tc_launch<BN_, TC_GELU_SPLIT3>(A, W, bias, out, ldc, M, N, K, aux, st); break; \
    default: tc_launch<BN_, TC_F32>(A, W, bias, out, ldc, M, N, K, aux, st); break;                         \
  }
  // 128x256 tiles keep the operand read rate under the shared-memory limit, but a small problem needs enough tiles to
  // cover the 148 SMs: fall back to 128x128 when the 256-wide tiling would leave most of the chip idle.
  const int64_t tiles256 = ((M + 127) / 128) * (N / 256);
  if (N % 256 == 0 && tiles256 >= kSMs) { VX_TC(256) } else { VX_TC(128) }
#undef VX_TC
}

// ---- operand packing -------------------------------------------------------------------------------
// fp32 row-major [rows, K] -> bf16 tile image (rows padded to 128 with zeros).  Used once per weight at load.
__global__ void pack_tile_image_kernel(const float* __restrict__ src, int64_t rows, int K, uint8_t* __restrict__ img,
                                       int64_t rows_pad) {
  const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 elements)
  const int64_t total = rows_pad * (K >> 3);
  if (chunk >= total) return;
  const int64_t r = chunk / (K >> 3);
  const int k = (int)(chunk % (K >> 3)) * 8;
  uint4 p = make_uint4(0, 0, 0, 0);
  if (r < rows) {
    float v[8];
    load8(src + r * K + k, v);
    auto pk = [](float a, float b) {
      return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
    };
    p.x = pk(v[0], v[1]); p.y = pk(v[2], v[3]); p.z = pk(v[4], v[5]); p.w = pk(v[6], v[7]);
  }
  *reinterpret_cast<uint4*>(img + tile_image_offset(r, k, K)) = p;
}
void launch_pack_tile_image(const float* src, int64_t rows, int K, void* img, cudaStream_t st) {
  const int64_t rows_pad = (rows + 127) / 128 * 128;
  const int64_t total = rows_pad * (K >> 3);
  pack_tile_image_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, rows, K, (uint8_t*)img, rows_pad);
}

// fp32 weight [rows, K] -> split-3 bf16 tile image [rows_pad, 3K] = [hi | lo | hi] (rows padded to 128 with zeros): the B
// operand of the bf16 x 3 contractions (Vocos), against activation images [hi | hi | lo].  ld = row stride of src (>= K_src);
// columns >= K_src up to K are zero padding.
__global__ void pack_split3_weight_kernel(const float* __restrict__ src, int64_t rows, int K_src, int64_t ld, int K,
                                          uint8_t* __restrict__ img, int64_t rows_pad) {
  const int64_t chunk = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // one 16-byte chunk (8 elements) of the [rows_pad, K] grid
  const int64_t total = rows_pad * (K >> 3);
  if (chunk >= total) return;
  const int64_t r = chunk / (K >> 3);
  const int k = (int)(chunk % (K >> 3)) * 8;
  uint32_t hw[4] = {0, 0, 0, 0}, lw[4] = {0, 0, 0, 0};
  if (r < rows) {
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float x0 = (k + 2 * e < K_src) ? src[r * ld + k + 2 * e] : 0.f, x1 = (k + 2 * e + 1 < K_src) ? src[r * ld + k + 2 * e + 1] : 0.f;
      const __nv_bfloat16 h0 = __float2bfloat16_rn(x0), h1 = __float2bfloat16_rn(x1);
      const __nv_bfloat16 l0 = __float2bfloat16_rn(x0 - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x1 - __bfloat162float(h1));
      hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
      lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
    }
  }
  const uint4 hp = make_uint4(hw[0], hw[1], hw[2], hw[3]), lp = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  *reinterpret_cast<uint4*>(img + tile_image_offset(r, k, 3 * K)) = hp;
  *reinterpret_cast<uint4*>(img + tile_image_offset(r, K + k, 3 * K)) = lp;
  *reinterpret_cast<uint4*>(img + tile_image_offset(r, 2 * K + k, 3 * K)) = hp;
}
void launch_pack_split3_weight(const float* src, int64_t rows, int K_src, int64_t ld, int K, void* img, cudaStream_t st) {
  const int64_t rows_pad = (rows + 127) / 128 * 128;
  const int64_t total = rows_pad * (K >> 3);
  pack_split3_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(src, rows, K_src, ld, K, (uint8_t*)img, rows_pad);
}

// LayerNorm / AdaLN rows -> bf16 tile image (the A operand of the next GEMM); same math as layernorm_kernel.
// One warp per row (values in registers, shuffle reductions only); HBM/L2 bound: 4 KB in, 2 KB out per row.  Persistent: a fixed
// grid of 8-warp CTAs strides over the rows and every warp requests its NEXT row before it reduces the current one (the
// one-row-per-warp launch ran at half the copy rate: 27.7 us for 15,360 rows, profiles/r02_launches_cfg4_b8.md), and the affine
// vectors (gamma, beta, AdaLN weight / bias: the same 16 KB for every row) are read once per warp, not once per row.
__device__ __forceinline__ void ln_row_load(const float* __restrict__ x, int64_t src, int lane, float4 (&v)[8]) {
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const float* p = x + src * kD + it * 256 + lane * 8;
    v[2 * it] = *reinterpret_cast<const float4*>(p);
    v[2 * it + 1] = *reinterpret_cast<const float4*>(p + 4);
  }
}
__global__ void __launch_bounds__(256) layernorm_image_kernel(const float* __restrict__ x,
                                                              const int32_t* __restrict__ row_map, int64_t rows,
                                                              const float* __restrict__ gamma,
                                                              const float* __restrict__ beta,
                                                              const float* __restrict__ ada_w,
                                                              const float* __restrict__ ada_b, float eps,
                                                              uint8_t* __restrict__ img) {
  const int lane = threadIdx.x & 31;
  const int64_t nwarps = (int64_t)gridDim.x * 8;
  int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  // y = ((v - mean) * rstd) * g + be, then (AdaLN) aw * y + ab, in exactly this order (modules/transformer.py:17-80,83-108);
  // the four affine vectors live in shared memory for the CTA's lifetime
  __shared__ __align__(16) float aff[4][kD];
  for (int i = threadIdx.x; i < kD; i += 256) {
    aff[0][i] = gamma[i]; aff[1][i] = beta[i];
    if (ada_w) { aff[2][i] = ada_w[i]; aff[3][i] = ada_b[i]; }
  }
  __syncthreads();
  if (r >= rows) return;
  auto pk = [](float a, float b) {
    return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
  };
  float4 cur[8], nxt[8];
  ln_row_load(x, row_map ? row_map[r] : r, lane, cur);
  for (; r < rows; r += nwarps) {
    const int64_t rn = r + nwarps;
    if (rn < rows) ln_row_load(x, row_map ? row_map[rn] : rn, lane, nxt);
    float v[4][8];
    float s = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const float4 a = cur[2 * it], b = cur[2 * it + 1];
      v[it][0] = a.x; v[it][1] = a.y; v[it][2] = a.z; v[it][3] = a.w; v[it][4] = b.x; v[it][5] = b.y; v[it][6] = b.z; v[it][7] = b.w;
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[it][i];
    }
    const float mean = warp_sum(s) / (float)kD;
    float qv = 0.f;
#pragma unroll
    for (int it = 0; it < 4; ++it)
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = v[it][i] - mean; qv += d * d; }
    const float rstd = 1.0f / sqrtf(warp_sum(qv) / (float)kD + eps);
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const int c0 = it * 256 + lane * 8;
      float y[8], g[8], be[8];
      *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(&aff[0][c0]);
      *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(&aff[0][c0 + 4]);
      *reinterpret_cast<float4*>(be) = *reinterpret_cast<const float4*>(&aff[1][c0]);
      *reinterpret_cast<float4*>(be + 4) = *reinterpret_cast<const float4*>(&aff[1][c0 + 4]);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = ((v[it][i] - mean) * rstd) * g[i] + be[i];
      if (ada_w) {
        *reinterpret_cast<float4*>(g) = *reinterpret_cast<const float4*>(&aff[2][c0]);
        *reinterpret_cast<float4*>(g + 4) = *reinterpret_cast<const float4*>(&aff[2][c0 + 4]);
        *reinterpret_cast<float4*>(be) = *reinterpret_cast<const float4*>(&aff[3][c0]);
        *reinterpret_cast<float4*>(be + 4) = *reinterpret_cast<const float4*>(&aff[3][c0 + 4]);
#pragma unroll
        for (int i = 0; i < 8; ++i) y[i] = g[i] * y[i] + be[i];
      }
      uint4 p;
      p.x = pk(y[0], y[1]); p.y = pk(y[2], y[3]); p.z = pk(y[4], y[5]); p.w = pk(y[6], y[7]);
      *reinterpret_cast<uint4*>(img + tile_image_offset(r, c0, kD)) = p;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) cur[i] = nxt[i];
  }
}
void launch_layernorm_image(const float* x, const int32_t* row_map, int64_t rows, const float* gamma,
                            const float* beta, const float* ada_w, const float* ada_b, float eps, void* img,
                            cudaStream_t st) {
  if (rows <= 0) return;
  const int64_t want = (rows + 7) / 8;
  const unsigned grid = (unsigned)std::min<int64_t>(want, (int64_t)kSMs * 2);   // 2 CTAs of 8 warps per SM resident (110 registers)
  layernorm_image_kernel<<<grid, 256, 0, st>>>(x, row_map, rows, gamma, beta, ada_w, ada_b, eps, (uint8_t*)img);
}

}  // namespace vx
