// fp32-accumulate SIMT contractions: C[M,N] = A[M,K] * W[N,K]^T (+bias, epilogue).
//
// * gemm_kernel : 128x128x16 tiles, 8x8 register micro-tiles, register-staged double buffering.
//   This is the exact-arithmetic path: fp32 parity mode of the transformer stacks and every Vocos
//   contraction (waveform RMS tolerance 1e-4 rules out bf16 operands there).  In bf16 mode the large
//   transformer GEMMs go to the tcgen05 kernel (gemm_tc.cu) instead.
// * gemv_kernel : M <= 8 rows (AR decode step).  Weight-bandwidth bound: one warp per output feature,
//   128-bit coalesced weight loads, activations (optionally LayerNorm-ed in the prologue) in shared memory,
//   warp-shuffle reduction.  modules/activation.py:144,166; modules/transformer.py:371-373.
#include "common.cuh"
#include "kernels.h"

namespace vx {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

__device__ __forceinline__ float gelu_exact(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752440f)); }

template <typename WT>
__global__ void __launch_bounds__(256) gemm_kernel(const float* __restrict__ A, int64_t lda,
                                                   const WT* __restrict__ W, const float* __restrict__ bias,
                                                   float* __restrict__ C, int64_t ldc, int64_t M, int N, int K,
                                                   int epi, const float* __restrict__ scale) {
  __shared__ __align__(16) float sA[2][BK][BM + PAD];
  __shared__ __align__(16) float sB[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int64_t m0 = (int64_t)blockIdx.y * BM;
  const int n0 = blockIdx.x * BN;

  // A tile loader: 128 rows x 16 k fp32 -> 2 float4 per thread
  const int a_row = tid >> 2, a_kq = (tid & 3) * 4;
  float4 ra[2];
  float rb[8];
  // W tile loader: fp32 -> like A; bf16 -> 128 rows x 2 chunks of 8
  const int w_row = (sizeof(WT) == 4) ? (tid >> 2) : (tid >> 1);
  const int w_kq = (sizeof(WT) == 4) ? (tid & 3) * 4 : (tid & 1) * 8;

  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const int64_t r = m0 + a_row + 64 * i;
      ra[i] = (r < M) ? *reinterpret_cast<const float4*>(A + r * lda + k0 + a_kq) : make_float4(0, 0, 0, 0);
    }
    if constexpr (sizeof(WT) == 4) {
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int n = n0 + w_row + 64 * i;
        float t4[4] = {0, 0, 0, 0};
        if (n < N) load4(W + (size_t)n * K + k0 + w_kq, t4);
        rb[4 * i] = t4[0]; rb[4 * i + 1] = t4[1]; rb[4 * i + 2] = t4[2]; rb[4 * i + 3] = t4[3];
      }
    } else {
      const int n = n0 + w_row;
      if (n < N) load8(W + (size_t)n * K + k0 + w_kq, rb);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) rb[i] = 0.f;
      }
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      sA[buf][a_kq + 0][a_row + 64 * i] = ra[i].x;
      sA[buf][a_kq + 1][a_row + 64 * i] = ra[i].y;
      sA[buf][a_kq + 2][a_row + 64 * i] = ra[i].z;
      sA[buf][a_kq + 3][a_row + 64 * i] = ra[i].w;
    }
    if constexpr (sizeof(WT) == 4) {
#pragma unroll
      for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int c = 0; c < 4; ++c) sB[buf][w_kq + c][w_row + 64 * i] = rb[4 * i + c];
    } else {
#pragma unroll
      for (int c = 0; c < 8; ++c) sB[buf][w_kq + c][w_row] = rb[c];
    }
  };

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  const int nk = K / BK;
  gload(0);
  sstore(0);
  __syncthreads();
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&sA[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&sA[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&sB[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&sB[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int64_t r = m0 + ((i < 4) ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (r >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int c0 = n0 + jh * 64 + tx * 4;
      float o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int c = c0 + j;
        float v = acc[i][jh * 4 + j];
        if (c < N) {
          if (bias) v += bias[c];
          if (epi == EPI_RELU) v = fmaxf(v, 0.f);
          else if (epi == EPI_GELU) v = gelu_exact(v);
          else if (epi == EPI_RESIDUAL) v += C[r * ldc + c];
          else if (epi == EPI_SCALE_RESIDUAL) v = C[r * ldc + c] + scale[c] * v;
        }
        o[j] = v;
      }
      if (c0 + 3 < N && (ldc & 3) == 0) {
        *reinterpret_cast<float4*>(C + r * ldc + c0) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c0 + j < N) C[r * ldc + c0 + j] = o[j];
      }
    }
  }
}

void launch_gemm(const float* A, int64_t lda, const void* W, int wt, const float* bias, float* C, int64_t ldc,
                 int64_t M, int N, int K, int epi, const float* scale, cudaStream_t st) {
  if (M <= 0 || N <= 0) return;
  dim3 grid((N + BN - 1) / BN, (unsigned)((M + BM - 1) / BM));
  if (wt == 0)
    gemm_kernel<float><<<grid, 256, 0, st>>>(A, lda, (const float*)W, bias, C, ldc, M, N, K, epi, scale);
  else
    gemm_kernel<bf16><<<grid, 256, 0, st>>>(A, lda, (const bf16*)W, bias, C, ldc, M, N, K, epi, scale);
}

// ------------------------------------------------------------------------------------------------
// Decode GEMV: M <= 8 activation rows against an [N,K] weight; optional LayerNorm prologue (eps 1e-5).
// ------------------------------------------------------------------------------------------------
constexpr int GV_WARPS = 8;

template <typename WT, int MR>
__global__ void __launch_bounds__(GV_WARPS * 32) gemv_kernel(const float* __restrict__ A, int64_t lda,
                                                             const WT* __restrict__ W,
                                                             const float* __restrict__ bias,
                                                             float* __restrict__ C, int64_t ldc, int M, int N,
                                                             int K, int epi, const float* __restrict__ ln_g,
                                                             const float* __restrict__ ln_b) {
  extern __shared__ __align__(16) float sx[];  // [MR][K]
  __shared__ float red[32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int m = 0; m < MR; ++m) {
    if (m < M) {
      if (ln_g) {  // K == 1024 here: 4 values per thread
        float v[4], s = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { v[i] = A[m * lda + tid + i * 256]; s += v[i]; }
        const float mean = block_sum(s, red) / (float)K;
        float q = 0.f;
#pragma unroll
        for (int i = 0; i < 4; ++i) { const float d = v[i] - mean; q += d * d; }
        const float rstd = 1.0f / sqrtf(block_sum(q, red) / (float)K + 1e-5f);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int c = tid + i * 256;
          sx[m * K + c] = __fadd_rn(__fmul_rn((v[i] - mean) * rstd, ln_g[c]), ln_b[c]);
        }
      } else {
        for (int c = tid; c < K; c += GV_WARPS * 32) sx[m * K + c] = A[m * lda + c];
      }
    } else {
      for (int c = tid; c < K; c += GV_WARPS * 32) sx[m * K + c] = 0.f;
    }
  }
  __syncthreads();
  const int n = blockIdx.x * GV_WARPS + warp;
  if (n >= N) return;
  const WT* wr = W + (size_t)n * K;
  float acc[MR];
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
  for (int k = lane * 8; k < K; k += 256) {
    float w[8];
    load8(wr + k, w);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * K + k]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * K + k + 4]);
      acc[m] = fmaf(w[0], x0.x, acc[m]); acc[m] = fmaf(w[1], x0.y, acc[m]);
      acc[m] = fmaf(w[2], x0.z, acc[m]); acc[m] = fmaf(w[3], x0.w, acc[m]);
      acc[m] = fmaf(w[4], x1.x, acc[m]); acc[m] = fmaf(w[5], x1.y, acc[m]);
      acc[m] = fmaf(w[6], x1.z, acc[m]); acc[m] = fmaf(w[7], x1.w, acc[m]);
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = warp_sum(acc[m]);
  if (lane == 0) {
    const float bv = bias ? bias[n] : 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      if (m < M) {
        float v = acc[m] + bv;
        if (epi == EPI_RELU) v = fmaxf(v, 0.f);
        else if (epi == EPI_RESIDUAL) v += C[m * ldc + n];
        C[m * ldc + n] = v;
      }
    }
  }
}

template <typename WT, int MR>
static void gemv_launch(const float* A, int64_t lda, const void* W, const float* bias, float* C, int64_t ldc, int M,
                        int N, int K, int epi, const float* g, const float* b, cudaStream_t st) {
  const size_t smem = (size_t)MR * K * sizeof(float);
  static bool attr_done[64] = {};  // per template instance and device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    cudaFuncSetAttribute(gemv_kernel<WT, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, 8 * 4096 * 4);
    attr_done[dev & 63] = true;
  }
  gemv_kernel<WT, MR><<<(N + GV_WARPS - 1) / GV_WARPS, GV_WARPS * 32, smem, st>>>(A, lda, (const WT*)W, bias, C, ldc,
                                                                                 M, N, K, epi, g, b);
}

void launch_gemv(const float* A, int64_t lda, const void* W, int wt, const float* bias, float* C, int64_t ldc,
                 int M, int N, int K, int epi, const float* ln_gamma, const float* ln_beta, cudaStream_t st) {
  // MR in {1,2,4,8}: padded rows cost FMAs only, the weight stream is what matters.
  const int mr = (M <= 1) ? 1 : (M <= 2) ? 2 : (M <= 4) ? 4 : 8;
#define VX_GV(WT_, MR_) gemv_launch<WT_, MR_>(A, lda, W, bias, C, ldc, M, N, K, epi, ln_gamma, ln_beta, st)
  if (wt == 0) {
    if (mr == 1) VX_GV(float, 1); else if (mr == 2) VX_GV(float, 2); else if (mr == 4) VX_GV(float, 4); else VX_GV(float, 8);
  } else {
    if (mr == 1) VX_GV(bf16, 1); else if (mr == 2) VX_GV(bf16, 2); else if (mr == 4) VX_GV(bf16, 4); else VX_GV(bf16, 8);
  }
#undef VX_GV
}

}  // namespace vx
