// Vocos decode kernels (third-party `vocos` package semantics; see oracle/vocos_oracle.py header).
// Activations are time-major [B*T, C] fp32; every contraction (k=7 embed conv via im2col, the two
// point-wise linears, the head and the inverse DFT expressed as a [1280 x 1282] basis contraction) runs
// on the fp32 SIMT GEMM (gemm_simt.cu) because the waveform tolerance (RMS 1e-4) rules out bf16 operands.
#include "common.cuh"
#include "kernels.h"
#include "tc_common.cuh"

namespace vx {

constexpr int VC_IN = 128, VC_NFFT = 1280, VC_HOP = 320, VC_BINS = 641;

// codes [8,B,T] -> feats [B*T,128]: sum over the 8 codebooks of table[code + 1024*k]
// (vocos pretrained.py codes_to_features; call site utils/generation.py:148-149).
__global__ void __launch_bounds__(128) vocos_features_kernel(const int32_t* __restrict__ codes, int B, int T,
                                                             const float* __restrict__ table,
                                                             float* __restrict__ feats) {
  const int64_t r = blockIdx.x;  // b*T + t
  const int c = threadIdx.x;
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kNQ; ++k) {
    const int code = min(max(codes[(size_t)k * B * T + r], 0), 1023);   // range is checked by the caller (engine.cu); never index out of the table
    s = __fadd_rn(s, table[((size_t)code + 1024 * k) * VC_IN + c]);
  }
  feats[r * VC_IN + c] = s;
}
void launch_vocos_features(const int32_t* codes, int B, int T, const float* table, float* feats, cudaStream_t st) {
  vocos_features_kernel<<<(unsigned)((size_t)B * T), VC_IN, 0, st>>>(codes, B, T, table, feats);
}

// im2col for Conv1d(k=7, padding=3): out[b*T+t][j*C + ci] = in[b][t+j-3][ci] (0 outside the utterance)
__global__ void __launch_bounds__(128) vocos_im2col7_kernel(const float* __restrict__ in, int B, int T, int C,
                                                            float* __restrict__ out) {
  const int64_t r = blockIdx.x;
  const int t = (int)(r % T);
  const int j = blockIdx.y;
  const int ts = t + j - 3;
  for (int c = threadIdx.x; c < C; c += 128)
    out[r * (7 * C) + j * C + c] = (ts >= 0 && ts < T) ? in[(r + (j - 3)) * C + c] : 0.f;
}
void launch_vocos_im2col7(const float* in, int B, int T, int C, float* out, cudaStream_t st) {
  vocos_im2col7_kernel<<<dim3((unsigned)((size_t)B * T), 7), 128, 0, st>>>(in, B, T, C, out);
}

// depthwise Conv1d(C, C, k=7, padding=3, groups=C); w [C,7]
__global__ void __launch_bounds__(128) vocos_dwconv7_kernel(const float* __restrict__ x, int B, int T, int C,
                                                            const float* __restrict__ w,
                                                            const float* __restrict__ bias,
                                                            float* __restrict__ out) {
  const int64_t r = blockIdx.x;
  const int t = (int)(r % T);
  for (int c = threadIdx.x; c < C; c += 128) {
    float s = bias[c];
#pragma unroll
    for (int j = 0; j < 7; ++j) {
      const int ts = t + j - 3;
      if (ts >= 0 && ts < T) s = fmaf(w[c * 7 + j], x[(r + (j - 3)) * C + c], s);
    }
    out[r * C + c] = s;
  }
}
void launch_vocos_dwconv7(const float* x, int B, int T, int C, const float* w, const float* bias, float* out,
                          cudaStream_t st) {
  vocos_dwconv7_kernel<<<(unsigned)((size_t)B * T), 128, 0, st>>>(x, B, T, C, w, bias, out);
}

// head_out [rows,1282] = [log-mag(641) | phase(641)] -> spec [rows, ldspec] = [re(641) | im(641) | 0...]
// mag = min(exp(m), 100); re = mag cos p; im = mag sin p   (vocos heads.py ISTFTHead.forward)
__global__ void __launch_bounds__(256) vocos_spectrum_kernel(const float* __restrict__ head_out,
                                                             float* __restrict__ spec, int ldspec) {
  const int64_t r = blockIdx.x;
  const float* h = head_out + r * (2 * VC_BINS);
  float* s = spec + r * ldspec;
  for (int k = threadIdx.x; k < ldspec; k += 256) {
    if (k < VC_BINS) {
      const float mag = fminf(expf(h[k]), 100.0f);
      const float p = h[VC_BINS + k];
      s[k] = mag * cosf(p);
      s[VC_BINS + k] = mag * sinf(p);
    } else if (k >= 2 * VC_BINS) {
      s[k] = 0.f;
    }
  }
}
void launch_vocos_spectrum(const float* head_out, int64_t rows, float* spec, int ldspec, cudaStream_t st) {
  vocos_spectrum_kernel<<<(unsigned)rows, 256, 0, st>>>(head_out, spec, ldspec);
}

// =================================================================================================================
// Tensor-core path (engine bf16 mode): every Vocos contraction runs on the tcgen05 GEMM (gemm_tc.cu) with bf16 x 3 split
// operands -- activation image [rows, 3K] = [hi | hi | lo] against weight image [N, 3K] = [hi | lo | hi], i.e.
// a*b ~ a_hi*b_hi + a_hi*b_lo + a_lo*b_hi with fp32 accumulation in TMEM (relative error ~2^-16: the waveform bar is RMS 1e-4).
// The kernels below produce those activation images directly from the fp32 stream.
// =================================================================================================================
__device__ __forceinline__ void split3_store8(uint8_t* img, int64_t r, int k, int K, const float (&x)[8]) {
  uint32_t hw[4], lw[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
    const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * e] - __bfloat162float(h0)), l1 = __float2bfloat16_rn(x[2 * e + 1] - __bfloat162float(h1));
    hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
    lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
  }
  const uint4 hp = make_uint4(hw[0], hw[1], hw[2], hw[3]), lp = make_uint4(lw[0], lw[1], lw[2], lw[3]);
  *reinterpret_cast<uint4*>(img + tc::tile_image_offset(r, k, 3 * K)) = hp;
  *reinterpret_cast<uint4*>(img + tc::tile_image_offset(r, K + k, 3 * K)) = hp;
  *reinterpret_cast<uint4*>(img + tc::tile_image_offset(r, 2 * K + k, 3 * K)) = lp;
}

// im2col for Conv1d(128 -> 384, k=7, padding=3) straight into the split-3 image: column j*128 + ci = feats[t+j-3][ci]
__global__ void __launch_bounds__(128) vocos_im2col_split3_kernel(const float* __restrict__ feats, int B, int T, uint8_t* __restrict__ img) {
  const int64_t r = blockIdx.x;
  const int t = (int)(r % T);
  const int chunk = threadIdx.x;          // 112 chunks of 8 columns (896 = 7 taps x 128 channels)
  if (chunk >= 112) return;
  const int j = chunk >> 4, c0 = (chunk & 15) * 8;
  const int ts = t + j - 3;
  float x[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ts >= 0 && ts < T) load8(feats + (r + (j - 3)) * VC_IN + c0, x);
  split3_store8(img, r, j * VC_IN + c0, 7 * VC_IN, x);
}
void launch_vocos_im2col_split3(const float* feats, int B, int T, void* img, cudaStream_t st) {
  vocos_im2col_split3_kernel<<<(unsigned)((size_t)B * T), 128, 0, st>>>(feats, B, T, (uint8_t*)img);
}

// depthwise Conv1d(384, k=7, padding=3) + AdaLayerNorm (LN without affine, eps 1e-6, * scale + shift) -> split-3 image.
// One warp per frame; a lane owns 8-channel chunks lane and lane + 32 (48 chunks); dw_wT is the tap-major copy [7][384].
__global__ void __launch_bounds__(256) vocos_dwconv_ln_split3_kernel(const float* __restrict__ x, int B, int T, const float* __restrict__ dw_wT,
                                                                     const float* __restrict__ dw_b, const float* __restrict__ scale,
                                                                     const float* __restrict__ shift, uint8_t* __restrict__ img) {
  constexpr int C = 384;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= (int64_t)B * T) return;
  const int t = (int)(r % T);
  float h[2][8];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c0 = (lane + 32 * u) * 8;
    const bool on = c0 < C;
    if (on) load8(dw_b + c0, h[u]);
    else { for (int i = 0; i < 8; ++i) h[u][i] = 0.f; }
    if (on) {
#pragma unroll
      for (int j = 0; j < 7; ++j) {
        const int ts = t + j - 3;
        if (ts >= 0 && ts < T) {
          float w[8], xv[8];
          load8(dw_wT + j * C + c0, w);
          load8(x + (r + (j - 3)) * C + c0, xv);
#pragma unroll
          for (int i = 0; i < 8; ++i) h[u][i] = fmaf(w[i], xv[i], h[u][i]);
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) s += h[u][i];
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if ((lane + 32 * u) * 8 < C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = h[u][i] - mean; q += d * d; }
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + 1e-6f);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c0 = (lane + 32 * u) * 8;
    if (c0 < C) {
      float sc[8], sh[8], y[8];
      load8(scale + c0, sc);
      load8(shift + c0, sh);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = __fadd_rn(__fmul_rn((h[u][i] - mean) * rstd, sc[i]), sh[i]);
      split3_store8(img, r, c0, C, y);
    }
  }
}
void launch_vocos_dwconv_ln_split3(const float* x, int B, int T, const float* dw_wT, const float* dw_b, const float* scale,
                                   const float* shift, void* img, cudaStream_t st) {
  const int64_t R = (int64_t)B * T;
  vocos_dwconv_ln_split3_kernel<<<(unsigned)((R + 7) / 8), 256, 0, st>>>(x, B, T, dw_wT, dw_b, scale, shift, (uint8_t*)img);
}

// LayerNorm(384, eps 1e-6, affine) -> split-3 image (backbone.final_layer_norm feeding head.out)
__global__ void __launch_bounds__(256) vocos_ln_split3_kernel(const float* __restrict__ x, int64_t rows, const float* __restrict__ gamma,
                                                              const float* __restrict__ beta, uint8_t* __restrict__ img) {
  constexpr int C = 384;
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= rows) return;
  float h[2][8];
  float s = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c0 = (lane + 32 * u) * 8;
    if (c0 < C) {
      load8(x + r * C + c0, h[u]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += h[u][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) h[u][i] = 0.f;
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int u = 0; u < 2; ++u)
    if ((lane + 32 * u) * 8 < C) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { const float d = h[u][i] - mean; q += d * d; }
    }
  const float rstd = 1.0f / sqrtf(warp_sum(q) / (float)C + 1e-6f);
#pragma unroll
  for (int u = 0; u < 2; ++u) {
    const int c0 = (lane + 32 * u) * 8;
    if (c0 < C) {
      float g[8], b[8], y[8];
      load8(gamma + c0, g);
      load8(beta + c0, b);
#pragma unroll
      for (int i = 0; i < 8; ++i) y[i] = __fadd_rn(__fmul_rn((h[u][i] - mean) * rstd, g[i]), b[i]);
      split3_store8(img, r, c0, C, y);
    }
  }
}
void launch_vocos_ln_split3(const float* x, int64_t rows, const float* gamma, const float* beta, void* img, cudaStream_t st) {
  vocos_ln_split3_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(x, rows, gamma, beta, (uint8_t*)img);
}

// head_out [rows, ldh] = [log-mag(641) | phase(641) | pad] -> split-3 image of [re(641) | im(641) | 0 .. ] with K = 1344
__global__ void __launch_bounds__(192) vocos_spectrum_split3_kernel(const float* __restrict__ head_out, int ldh, uint8_t* __restrict__ img) {
  constexpr int K = 1344;
  const int64_t r = blockIdx.x;
  const float* h = head_out + r * ldh;
  const int chunk = threadIdx.x;   // 168 chunks of 8 columns
  if (chunk >= K / 8) return;
  float x[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = chunk * 8 + i;
    float v = 0.f;
    if (c < 2 * VC_BINS) {
      const int k = (c < VC_BINS) ? c : c - VC_BINS;
      const float mag = fminf(expf(h[k]), 100.0f);
      const float ph = h[VC_BINS + k];
      v = mag * ((c < VC_BINS) ? cosf(ph) : sinf(ph));
    }
    x[i] = v;
  }
  split3_store8(img, r, chunk * 8, K, x);
}
void launch_vocos_spectrum_split3(const float* head_out, int ldh, int64_t rows, void* img, cudaStream_t st) {
  vocos_spectrum_split3_kernel<<<(unsigned)rows, 192, 0, st>>>(head_out, ldh, (uint8_t*)img);
}

// frames [B*T,1280] (irfft output) * hann, overlap-add with hop 320, trim pad=480 each side, divide by
// the window-square envelope (vocos spectral_ops.py ISTFT.forward, padding == "same").
__global__ void __launch_bounds__(256) vocos_overlap_add_kernel(const float* __restrict__ frames, int B, int T,
                                                                const float* __restrict__ window,
                                                                float* __restrict__ wave) {
  const int b = blockIdx.y;
  const int s = blockIdx.x * 256 + threadIdx.x;
  const int nout = T * VC_HOP;
  if (s >= nout) return;
  const int sp = s + (VC_NFFT - VC_HOP) / 2;  // index in the untrimmed signal
  int t_hi = sp / VC_HOP;
  if (t_hi > T - 1) t_hi = T - 1;
  int t_lo = (sp - VC_NFFT + VC_HOP) / VC_HOP;  // ceil((sp-1279)/320)
  if (t_lo < 0) t_lo = 0;
  float acc = 0.f, env = 0.f;
  for (int t = t_lo; t <= t_hi; ++t) {
    const int n = sp - t * VC_HOP;
    if (n >= 0 && n < VC_NFFT) {
      const float w = window[n];
      acc = __fadd_rn(acc, __fmul_rn(frames[((size_t)b * T + t) * VC_NFFT + n], w));
      env = __fadd_rn(env, __fmul_rn(w, w));
    }
  }
  wave[(size_t)b * nout + s] = acc / env;
}
void launch_vocos_overlap_add(const float* frames, int B, int T, const float* window, float* wave, cudaStream_t st) {
  vocos_overlap_add_kernel<<<dim3((T * VC_HOP + 255) / 256, B), 256, 0, st>>>(frames, B, T, window, wave);
}

}  // namespace vx
