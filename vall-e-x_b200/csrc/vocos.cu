// Vocos decode kernels (third-party `vocos` package semantics; see oracle/vocos_oracle.py header).
// Activations are time-major [B*T, C] fp32; every contraction (k=7 embed conv via im2col, the two
// point-wise linears, the head and the inverse DFT expressed as a [1280 x 1282] basis contraction) runs
// on the fp32 SIMT GEMM (gemm_simt.cu) because the waveform tolerance (RMS 1e-4) rules out bf16 operands.
#include "common.cuh"
#include "kernels.h"

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
