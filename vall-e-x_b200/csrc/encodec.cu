// EnCodec 24 kHz encoder + residual vector quantiser on the B200: the audio half of prompt making.
// Replaces `AudioTokenizer.encode` / `tokenize_audio` (reference data/tokenizer.py:63-111) as used by
// `make_prompt` (utils/prompt_making.py:57-84).  The arithmetic is the third-party `encodec` package's SEANet encoder
// (causal reflect-padded convolutions, ELU, 2-layer LSTM with skip) and Euclidean RVQ; oracle/encodec_oracle.py restates it
// and is pinned to transformers' port of the same model.
//
// Everything is fp32 on CUDA cores: the outputs are integer codebook indices that have to reproduce the reference's argmax
// decisions, a 15-second prompt is ~45 GFLOP, and the sequential LSTM (T <= 1125 dependent steps) bounds the latency anyway.
//   conv1d_kernel     direct convolution, [C][T] activations, output tile CO_T x T_T per CTA, 4x4 outputs per thread;
//                     the input window and the weight slab of 8 input channels are staged in shared memory; ELU of the
//                     INPUT, the causal reflect padding (as an index map), bias and the residual add are fused.
//   lstm_kernel       one cooperative launch for both layers: 64 CTAs per layer, a CTA keeps the 32 gate rows of its 8 hidden
//                     units in shared memory (layer 1: [W_ih | W_hh], 128 KB) for the whole sequence; layer 1 runs one
//                     step behind layer 0, so the sequence costs T+1 grid barriers instead of 2T.
//   rvq_kernel        one CTA per frame walks the 8 codebooks: distances in the reference's operation order
//                     -(|r|^2 - 2 r.e + |e|^2), first maximum, residual update.
#include <cooperative_groups.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vallex_b200.h"
#include "common.cuh"

using namespace vx;

int vx_fail(int code, const std::string& msg);  // engine.cu: sets the thread-local vx_last_error() string

namespace {

#define ECK(call)                                                                                                     \
  do {                                                                                                                \
    cudaError_t _e = (call);                                                                                          \
    if (_e != cudaSuccess)                                                                                            \
      return vx_fail(VX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e) + " @encodec:" + std::to_string(__LINE__)); \
  } while (0)

constexpr int kHop = 320;
constexpr int kLstmH = 512;
constexpr int kEncDim = 128;
constexpr int kBins = 1024;
constexpr int kMaxQ = 32;
constexpr int kCiChunk = 8;

__device__ __forceinline__ float elu1(float v) { return v > 0.f ? v : expm1f(v); }

// Index map of `pad1d(x, (left, right), 'reflect')` incl. the short-input branch (zeros appended up to max_pad + 1 samples
// before reflecting; encodec/modules/conv.py pad1d): p = position in the padded signal -> source index, or -1 for an appended zero.
__device__ __forceinline__ int reflect_src(int p, int left, int len, int len_ext) {
  int q = p - left;
  if (q < 0) q = -q;
  if (q >= len_ext) q = 2 * (len_ext - 1) - q;
  return q < len ? q : -1;
}

struct ConvArgs {
  const float* x;      // [ci][len]
  const float* w;      // [co][ci][k]
  const float* bias;   // [co]
  const float* add;    // nullable [co][tout]: out = conv + add
  float* out;          // [co][tout]
  int ci, co, k, stride, len, tout, left, len_ext, elu_in;
};

// CO_T output channels x T_T = (256 / (CO_T/4)) * 4 output positions per CTA; 256 threads, 4 x 4 outputs per thread.
template <int CO_T>
__global__ void __launch_bounds__(256) conv1d_kernel(ConvArgs a) {
  constexpr int TCO = CO_T / 4, TT = 256 / TCO, T_T = TT * 4;
  extern __shared__ float smem[];
  const int win = (T_T - 1) * a.stride + a.k;          // input window of the tile
  const int wq = (win + a.stride - 1) / a.stride, winp = wq * a.stride;
  float* xs = smem;                                     // [kCiChunk][stride phase][wq]: position p at (p % stride) * wq + p / stride,
  float* ws = smem + kCiChunk * winp;                   //   so the lanes of a warp (consecutive outputs) read consecutive words
                                                        // ws: [kCiChunk][k][CO_T]
  const int t0 = blockIdx.x * T_T, co0 = blockIdx.y * CO_T;
  const int tx = threadIdx.x % TT, ty = threadIdx.x / TT;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int p0 = t0 * a.stride;                         // first padded position of the window
  for (int c0 = 0; c0 < a.ci; c0 += kCiChunk) {
    const int nc = min(kCiChunk, a.ci - c0);
    __syncthreads();
    for (int i = threadIdx.x; i < nc * win; i += 256) {
      const int c = i / win, p = i - c * win;
      const int src = reflect_src(p0 + p, a.left, a.len, a.len_ext);
      float v = 0.f;
      if (src >= 0) {
        v = __ldg(a.x + (size_t)(c0 + c) * a.len + src);
        if (a.elu_in) v = elu1(v);
      }
      xs[c * winp + (p % a.stride) * wq + p / a.stride] = v;
    }
    for (int i = threadIdx.x; i < nc * a.k * CO_T; i += 256) {
      const int o = i % CO_T, ck = i / CO_T, c = ck / a.k, kk = ck - c * a.k;
      const int co = co0 + o;
      ws[i] = co < a.co ? __ldg(a.w + ((size_t)co * a.ci + c0 + c) * a.k + kk) : 0.f;
    }
    __syncthreads();
    for (int c = 0; c < nc; ++c) {
      const float* xr = xs + c * winp + tx;                 // thread's outputs: tx + j*TT
      const float* wr = ws + c * a.k * CO_T + ty * 4;
      int ph = 0, off = 0;                                  // tap kk = off * stride + ph
      for (int kk = 0; kk < a.k; ++kk) {
        const float4 wv = *reinterpret_cast<const float4*>(wr + kk * CO_T);
        float xv[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) xv[j] = xr[ph * wq + off + j * TT];
        if (++ph == a.stride) { ph = 0; ++off; }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[0][j] = fmaf(wv.x, xv[j], acc[0][j]);
          acc[1][j] = fmaf(wv.y, xv[j], acc[1][j]);
          acc[2][j] = fmaf(wv.z, xv[j], acc[2][j]);
          acc[3][j] = fmaf(wv.w, xv[j], acc[3][j]);
        }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty * 4 + i;
    if (co >= a.co) continue;
    const float b = a.bias ? __ldg(a.bias + co) : 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int t = t0 + tx + j * TT;
      if (t >= a.tout) continue;
      float v = acc[i][j] + b;
      if (a.add) v += a.add[(size_t)co * a.tout + t];
      a.out[(size_t)co * a.tout + t] = v;
    }
  }
}

// ---- LSTM ------------------------------------------------------------------------------------------------------------
struct LstmArgs {
  const float* pre0;     // [4H][T]  W_ih0 x_t + b_ih0 + b_hh0 (conv1d_kernel, k = 1)
  const float* w_hh0;    // [4H][H]
  const float* w_ih1;    // [4H][H]
  const float* w_hh1;    // [4H][H]
  const float* bias1;    // [4H]     b_ih1 + b_hh1
  const float* skip;     // [H][T]   the LSTM block's input (SLSTM adds it back)
  float* h0;             // [T][H]   layer-0 outputs
  float* h1;             // [T][H]   layer-1 outputs
  float* out;            // [H][T]   h1 + skip, conv layout
  unsigned int* bar;
  int T;
};
constexpr int kLstmCtasPerLayer = 64;
constexpr int kLstmUnits = kLstmH / kLstmCtasPerLayer;  // 8 hidden units = 32 gate rows per CTA
constexpr int kLstmThreads = 256;

__device__ __forceinline__ float sigmoidf_(float v) { return 1.f / (1.f + expf(-v)); }

__global__ void __launch_bounds__(kLstmThreads) lstm_kernel(LstmArgs a) {
  extern __shared__ float smem[];
  const int layer = blockIdx.x / kLstmCtasPerLayer, cta = blockIdx.x % kLstmCtasPerLayer;
  const int K = layer == 0 ? kLstmH : 2 * kLstmH;
  float* wsm = smem;                       // [32][K] rows: gate g (i,f,g,o) of unit u at row g*8+u
  float* vin = smem + 32 * 2 * kLstmH;     // [K] input vector of the step
  __shared__ float gates[32];
  __shared__ float cstate[kLstmUnits];
  const int u0 = cta * kLstmUnits;
  for (int i = threadIdx.x; i < 32 * K; i += kLstmThreads) {
    const int r = i / K, k = i - r * K;
    const int row = (r / kLstmUnits) * kLstmH + u0 + (r % kLstmUnits);
    float v;
    if (layer == 0) v = a.w_hh0[(size_t)row * kLstmH + k];
    else v = k < kLstmH ? a.w_ih1[(size_t)row * kLstmH + k] : a.w_hh1[(size_t)row * kLstmH + k - kLstmH];
    wsm[i] = v;
  }
  if (threadIdx.x < kLstmUnits) cstate[threadIdx.x] = 0.f;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  unsigned int epoch = 0;
  // global step s: layer 0 computes t = s (s < T), layer 1 computes t = s - 1 (1 <= s <= T)
  for (int s = 0; s <= a.T; ++s) {
    const int t = s - layer;
    const bool active = t >= 0 && t < a.T;
    if (active) {
      // input vector: layer 0: h0[t-1]; layer 1: [h0[t] | h1[t-1]]
      for (int k = threadIdx.x; k < K; k += kLstmThreads) {
        float v;
        if (layer == 0) v = t > 0 ? __ldcg(a.h0 + (size_t)(t - 1) * kLstmH + k) : 0.f;
        else if (k < kLstmH) v = __ldcg(a.h0 + (size_t)t * kLstmH + k);
        else v = t > 0 ? __ldcg(a.h1 + (size_t)(t - 1) * kLstmH + k - kLstmH) : 0.f;
        vin[k] = v;
      }
      __syncthreads();
      // 8 warps x 4 rows
#pragma unroll
      for (int rr = 0; rr < 4; ++rr) {
        const int r = warp * 4 + rr;
        const float* wr = wsm + (size_t)r * K;
        float sacc = 0.f;
        for (int k = lane * 4; k < K; k += 128) {
          const float4 wv = *reinterpret_cast<const float4*>(wr + k);
          const float4 xv = *reinterpret_cast<const float4*>(vin + k);
          sacc = fmaf(wv.x, xv.x, sacc); sacc = fmaf(wv.y, xv.y, sacc);
          sacc = fmaf(wv.z, xv.z, sacc); sacc = fmaf(wv.w, xv.w, sacc);
        }
        sacc = warp_sum(sacc);
        if (lane == 0) {
          const int row = (r / kLstmUnits) * kLstmH + u0 + (r % kLstmUnits);
          gates[r] = sacc + (layer == 0 ? __ldg(a.pre0 + (size_t)row * a.T + t) : __ldg(a.bias1 + row));
        }
      }
      __syncthreads();
      if (threadIdx.x < kLstmUnits) {
        const int u = threadIdx.x;
        const float ig = sigmoidf_(gates[u]), fg = sigmoidf_(gates[kLstmUnits + u]);
        const float gg = tanhf(gates[2 * kLstmUnits + u]), og = sigmoidf_(gates[3 * kLstmUnits + u]);
        const float c = fg * cstate[u] + ig * gg;
        cstate[u] = c;
        const float h = og * tanhf(c);
        if (layer == 0) a.h0[(size_t)t * kLstmH + u0 + u] = h;
        else {
          a.h1[(size_t)t * kLstmH + u0 + u] = h;
          a.out[(size_t)(u0 + u) * a.T + t] = h + a.skip[(size_t)(u0 + u) * a.T + t];
        }
      }
    }
    // grid barrier (cooperative launch: all CTAs are co-resident)
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch += gridDim.x;
      __threadfence();
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
      unsigned int v, spins = 0;
      do {
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.bar) : "memory");
        if (++spins > (1u << 26)) __trap();
      } while ((int)(v - epoch) < 0);
    }
    __syncthreads();
  }
}

// ---- RVQ ---------------------------------------------------------------------------------------------------------------
// e2[q][j] = sum_d embed[q][j][d]^2 in the reference's reduction shape (`embed.pow(2).sum(0)` over d)
__global__ void rvq_norm_kernel(const float* __restrict__ embed, float* __restrict__ e2, int rows) {
  const int r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float s = 0.f;
  for (int d = 0; d < kEncDim; ++d) { const float v = embed[(size_t)r * kEncDim + d]; s = __fadd_rn(s, __fmul_rn(v, v)); }
  e2[r] = s;
}

__global__ void __launch_bounds__(256) rvq_kernel(const float* __restrict__ emb /*[128][T]*/, const float* __restrict__ embed /*[nq][1024][128]*/,
                                                   const float* __restrict__ e2, int T, int nq, int32_t* __restrict__ codes /*[nq][T]*/) {
  __shared__ float res[kEncDim];
  __shared__ float redv[8];
  __shared__ int redi[8];
  __shared__ float red[32];
  __shared__ int best_s;
  const int t = blockIdx.x;
  if (threadIdx.x < kEncDim) res[threadIdx.x] = emb[(size_t)threadIdx.x * T + t];
  __syncthreads();
  for (int q = 0; q < nq; ++q) {
    const float* eq = embed + (size_t)q * kBins * kEncDim;
    const float s = block_sum(threadIdx.x < kEncDim ? __fmul_rn(res[threadIdx.x], res[threadIdx.x]) : 0.f, red);
    float bv = -INFINITY;
    int bi = 0;
    for (int j = threadIdx.x; j < kBins; j += 256) {     // ascending j per thread: strict > keeps the first maximum
      const float4* er = reinterpret_cast<const float4*>(eq + (size_t)j * kEncDim);
      float d = 0.f;
#pragma unroll 8
      for (int k = 0; k < kEncDim / 4; ++k) {
        const float4 ev = __ldg(er + k);
        d = fmaf(res[4 * k], ev.x, d); d = fmaf(res[4 * k + 1], ev.y, d);
        d = fmaf(res[4 * k + 2], ev.z, d); d = fmaf(res[4 * k + 3], ev.w, d);
      }
      // dist = -(|r|^2 - 2 r.e + |e|^2), each operation rounded separately like the reference's tensor expression
      const float dist = -__fadd_rn(__fsub_rn(s, __fmul_rn(2.f, d)), __ldg(e2 + q * kBins + j));
      if (dist > bv) { bv = dist; bi = j; }
    }
    // block argmax, ties -> smallest index
    for (int o = 16; o > 0; o >>= 1) {
      const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { redv[threadIdx.x >> 5] = bv; redi[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float v = redv[0]; int i = redi[0];
      for (int w = 1; w < 8; ++w) if (redv[w] > v || (redv[w] == v && redi[w] < i)) { v = redv[w]; i = redi[w]; }
      best_s = i;
      codes[(size_t)q * T + t] = i;
    }
    __syncthreads();
    if (threadIdx.x < kEncDim) res[threadIdx.x] = __fsub_rn(res[threadIdx.x], __ldg(eq + (size_t)best_s * kEncDim + threadIdx.x));
    __syncthreads();
  }
}

__global__ void check_finite_kernel(const float* __restrict__ x, size_t n, int* flag) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  bool bad = false;
  for (; i < n; i += (size_t)gridDim.x * blockDim.x) bad |= !isfinite(x[i]);
  if (bad) atomicOr(flag, 1);
}

struct ConvW {
  std::string prefix;
  int co, ci, k, stride;
  float* w = nullptr;
  float* b = nullptr;
};

}  // namespace

struct vx_encodec {
  int device = 0;
  cudaStream_t stream = nullptr;
  std::map<std::string, std::vector<float>> host;          // staged tensors until finalize
  std::map<std::string, std::vector<int64_t>> shapes;
  std::vector<ConvW> convs;                                 // execution order (weights.encodec_conv_specs)
  float *w_ih0 = nullptr, *w_hh0 = nullptr, *b0 = nullptr, *w_ih1 = nullptr, *w_hh1 = nullptr, *b1 = nullptr;
  float *embed = nullptr, *e2 = nullptr;
  int n_codebooks = 0;
  bool finalized = false;
  // workspaces
  float *buf[3] = {nullptr, nullptr, nullptr};
  size_t buf_elems = 0;
  float *pre0 = nullptr, *h0 = nullptr, *h1 = nullptr, *emb = nullptr, *wav = nullptr;
  int32_t* codes = nullptr;
  int64_t cap_samples = 0;
  unsigned int* bar = nullptr;
  int* flag = nullptr;
  int64_t last_T = 0;
  int64_t launches = 0;
  float last_ms = 0.f;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

int conv_launch(vx_encodec* e, const ConvW& c, const float* x, int len, const float* add, float* out, int elu_in, int* tout_out,
                cudaStream_t st) {
  ConvArgs a;
  a.x = x; a.w = c.w; a.bias = c.b; a.add = add; a.out = out;
  a.ci = c.ci; a.co = c.co; a.k = c.k; a.stride = c.stride; a.len = len; a.elu_in = elu_in;
  // SConv1d (causal): padding_total = k - stride on the left; on the right whatever completes the last frame
  a.left = c.k - c.stride;
  a.tout = (len + c.stride - 1) / c.stride;
  const int right = a.tout * c.stride - len;
  const int max_pad = std::max(a.left, right);
  a.len_ext = len <= max_pad ? max_pad + 1 : len;
  *tout_out = a.tout;
  const int co_t = c.co <= 16 ? 16 : (c.co <= 32 ? 32 : 64);
  const int tco = co_t / 4, tt = 256 / tco, t_t = tt * 4;
  const int win = (t_t - 1) * c.stride + c.k;
  const int winp = (win + c.stride - 1) / c.stride * c.stride;
  const size_t smem = (size_t)(kCiChunk * winp + kCiChunk * c.k * co_t) * sizeof(float);
  dim3 grid((a.tout + t_t - 1) / t_t, (c.co + co_t - 1) / co_t);
  if (co_t == 16) conv1d_kernel<16><<<grid, 256, smem, st>>>(a);
  else if (co_t == 32) conv1d_kernel<32><<<grid, 256, smem, st>>>(a);
  else conv1d_kernel<64><<<grid, 256, smem, st>>>(a);
  e->launches++;
  ECK(cudaGetLastError());
  return VX_OK;
}

int upload(const std::vector<float>& h, float** d) {
  ECK(cudaMalloc(d, h.size() * sizeof(float)));
  ECK(cudaMemcpy(*d, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
  return VX_OK;
}

// (prefix, co, ci, k, stride) in execution order -- the encodec_24khz SEANet encoder (ratios [8,5,4,2] reversed)
std::vector<ConvW> conv_specs() {
  std::vector<ConvW> v;
  auto add = [&](const std::string& p, int co, int ci, int k, int s) { ConvW c; c.prefix = p; c.co = co; c.ci = ci; c.k = k; c.stride = s; v.push_back(c); };
  add("encoder.model.0.conv.conv", 32, 1, 7, 1);
  const int ratios[4] = {2, 4, 5, 8};
  int dim = 32, idx = 1;
  for (int r : ratios) {
    const std::string m = "encoder.model." + std::to_string(idx);
    add(m + ".block.1.conv.conv", dim / 2, dim, 3, 1);
    add(m + ".block.3.conv.conv", dim, dim / 2, 1, 1);
    add(m + ".shortcut.conv.conv", dim, dim, 1, 1);
    add("encoder.model." + std::to_string(idx + 2) + ".conv.conv", 2 * dim, dim, 2 * r, r);
    dim *= 2; idx += 3;
  }
  add("encoder.model." + std::to_string(idx + 2) + ".conv.conv", kEncDim, dim, 7, 1);
  return v;
}

}  // namespace

extern "C" {

int vx_encodec_create(int device, vx_encodec** out) {
  if (!out) return vx_fail(VX_ERR_ARG, "vx_encodec_create: out is null");
  cudaDeviceProp prop;
  ECK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return vx_fail(VX_ERR_STATE, "vx_encodec_create: this library targets sm_100a (B200) only; no CPU or other-GPU fallback");
  ECK(cudaSetDevice(device));
  vx_encodec* e = new vx_encodec();
  e->device = device;
  ECK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  ECK(cudaEventCreate(&e->ev0));
  ECK(cudaEventCreate(&e->ev1));
  ECK(cudaMalloc(&e->bar, 256));
  ECK(cudaMemset(e->bar, 0, 256));
  ECK(cudaMalloc(&e->flag, sizeof(int)));
  e->convs = conv_specs();
  *out = e;
  return VX_OK;
}

void vx_encodec_destroy(vx_encodec* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  for (auto& c : e->convs) { cudaFree(c.w); cudaFree(c.b); }
  float* ptrs[] = {e->w_ih0, e->w_hh0, e->b0, e->w_ih1, e->w_hh1, e->b1, e->embed, e->e2, e->buf[0], e->buf[1], e->buf[2],
                   e->pre0, e->h0, e->h1, e->emb, e->wav};
  for (float* p : ptrs) cudaFree(p);
  cudaFree(e->codes); cudaFree(e->bar); cudaFree(e->flag);
  cudaEventDestroy(e->ev0); cudaEventDestroy(e->ev1);
  cudaStreamDestroy(e->stream);
  delete e;
}

int vx_encodec_load_weight(vx_encodec* e, const char* key, const float* data, const int64_t* shape, int ndim) {
  if (!e || !key || !data || !shape || ndim < 0 || ndim > 4) return vx_fail(VX_ERR_ARG, "vx_encodec_load_weight: bad argument");
  if (e->finalized) return vx_fail(VX_ERR_STATE, "vx_encodec_load_weight: weights already finalized");
  int64_t n = 1;
  for (int i = 0; i < ndim; ++i) n *= shape[i];
  e->host[key].assign(data, data + n);
  e->shapes[key].assign(shape, shape + ndim);
  return VX_OK;
}

int vx_encodec_finalize(vx_encodec* e) {
  if (!e) return vx_fail(VX_ERR_ARG, "vx_encodec_finalize: null engine");
  if (e->finalized) return VX_OK;
  ECK(cudaSetDevice(e->device));
  auto need = [&](const std::string& k, std::vector<int64_t> shp) -> const std::vector<float>* {
    auto it = e->host.find(k);
    if (it == e->host.end()) return nullptr;
    if (e->shapes[k] != shp) return nullptr;
    return &it->second;
  };
  for (auto& c : e->convs) {
    // plain weight (the reference strips weight norm, data/tokenizer.py:31-60) or weight_g / weight_v: w = g * v / ||v||
    std::vector<float> w;
    if (auto p = need(c.prefix + ".weight", {c.co, c.ci, c.k})) w = *p;
    else {
      auto g = need(c.prefix + ".weight_g", {c.co, 1, 1});
      auto v = need(c.prefix + ".weight_v", {c.co, c.ci, c.k});
      if (!g || !v) return vx_fail(VX_ERR_STATE, "vx_encodec_finalize: missing or mis-shaped " + c.prefix + ".weight (or weight_g/weight_v)");
      w.resize(v->size());
      const size_t per = (size_t)c.ci * c.k;
      for (int o = 0; o < c.co; ++o) {
        double s = 0.0;
        for (size_t i = 0; i < per; ++i) s += (double)(*v)[o * per + i] * (*v)[o * per + i];
        const float scale = (*g)[o] / (float)std::sqrt(s);
        for (size_t i = 0; i < per; ++i) w[o * per + i] = (*v)[o * per + i] * scale;
      }
    }
    auto b = need(c.prefix + ".bias", {c.co});
    if (!b) return vx_fail(VX_ERR_STATE, "vx_encodec_finalize: missing or mis-shaped " + c.prefix + ".bias");
    if (int rc = upload(w, &c.w)) return rc;
    if (int rc = upload(*b, &c.b)) return rc;
  }
  const std::string lp = "encoder.model.13.lstm.";
  const int64_t H = kLstmH;
  struct { const char* n; std::vector<int64_t> s; float** d; } mats[] = {
      {"weight_ih_l0", {4 * H, H}, &e->w_ih0}, {"weight_hh_l0", {4 * H, H}, &e->w_hh0},
      {"weight_ih_l1", {4 * H, H}, &e->w_ih1}, {"weight_hh_l1", {4 * H, H}, &e->w_hh1}};
  for (auto& m : mats) {
    auto p = need(lp + m.n, m.s);
    if (!p) return vx_fail(VX_ERR_STATE, std::string("vx_encodec_finalize: missing or mis-shaped ") + lp + m.n);
    if (int rc = upload(*p, m.d)) return rc;
  }
  for (int l = 0; l < 2; ++l) {
    auto bi = need(lp + "bias_ih_l" + std::to_string(l), {4 * H});
    auto bh = need(lp + "bias_hh_l" + std::to_string(l), {4 * H});
    if (!bi || !bh) return vx_fail(VX_ERR_STATE, "vx_encodec_finalize: missing or mis-shaped LSTM bias of layer " + std::to_string(l));
    std::vector<float> s(4 * H);
    for (int i = 0; i < 4 * H; ++i) s[i] = (*bi)[i] + (*bh)[i];
    if (int rc = upload(s, l == 0 ? &e->b0 : &e->b1)) return rc;
  }
  std::vector<float> tables;
  int q = 0;
  for (; q < kMaxQ; ++q) {
    auto p = need("quantizer.vq.layers." + std::to_string(q) + "._codebook.embed", {kBins, kEncDim});
    if (!p) break;
    tables.insert(tables.end(), p->begin(), p->end());
  }
  if (q == 0) return vx_fail(VX_ERR_STATE, "vx_encodec_finalize: no quantizer.vq.layers.<q>._codebook.embed [1024,128] tables");
  e->n_codebooks = q;
  if (int rc = upload(tables, &e->embed)) return rc;
  ECK(cudaMalloc(&e->e2, (size_t)q * kBins * sizeof(float)));
  rvq_norm_kernel<<<(q * kBins + 255) / 256, 256, 0, e->stream>>>(e->embed, e->e2, q * kBins);
  ECK(cudaGetLastError());
  ECK(cudaStreamSynchronize(e->stream));
  ECK(cudaFuncSetAttribute(lstm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (32 * 2 * kLstmH + 2 * kLstmH) * (int)sizeof(float)));
  ECK(cudaFuncSetAttribute(conv1d_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  ECK(cudaFuncSetAttribute(conv1d_kernel<32>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  ECK(cudaFuncSetAttribute(conv1d_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
  e->host.clear();
  e->shapes.clear();
  e->finalized = true;
  return VX_OK;
}

int vx_encodec_num_codebooks(vx_encodec* e) { return e ? e->n_codebooks : 0; }

/* frames of an n-sample waveform: every strided convolution rounds up (right padding completes the last frame) */
int64_t vx_encodec_frames(int64_t n_samples) {
  int64_t t = n_samples;
  const int ratios[4] = {2, 4, 5, 8};
  for (int r : ratios) t = (t + r - 1) / r;
  return t;
}

static int ensure_capacity(vx_encodec* e, int64_t n) {
  if (n <= e->cap_samples) return VX_OK;
  float** all[] = {&e->buf[0], &e->buf[1], &e->buf[2], &e->pre0, &e->h0, &e->h1, &e->emb, &e->wav};
  for (auto p : all) { cudaFree(*p); *p = nullptr; }
  cudaFree(e->codes); e->codes = nullptr;
  const int64_t T = vx_encodec_frames(n);
  // largest activation: 32 channels at the input rate (later stages: 64 x N/2, 128 x N/8, ...)
  e->buf_elems = (size_t)32 * n + 4096;
  for (int i = 0; i < 3; ++i) ECK(cudaMalloc(&e->buf[i], e->buf_elems * sizeof(float)));
  ECK(cudaMalloc(&e->wav, (size_t)n * sizeof(float)));
  ECK(cudaMalloc(&e->pre0, (size_t)4 * kLstmH * T * sizeof(float)));
  ECK(cudaMalloc(&e->h0, (size_t)kLstmH * T * sizeof(float)));
  ECK(cudaMalloc(&e->h1, (size_t)kLstmH * T * sizeof(float)));
  ECK(cudaMalloc(&e->emb, (size_t)kEncDim * T * sizeof(float)));
  ECK(cudaMalloc(&e->codes, (size_t)kMaxQ * T * sizeof(int32_t)));
  e->cap_samples = n;
  return VX_OK;
}

int vx_encodec_encode(vx_encodec* e, const float* wav, int64_t n_samples, int n_q, int32_t* out_codes, int mem, void* stream) {
  if (!e || !wav || !out_codes) return vx_fail(VX_ERR_ARG, "vx_encodec_encode: null argument");
  if (!e->finalized) return vx_fail(VX_ERR_STATE, "vx_encodec_encode: call vx_encodec_finalize first");
  if (n_samples < 1 || n_samples > (int64_t)1 << 26) return vx_fail(VX_ERR_ARG, "vx_encodec_encode: n_samples out of range");
  if (n_q < 1 || n_q > e->n_codebooks) return vx_fail(VX_ERR_ARG, "vx_encodec_encode: n_q must be in [1, loaded codebooks]");
  ECK(cudaSetDevice(e->device));
  cudaStream_t st = stream ? (cudaStream_t)stream : e->stream;
  if (int rc = ensure_capacity(e, n_samples)) return rc;
  const int64_t T = vx_encodec_frames(n_samples);
  e->launches = 0;
  ECK(cudaEventRecord(e->ev0, st));
  ECK(cudaMemsetAsync(e->flag, 0, sizeof(int), st));
  if (mem == VX_HOST) ECK(cudaMemcpyAsync(e->wav, wav, (size_t)n_samples * sizeof(float), cudaMemcpyHostToDevice, st));
  else ECK(cudaMemcpyAsync(e->wav, wav, (size_t)n_samples * sizeof(float), cudaMemcpyDeviceToDevice, st));
  check_finite_kernel<<<148, 256, 0, st>>>(e->wav, (size_t)n_samples, e->flag);
  e->launches++;
  // SEANet encoder: conv, 4 x [resnet block (ELU, k3, ELU, k1; + shortcut k1), ELU, strided conv], LSTM, ELU, conv
  float *x = e->buf[0], *y = e->buf[1], *z = e->buf[2];
  int len = (int)n_samples, tout = 0;
  size_t ci = 0;
  if (int rc = conv_launch(e, e->convs[ci++], e->wav, len, nullptr, x, 0, &tout, st)) return rc;
  for (int s = 0; s < 4; ++s) {
    if (int rc = conv_launch(e, e->convs[ci++], x, len, nullptr, y, 1, &tout, st)) return rc;     // block.1: ELU -> k3
    if (int rc = conv_launch(e, e->convs[ci++], y, len, nullptr, z, 1, &tout, st)) return rc;     // block.3: ELU -> k1
    if (int rc = conv_launch(e, e->convs[ci++], x, len, z, y, 0, &tout, st)) return rc;           // shortcut(x) + block(x)
    if (int rc = conv_launch(e, e->convs[ci++], y, len, nullptr, x, 1, &tout, st)) return rc;     // ELU -> strided conv
    len = tout;
  }
  if (len != T) return vx_fail(VX_ERR_STATE, "vx_encodec_encode: internal frame-count mismatch");
  // LSTM: input projection of layer 0 for all t as a k=1 convolution, then the recurrent kernel
  ConvW proj; proj.co = 4 * kLstmH; proj.ci = kLstmH; proj.k = 1; proj.stride = 1; proj.w = e->w_ih0; proj.b = e->b0;
  if (int rc = conv_launch(e, proj, x, len, nullptr, e->pre0, 0, &tout, st)) return rc;
  LstmArgs la;
  la.pre0 = e->pre0; la.w_hh0 = e->w_hh0; la.w_ih1 = e->w_ih1; la.w_hh1 = e->w_hh1; la.bias1 = e->b1; la.skip = x;
  la.h0 = e->h0; la.h1 = e->h1; la.out = y; la.bar = e->bar; la.T = (int)T;
  ECK(cudaMemsetAsync(e->bar, 0, sizeof(unsigned int), st));
  void* kargs[] = {&la};
  const size_t lsm = (size_t)(32 * 2 * kLstmH + 2 * kLstmH) * sizeof(float);
  ECK(cudaLaunchCooperativeKernel((void*)lstm_kernel, dim3(2 * kLstmCtasPerLayer), dim3(kLstmThreads), kargs, lsm, st));
  e->launches++;
  if (int rc = conv_launch(e, e->convs[ci++], y, len, nullptr, e->emb, 1, &tout, st)) return rc;  // ELU -> final conv
  rvq_kernel<<<(unsigned)T, 256, 0, st>>>(e->emb, e->embed, e->e2, (int)T, n_q, e->codes);
  e->launches++;
  ECK(cudaGetLastError());
  if (mem == VX_HOST) ECK(cudaMemcpyAsync(out_codes, e->codes, (size_t)n_q * T * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  else ECK(cudaMemcpyAsync(out_codes, e->codes, (size_t)n_q * T * sizeof(int32_t), cudaMemcpyDeviceToDevice, st));
  int flag = 0;
  ECK(cudaMemcpyAsync(&flag, e->flag, sizeof(int), cudaMemcpyDeviceToHost, st));
  ECK(cudaEventRecord(e->ev1, st));
  ECK(cudaStreamSynchronize(st));
  ECK(cudaEventElapsedTime(&e->last_ms, e->ev0, e->ev1));
  e->last_T = T;
  if (flag) return vx_fail(VX_ERR_ARG, "vx_encodec_encode: waveform contains non-finite samples");
  return VX_OK;
}

int vx_encodec_embeddings(vx_encodec* e, float* dst, int64_t max_elems, int64_t* out_elems) {
  if (!e || !dst || !e->emb || e->last_T == 0) return vx_fail(VX_ERR_STATE, "vx_encodec_embeddings: nothing encoded yet");
  const int64_t n = (int64_t)kEncDim * e->last_T;
  if (out_elems) *out_elems = n;
  if (n > max_elems) return vx_fail(VX_ERR_CAPACITY, "vx_encodec_embeddings: destination too small");
  ECK(cudaSetDevice(e->device));
  ECK(cudaMemcpy(dst, e->emb, (size_t)n * sizeof(float), cudaMemcpyDeviceToHost));
  return VX_OK;
}

int vx_encodec_last_timing(vx_encodec* e, float* ms, int64_t* frames, int64_t* kernel_launches) {
  if (!e) return vx_fail(VX_ERR_ARG, "vx_encodec_last_timing: null engine");
  if (ms) *ms = e->last_ms;
  if (frames) *frames = e->last_T;
  if (kernel_launches) *kernel_launches = e->launches;
  return VX_OK;
}

}  // extern "C"
