// vallex_b200 engine: C ABI (include/vallex_b200.h), weight ingestion, workspaces and the host-side
// orchestration of the AR prefill / AR decode / NAR / Vocos kernel sequences.  One engine per GPU.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <vector>

#include "../../include/vallex_b200.h"
#include "common.cuh"
#include "decode_persist.h"
#include "kernels.h"
#include "tc_common.cuh"

using namespace vx;

namespace {

thread_local std::string g_err;
int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
}  // namespace
int vx_fail(int code, const std::string& msg) { return fail(code, msg); }  // for the other translation units (encodec.cu)
namespace {
#define CK(call)                                                                                   \
  do {                                                                                             \
    cudaError_t _e = (call);                                                                       \
    if (_e != cudaSuccess)                                                                         \
      return fail(VX_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(_e) + " @" + std::to_string(__LINE__)); \
  } while (0)

__global__ void f32_to_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) out[i] = __float2bfloat16_rn(in[i]);
}
// out[c][r] = in[r][c] (bf16), 32x32 tiles through shared memory
__global__ void transpose_bf16_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int rows, int cols) {
  __shared__ bf16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    if (r < rows && c < cols) tile[i][threadIdx.x] = in[(size_t)r * cols + c];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < rows && c < cols) out[(size_t)c * rows + r] = tile[threadIdx.x][i];
  }
}
__global__ void write_cb0_kernel(Batch bt, const int32_t* __restrict__ cb0, const int32_t* __restrict__ tl, int Tmax) {
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (threadIdx.x == 0 && blockIdx.x == 0) bt.T[b] = tl[b];
  if (t < tl[b]) bt.y[(size_t)b * bt.ymax + 1 + bt.P[b] + t] = cb0[(size_t)b * Tmax + t];
}

// ids index embedding tables directly: flag |= bit when a[b*stride + i] (i < lens[b]*inner, or i < count when lens is null)
// lies outside [lo, hi).  The reference's F.embedding raises IndexError for those (models/vallex.py:497-507,605-632).
__global__ void check_range_kernel(const int32_t* __restrict__ a, int64_t stride, const int32_t* __restrict__ lens, int inner,
                                   int64_t count, int lo, int hi, int bit, int32_t* __restrict__ flag) {
  const int b = blockIdx.y;
  const int64_t n = lens ? (int64_t)lens[b] * inner : count;
  bool bad = false;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = a[(size_t)b * stride + i];
    bad |= (v < lo || v >= hi);
  }
  if (bad) atomicOr(flag, bit);
}

// best_of beams (models/vallex.py:525-527): after ONE prefill of the utterance, beam b > 0 gets a copy of beam 0's decode state:
// KV rows [0, n) of every layer and head, the token buffer, the first-step logits.  grid (layers * heads, beams - 1).
template <typename KVT>
__global__ void fork_kv_kernel(KVT* __restrict__ kc, KVT* __restrict__ vc, size_t kv_layer_elems, int ctx_cap, int n) {
  const int layer = blockIdx.x / kH, h = blockIdx.x % kH, b = blockIdx.y + 1;
  const size_t src = (size_t)layer * kv_layer_elems + (size_t)h * ctx_cap * kDH;
  const size_t dst = (size_t)layer * kv_layer_elems + ((size_t)b * kH + h) * ctx_cap * kDH;
  const size_t cnt = (size_t)n * kDH * sizeof(KVT) / 16;   // rows are 128 B (bf16) / 256 B (fp32): whole uint4s
  const uint4* ks = reinterpret_cast<const uint4*>(kc + src);
  const uint4* vs = reinterpret_cast<const uint4*>(vc + src);
  uint4* kd = reinterpret_cast<uint4*>(kc + dst);
  uint4* vd = reinterpret_cast<uint4*>(vc + dst);
  for (size_t i = threadIdx.x; i < cnt; i += blockDim.x) { kd[i] = ks[i]; vd[i] = vs[i]; }
}
__global__ void fork_state_kernel(Batch bt, int nb, int ylen, float* __restrict__ logits, float* __restrict__ xs, int32_t* S, int32_t* P, int32_t* row0) {
  const int b = blockIdx.x + 1;
  for (int i = threadIdx.x; i < ylen; i += blockDim.x) bt.y[(size_t)b * bt.ymax + i] = bt.y[i];
  for (int i = threadIdx.x; i < kARLogits; i += blockDim.x) logits[(size_t)b * kARLogits + i] = logits[i];
  for (int i = threadIdx.x; i < kD; i += blockDim.x) xs[(size_t)b * kD + i] = xs[i];
  if (threadIdx.x == 0) {
    S[b] = S[0]; P[b] = P[0]; row0[b] = row0[0];
    bt.T[b] = 0; bt.finished[b] = 0; bt.sum_logprobs[b] = 0.f;
  }
}
// models/vallex.py:583-594: avg = sum_logprobs / len ** length_penalty with len = BOS + prompt + generated non-EOS tokens;
// best = first argmax, worst = first argmin; the chosen beam becomes utterance 0 (tokens, length) for the NAR stages.
__global__ void select_beam_kernel(Batch bt, int nb, float length_penalty, int return_worst, int32_t* __restrict__ out_pick) {
  __shared__ int s_pick;
  if (threadIdx.x == 0) {
    int pick = 0;
    float best = 0.f;
    for (int b = 0; b < nb; ++b) {
      const float len = (float)(1 + bt.P[b] + bt.T[b]);
      const float avg = bt.sum_logprobs[b] / powf(len, length_penalty);
      if (b == 0 || (return_worst ? avg < best : avg > best)) { best = avg; pick = b; }
    }
    s_pick = pick;
    out_pick[0] = pick;
    out_pick[1] = bt.T[pick];
  }
  __syncthreads();
  const int pick = s_pick;
  if (pick != 0) {
    const int n = 1 + bt.P[pick] + bt.T[pick];
    for (int i = threadIdx.x; i < n; i += blockDim.x) bt.y[i] = bt.y[(size_t)pick * bt.ymax + i];
    __syncthreads();
    if (threadIdx.x == 0) { bt.T[0] = bt.T[pick]; bt.sum_logprobs[0] = bt.sum_logprobs[pick]; }
  }
}

struct Weight {
  std::vector<int64_t> shape;
  float* f32 = nullptr;  // device fp32 (always present until finalize may drop it for bf16 matrices)
  bf16* b16 = nullptr;   // device bf16 row-major copy (bf16 mode: decode GEMV operands, tied embedding tables)
  bf16* img = nullptr;   // device bf16 tile image (bf16 mode: tcgen05 GEMM operands, see tc_common.cuh)
  size_t numel = 0;
};

struct Buf {
  void* p = nullptr;
  size_t cap = 0;
};

struct LayerW {
  const void *qkv_w, *out_w, *ff1_w, *ff2_w;
  const void *qkv_i, *out_i, *ff1_i, *ff2_i;  // tile images (bf16 mode)
  const void* ff2_t = nullptr;                // bf16 [4096][1024] transposed FFN2 weight (AR, fused decode variant)
  const float *qkv_b, *out_b, *ff1_b, *ff2_b, *n1_g, *n1_b, *n2_g, *n2_b;
  const float *ada1, *ada2;  // [7][2048] (NAR only)
};

}  // namespace

struct vx_engine {
  int device = 0, max_batch = 0, max_ctx = 0, mode = 0;
  cudaStream_t own_stream = nullptr;
  std::map<std::string, Weight> w;
  bool finalized = false, has_vocos = false, has_valle = false, force_simt = false;
  LayerW ar[kL], nar[kL];
  const float *ar_norm_g = nullptr, *ar_norm_b = nullptr, *nar_norm_g = nullptr, *nar_norm_b = nullptr;
  const float* nar_norm_ada = nullptr;  // [7][2048]
  float* ada_all = nullptr;             // [25][7][2048]
  const void* ar_predict = nullptr;
  const void* nar_predict[7] = {};
  const void* nar_predict_img[7] = {};
  EmbTables nar_emb{};
  const float* pe = nullptr;
  int pe_rows = 0;
  float* dft_basis = nullptr;  // [1280][1296]
  // tensor-core Vocos path (bf16 mode): split-3 weight images [hi | lo | hi] of every contraction, tap-major depthwise weights
  struct { bool ready = false; void *embed = nullptr, *pw1[8] = {}, *pw2[8] = {}, *head = nullptr, *dft = nullptr; float* dwT[8] = {}; float* head_bias = nullptr; } vtc;
  Buf v_imgA, v_imgB;

  // batch state
  bool batch_set = false, prefilled = false, decoded = false, forked = false;
  size_t kv_slot_off = 0;                            // element offset of the slot a single-utterance prefill writes (vx_slot_prefill)
  std::vector<int32_t> hfin;                         // host mirror of finished[] after the last decode launch
  Buf d_pending, d_zero;                             // [max_batch] fresh-prefill flags (continuous batching); one zero int
  int B = 0, Smax = 0, Pmax = 0, ymax = 0;
  std::vector<int> hS, hP, hT, h_ar_row0;
  int ar_rows = 0, ar_maxlen = 0;
  Batch bt{};
  Buf d_flag;                                        // [4] id range-check flags
  Buf d_text, d_lang, d_prompt, d_S, d_P, d_ar_row0, d_nar_row0, d_T, d_ctx, d_y, d_budget, d_fin, d_slp, d_nact;
  Buf x, xn, qkv, att, ffh, kc, vc;                 // prefill / NAR activations, KV cache
  Buf a_img, qkv16, att_img, h_img, t_img;           // bf16 mode: tile-image / bf16 activations
  Buf q_img, k_img, vt_img;                          // per-head attention operand images (tcgen05 attention)
  bool decode_fused = true;                          // 3-barrier persistent decode variant (bf16); VX_DECODE_FUSED=0 disables
  int decode_fused_maxb = 2;
  int decode_cluster_grid = -1;                      // CTAs of the batch-1 cluster kernel on this device (0: not launchable; -1: not asked yet)
  void* ff2_t_blob = nullptr;                        // 12 x [4096][1024] bf16
  bool attn_t5 = false;
  bool attn_t5_fuse = false;                         // the QKV GEMM epilogue writes the per-head operand images (NAR layers)
  int64_t full_rows = 0;                             // packed rows of the current full-sequence pass (128-aligned starts)
  Buf xs, xsn, qkvs, atts, ffs, logits, part, tick;  // decode-step scratch
  Buf p_xb, p_acc;                                 // fused decode variant: mid-layer residual, fixed-point accumulators
  Buf p_pf, p_bar, p_prof;                                   // persistent decode: FFN2 partials, barrier counter + steps
  int num_sms = 148;
  bool decode_graph = false;                         // VX_DECODE_GRAPH=1: per-step CUDA-graph path instead of the persistent kernel
  Buf lastmap, out_tok, forced, yemb, yrow0, trow0, tgtmap, nlog, codes, tl_dev, cb0_dev;
  Buf v_codes, v_feat, v_col, v_x, v_h, v_ff, v_head, v_spec, v_frames, v_wave;
  int32_t* h_pinned = nullptr;  // small pinned scratch for polling
  int ctx_cap = 0;
  size_t kv_layer_elems = 0;

  bool ctx_clamped = false;                          // last vx_ar_decode: some budget was cut by max_ctx (vx_ar_ctx_clamped)
  std::string decode_kernel;                         // name of the kernel the last vx_ar_decode launched
  float ms_prefill = 0, ms_decode = 0, ms_nar = 0, ms_vocos = 0;
  int64_t decode_steps = 0, launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace {

int ensure(Buf& b, size_t bytes) {
  if (bytes <= b.cap) return VX_OK;
  if (b.p) cudaFree(b.p);
  b.p = nullptr;
  b.cap = 0;
  const size_t want = (bytes + 255) & ~size_t(255);
  cudaError_t err = cudaMalloc(&b.p, want);
  if (err != cudaSuccess) return fail(VX_ERR_CUDA, std::string("cudaMalloc ") + std::to_string(want) + ": " + cudaGetErrorString(err));
  // A grown workspace starts zeroed, like the first allocation of a process usually is: recycled device memory holds stale bit
  // patterns (NaN / Inf among them), and padded operand rows that are multiplied by exact zeros must be finite.  Growth is rare
  // (the buffers only ever grow), so the synchronous fill costs nothing in steady state.
  err = cudaMemset(b.p, 0, want);
  if (err == cudaSuccess) err = cudaDeviceSynchronize();
  if (err != cudaSuccess) return fail(VX_ERR_CUDA, std::string("cudaMemset of a new workspace: ") + cudaGetErrorString(err));
  b.cap = want;
  return VX_OK;
}
#define ENS(buf, bytes)                        \
  do {                                         \
    int _r = ensure(buf, (size_t)(bytes));     \
    if (_r != VX_OK) return _r;                \
  } while (0)

int copy_in(void* dst, const void* src, size_t bytes, int mem, cudaStream_t st) {
  if (bytes == 0) return VX_OK;
  CK(cudaMemcpyAsync(dst, src, bytes, mem == VX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice, st));
  return VX_OK;
}
int copy_out(void* dst, const void* src, size_t bytes, int mem, cudaStream_t st) {
  if (bytes == 0) return VX_OK;
  CK(cudaMemcpyAsync(dst, src, bytes, mem == VX_HOST ? cudaMemcpyDeviceToHost : cudaMemcpyDeviceToDevice, st));
  return VX_OK;
}

size_t kv_esize(const vx_engine* e) { return e->mode == VX_BF16 ? 2 : 4; }

const Weight* findw(vx_engine* e, const std::string& k) {
  auto it = e->w.find(k);
  return it == e->w.end() ? nullptr : &it->second;
}

bool is_gemm_matrix(const std::string& k) {
  auto ends = [&](const char* s) {
    const size_t n = strlen(s);
    return k.size() >= n && k.compare(k.size() - n, n, s) == 0;
  };
  if (k.rfind("vocos.", 0) == 0) return false;
  if (ends("self_attn.in_proj_weight") || ends("self_attn.out_proj.weight") || ends("linear1.weight") ||
      ends("linear2.weight"))
    return true;
  if (k == "ar_predict_layer.weight") return true;
  if (k.rfind("nar_predict_layers.", 0) == 0) return true;
  for (int j = 2; j < 8; ++j)
    if (k == "nar_audio_embeddings." + std::to_string(j) + ".word_embeddings.weight") return true;
  return false;
}

// matrix pointer in the engine's GEMM storage type
const void* mat(vx_engine* e, const std::string& k) {
  const Weight* w = findw(e, k);
  if (!w) return nullptr;
  return (e->mode == VX_BF16 && w->b16) ? (const void*)w->b16 : (const void*)w->f32;
}
const void* matimg(vx_engine* e, const std::string& k) {
  const Weight* w = findw(e, k);
  return w ? (const void*)w->img : nullptr;
}
const float* vec(vx_engine* e, const std::string& k) {
  const Weight* w = findw(e, k);
  return w ? w->f32 : nullptr;
}

struct Req {
  std::string key;
  std::vector<int64_t> shape;
};

std::vector<Req> required_keys() {
  std::vector<Req> r;
  auto add = [&](const std::string& k, std::vector<int64_t> s) { r.push_back({k, s}); };
  add("ar_text_embedding.word_embeddings.weight", {kTextVocab, kD});
  add("nar_text_embedding.word_embeddings.weight", {kTextVocab, kD});
  add("ar_audio_embedding.word_embeddings.weight", {kVocab + 2, kD});
  add("ar_text_position.alpha", {1});
  add("ar_audio_position.alpha", {1});
  add("nar_text_position.alpha", {1});
  add("nar_audio_position.alpha", {1});
  add("ar_decoder.norm.weight", {kD});
  add("ar_decoder.norm.bias", {kD});
  add("ar_predict_layer.weight", {kVocab + 1, kD});
  add("ar_language_embedding.word_embeddings.weight", {3, kD});
  add("nar_language_embedding.word_embeddings.weight", {3, kD});
  add("nar_audio_embeddings.0.word_embeddings.weight", {kVocab + 1, kD});
  for (int j = 1; j < kNQ; ++j) add("nar_audio_embeddings." + std::to_string(j) + ".word_embeddings.weight", {kVocab, kD});
  for (int j = 0; j < kNQ - 1; ++j) {
    add("nar_predict_layers." + std::to_string(j) + ".weight", {kVocab, kD});
    add("nar_stage_embeddings." + std::to_string(j) + ".word_embeddings.weight", {1, kD});
  }
  for (int which = 0; which < 2; ++which) {
    for (int i = 0; i < kL; ++i) {
      const std::string p = std::string(which ? "nar" : "ar") + "_decoder.layers." + std::to_string(i);
      add(p + ".self_attn.in_proj_weight", {3 * kD, kD});
      add(p + ".self_attn.in_proj_bias", {3 * kD});
      add(p + ".self_attn.out_proj.weight", {kD, kD});
      add(p + ".self_attn.out_proj.bias", {kD});
      add(p + ".linear1.weight", {kFF, kD});
      add(p + ".linear1.bias", {kFF});
      add(p + ".linear2.weight", {kD, kFF});
      add(p + ".linear2.bias", {kD});
      for (const char* n : {".norm1", ".norm2"}) {
        if (which) {
          add(p + n + ".project_layer.weight", {2 * kD, kD});
          add(p + n + ".project_layer.bias", {2 * kD});
          add(p + n + ".norm.weight", {kD});
          add(p + n + ".norm.bias", {kD});
        } else {
          add(p + n + ".weight", {kD});
          add(p + n + ".bias", {kD});
        }
      }
    }
  }
  add("nar_decoder.norm.project_layer.weight", {2 * kD, kD});
  add("nar_decoder.norm.project_layer.bias", {2 * kD});
  add("nar_decoder.norm.norm.weight", {kD});
  add("nar_decoder.norm.norm.bias", {kD});
  return r;
}

std::vector<Req> vocos_keys() {
  std::vector<Req> r;
  auto add = [&](const std::string& k, std::vector<int64_t> s) { r.push_back({"vocos." + k, s}); };
  add("feature_extractor.codebook_weights", {16 * 1024, 128});
  add("backbone.embed.weight", {384, 128, 7});
  add("backbone.embed.bias", {384});
  add("backbone.norm.scale.weight", {4, 384});
  add("backbone.norm.shift.weight", {4, 384});
  for (int i = 0; i < 8; ++i) {
    const std::string p = "backbone.convnext." + std::to_string(i);
    add(p + ".dwconv.weight", {384, 1, 7});
    add(p + ".dwconv.bias", {384});
    add(p + ".norm.scale.weight", {4, 384});
    add(p + ".norm.shift.weight", {4, 384});
    add(p + ".pwconv1.weight", {1152, 384});
    add(p + ".pwconv1.bias", {1152});
    add(p + ".pwconv2.weight", {384, 1152});
    add(p + ".pwconv2.bias", {384});
    add(p + ".gamma", {384});
  }
  add("backbone.final_layer_norm.weight", {384});
  add("backbone.final_layer_norm.bias", {384});
  add("head.out.weight", {1282, 384});
  add("head.out.bias", {1282});
  add("head.istft.window", {1280});
  return r;
}

cudaStream_t pick(vx_engine* e, void* stream) { return stream ? (cudaStream_t)stream : e->own_stream; }

// C[M,N] = (LN?)(A) W^T + bias, choosing the GEMV (M<=8) or the tiled kernel.
int linear(vx_engine* e, const float* A, int64_t lda, const void* W, int wt, const float* bias, float* C,
           int64_t ldc, int64_t M, int N, int K, int epi, cudaStream_t st) {
  if (M <= 8) launch_gemv(A, lda, W, wt, bias, C, ldc, (int)M, N, K, epi, nullptr, nullptr, st);
  else launch_gemm(A, lda, W, wt, bias, C, ldc, M, N, K, epi, nullptr, st);
  e->launches++;
  return VX_OK;
}

// One pre-LN encoder layer over R packed rows (AR prefill: attn_mode 1 + KV scatter; NAR: attn_mode 0 + AdaLN).
// fp32 mode (or VX_FORCE_SIMT=1): exact SIMT kernels.  bf16 mode: LN -> bf16 tile image -> tcgen05 GEMMs with fused
// epilogues, mma.sync flash attention.
int full_layer(vx_engine* e, const LayerW& L, int64_t R, const int32_t* row0, int attn_mode, int maxlen,
               const float* ada1, const float* ada2, int kv_layer, cudaStream_t st, int64_t R0 = 0, int b0 = 0, int nb = -1) {
  // [R0, R0 + R) / utterances [b0, b0 + nb): the row range this call covers (NAR chunks, see vx_nar); default = everything
  Batch bt = e->bt;
  if (nb >= 0) { bt.B = nb; bt.S += b0; bt.P += b0; bt.T += b0; row0 += b0; }
  const int wt = e->mode;
  const size_t kvb = kv_esize(e);
  float* x = (float*)e->x.p + (size_t)R0 * kD;
  void* kc = kv_layer >= 0 ? (char*)e->kc.p + ((size_t)kv_layer * e->kv_layer_elems + e->kv_slot_off) * kvb : nullptr;
  void* vc = kv_layer >= 0 ? (char*)e->vc.p + ((size_t)kv_layer * e->kv_layer_elems + e->kv_slot_off) * kvb : nullptr;
  if (e->mode == VX_BF16 && !e->force_simt) {
    // tile images are [row tile][k tile] blocks: rows [R0, ...) of a [*, K] image start R0 * K * 2 bytes in (R0 % 128 == 0)
    char* a_img = (char*)e->a_img.p + (size_t)R0 * kD * 2;
    char* att_img = (char*)e->att_img.p + (size_t)R0 * kD * 2;
    char* h_img = (char*)e->h_img.p + (size_t)R0 * kFF * 2;
    launch_layernorm_image(x, nullptr, R, L.n1_g, L.n1_b, ada1, ada1 ? ada1 + kD : nullptr, 1e-5f, a_img, st);
    // VX_ATTN_T5=3 (experimental): NAR layers (no KV scatter) write the per-head operand images straight from the QKV GEMM
    const bool heads_epi = e->attn_t5 && e->attn_t5_fuse && kv_layer < 0;
    void *q_i = e->q_img.p, *k_i = e->k_img.p, *vt_i = e->vt_img.p;
    if (e->attn_t5 && e->attn_t5_fuse) {   // one allocation [q | k | v^T], each full_rows x 1024 bf16
      k_i = (char*)q_i + (size_t)e->full_rows * kD * 2;
      vt_i = (char*)q_i + (size_t)2 * e->full_rows * kD * 2;
    }
    if (heads_epi) {
      // per-head images [full_rows, 64] (q, k) and 64-key V^T blocks: row R0 of every head sits R0 * 128 bytes into its image
      launch_gemm_tc(a_img, L.qkv_i, L.qkv_b, (char*)q_i + (size_t)R0 * 128, e->full_rows, R, 3 * kD, kD, TC_EPI_QKV_HEADS, st);
    } else {
      if (R0 != 0) return fail(VX_ERR_STATE, "full_layer: row chunks need the QKV-heads epilogue");
      launch_gemm_tc(a_img, L.qkv_i, L.qkv_b, e->qkv16.p, 3 * kD, R, 3 * kD, kD, TC_EPI_BF16_ROW, st);
    }
    if (kv_layer >= 0) { launch_kv_scatter_bf16(bt, e->qkv16.p, kc, vc, maxlen, e->ctx_cap, st); e->launches++; }
    if (e->attn_t5) {
      if (!heads_epi) { launch_pack_qkv_heads(e->qkv16.p, e->full_rows, q_i, k_i, vt_i, st); e->launches++; }
      launch_attn_t5(bt, q_i, k_i, vt_i, e->full_rows, row0, attn_mode, maxlen, e->att_img.p, st);   // row0 is absolute: images un-offset
    } else {
      if (R0 != 0) return fail(VX_ERR_STATE, "full_layer: row chunks need the tcgen05 attention");
      launch_attn_tc(bt, e->qkv16.p, row0, attn_mode, maxlen, e->att_img.p, st);
    }
    launch_gemm_tc(att_img, L.out_i, L.out_b, x, kD, R, kD, kD, TC_EPI_F32_RESIDUAL, st);
    launch_layernorm_image(x, nullptr, R, L.n2_g, L.n2_b, ada2, ada2 ? ada2 + kD : nullptr, 1e-5f, a_img, st);
    launch_gemm_tc(a_img, L.ff1_i, L.ff1_b, h_img, kFF, R, kFF, kD, TC_EPI_BF16_IMAGE_RELU, st);
    launch_gemm_tc(h_img, L.ff2_i, L.ff2_b, x, kD, R, kD, kFF, TC_EPI_F32_RESIDUAL, st);
    e->launches += 7;
    return VX_OK;
  }
  if (R0 != 0) return fail(VX_ERR_STATE, "full_layer: row chunks exist in bf16 mode only");
  float* xn = (float*)e->xn.p; float* qkv = (float*)e->qkv.p; float* att = (float*)e->att.p; float* ffh = (float*)e->ffh.p;
  launch_layernorm(x, nullptr, R, kD, L.n1_g, L.n1_b, ada1, ada1 ? ada1 + kD : nullptr, 1e-5f, xn, st);
  launch_gemm(xn, kD, L.qkv_w, wt, L.qkv_b, qkv, 3 * kD, R, 3 * kD, kD, EPI_NONE, nullptr, st);
  if (kv_layer >= 0) { launch_kv_scatter(bt, qkv, kc, vc, wt, maxlen, e->ctx_cap, st); e->launches++; }
  launch_attn_full(bt, qkv, row0, attn_mode, maxlen, att, st);
  launch_gemm(att, kD, L.out_w, wt, L.out_b, x, kD, R, kD, kD, EPI_RESIDUAL, nullptr, st);
  launch_layernorm(x, nullptr, R, kD, L.n2_g, L.n2_b, ada2, ada2 ? ada2 + kD : nullptr, 1e-5f, xn, st);
  launch_gemm(xn, kD, L.ff1_w, wt, L.ff1_b, ffh, kFF, R, kFF, kD, EPI_RELU, nullptr, st);
  launch_gemm(ffh, kFF, L.ff2_w, wt, L.ff2_b, x, kD, R, kD, kFF, EPI_RESIDUAL, nullptr, st);
  e->launches += 7;
  return VX_OK;
}

// activation workspaces for R packed rows in the engine's mode
int ensure_full_buffers(vx_engine* e, int64_t R) {
  const int64_t Rp = (R + 127) / 128 * 128;
  ENS(e->x, Rp * kD * 4);
  if (e->mode == VX_BF16 && !e->force_simt) {
    ENS(e->a_img, Rp * kD * 2); ENS(e->qkv16, Rp * 3 * kD * 2); ENS(e->att_img, Rp * kD * 2); ENS(e->h_img, Rp * kFF * 2);
    if (e->attn_t5 && e->attn_t5_fuse) { ENS(e->q_img, (size_t)3 * Rp * kD * 2); }
    else if (e->attn_t5) { ENS(e->q_img, Rp * kD * 2); ENS(e->k_img, Rp * kD * 2); ENS(e->vt_img, Rp * kD * 2); }
    e->full_rows = Rp;
    ENS(e->xn, (size_t)std::max<int64_t>(e->max_batch, 1) * kD * 4);
  } else {
    ENS(e->xn, Rp * kD * 4); ENS(e->qkv, Rp * 3 * kD * 4); ENS(e->att, Rp * kD * 4); ENS(e->ffh, Rp * kFF * 4);
  }
  return VX_OK;
}

}  // namespace

extern "C" {

const char* vx_version(void) { return "vallex_b200 0.1 sm_100a"; }
const char* vx_last_error(void) { return g_err.c_str(); }

int vx_create(int device, int max_batch, int max_ctx, int dtype_mode, vx_engine** out) {
  if (!out || max_batch < 1 || max_batch > 1000 || max_ctx < 8 || (dtype_mode != VX_F32 && dtype_mode != VX_BF16))
    return fail(VX_ERR_ARG, "vx_create: bad arguments (1 <= max_batch <= 1000, max_ctx >= 8)");   // h_pinned holds max_batch + 1 ints
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(VX_ERR_ARG, "vx_create: no such CUDA device");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) return fail(VX_ERR_CUDA, "vx_create: this library only contains sm_100a code (B200 required)");
  vx_engine* e = new vx_engine();
  e->device = device;
  e->max_batch = max_batch;
  e->max_ctx = max_ctx;
  e->mode = dtype_mode;
  CK(cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&e->ev0));
  CK(cudaEventCreate(&e->ev1));
  CK(cudaMallocHost(&e->h_pinned, 4096));
  const char* fs = getenv("VX_FORCE_SIMT");
  e->force_simt = fs && fs[0] == '1';
  const char* fu = getenv("VX_DECODE_FUSED");   // unset: fused kernel for B <= 2 (validated sizes); 1: B <= 8; 0: never
  e->decode_fused = !(fu && fu[0] == '0');
  e->decode_fused_maxb = (fu && fu[0] == '1') ? 8 : 2;
  const char* t5 = getenv("VX_ATTN_T5");   // tcgen05/TMEM attention for prefill + NAR (default); VX_ATTN_T5=0: the mma.sync kernel
  e->attn_t5 = !(t5 && t5[0] == '0');
  e->attn_t5_fuse = e->attn_t5;             // the QKV GEMM epilogue writes the per-head operand images of the NAR layers
  const char* dg = getenv("VX_DECODE_GRAPH");
  e->decode_graph = dg && dg[0] == '1';
  e->num_sms = prop.multiProcessorCount;
  *out = e;
  return VX_OK;
}

void vx_destroy(vx_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  for (auto& kv : e->w) {
    if (kv.second.f32) cudaFree(kv.second.f32);
    if (kv.second.b16) cudaFree(kv.second.b16);
    if (kv.second.img) cudaFree(kv.second.img);
  }
  Buf* bufs[] = {&e->d_pending, &e->d_zero, &e->d_flag, &e->d_text, &e->d_lang, &e->d_prompt, &e->d_S, &e->d_P, &e->d_ar_row0, &e->d_nar_row0, &e->d_T,
                 &e->d_ctx, &e->d_y, &e->d_budget, &e->d_fin, &e->d_slp, &e->d_nact, &e->x, &e->xn, &e->qkv, &e->att,
                 &e->ffh, &e->kc, &e->vc, &e->a_img, &e->qkv16, &e->att_img, &e->h_img, &e->t_img, &e->q_img, &e->k_img, &e->vt_img, &e->xs, &e->xsn, &e->qkvs, &e->atts, &e->ffs, &e->logits, &e->part, &e->tick, &e->p_xb, &e->p_acc, &e->p_pf, &e->p_bar, &e->p_prof,
                 &e->lastmap, &e->out_tok, &e->forced, &e->yemb, &e->yrow0, &e->trow0, &e->tgtmap, &e->nlog, &e->codes,
                 &e->tl_dev, &e->cb0_dev, &e->v_codes, &e->v_feat, &e->v_col, &e->v_x, &e->v_h, &e->v_ff, &e->v_head,
                 &e->v_spec, &e->v_frames, &e->v_wave, &e->v_imgA, &e->v_imgB};
  for (Buf* b : bufs)
    if (b->p) cudaFree(b->p);
  if (e->ada_all) cudaFree(e->ada_all);
  if (e->ff2_t_blob) cudaFree(e->ff2_t_blob);
  if (e->dft_basis) cudaFree(e->dft_basis);
  for (void* q : {e->vtc.embed, e->vtc.head, e->vtc.dft, (void*)e->vtc.head_bias}) if (q) cudaFree(q);
  for (int i = 0; i < 8; ++i) { if (e->vtc.pw1[i]) cudaFree(e->vtc.pw1[i]); if (e->vtc.pw2[i]) cudaFree(e->vtc.pw2[i]); if (e->vtc.dwT[i]) cudaFree(e->vtc.dwT[i]); }
  if (e->h_pinned) cudaFreeHost(e->h_pinned);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->own_stream) cudaStreamDestroy(e->own_stream);
  delete e;
}

int vx_load_weight(vx_engine* e, const char* key, const float* data, const int64_t* shape, int ndim) {
  return vx_load_weight_mem(e, key, data, shape, ndim, VX_HOST);
}

int vx_load_weight_mem(vx_engine* e, const char* key, const float* data, const int64_t* shape, int ndim, int mem) {
  if (!e || !key || !data || !shape || ndim < 1 || ndim > 4 || (mem != VX_HOST && mem != VX_DEVICE))
    return fail(VX_ERR_ARG, "vx_load_weight: bad arguments");
  CK(cudaSetDevice(e->device));
  Weight wt;
  wt.numel = 1;
  for (int i = 0; i < ndim; ++i) {
    wt.shape.push_back(shape[i]);
    wt.numel *= (size_t)shape[i];
  }
  auto it = e->w.find(key);
  if (it != e->w.end()) {
    if (it->second.f32) cudaFree(it->second.f32);
    if (it->second.b16) cudaFree(it->second.b16);
    if (it->second.img) cudaFree(it->second.img);
    e->w.erase(it);
  }
  CK(cudaMalloc(&wt.f32, std::max<size_t>(wt.numel, 4) * sizeof(float)));
  CK(cudaMemcpy(wt.f32, data, wt.numel * sizeof(float), mem == VX_HOST ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice));
  e->w[key] = wt;
  e->finalized = false;
  return VX_OK;
}

int vx_finalize_weights(vx_engine* e) {
  if (!e) return fail(VX_ERR_ARG, "vx_finalize_weights: null engine");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->own_stream;
  e->has_vocos = findw(e, "vocos.head.out.weight") != nullptr;
  e->has_valle = false;
  for (auto& kv : e->w)
    if (kv.first.rfind("vocos.", 0) != 0 && kv.first != "pe_table") e->has_valle = true;
  if (!e->has_valle && !e->has_vocos) return fail(VX_ERR_STATE, "vx_finalize_weights: no weights loaded");
  if (e->has_valle) {
    for (const Req& r : required_keys()) {
      const Weight* w = findw(e, r.key);
      if (!w) return fail(VX_ERR_STATE, "missing checkpoint key: " + r.key);
      if (w->shape != r.shape) return fail(VX_ERR_ARG, "shape mismatch for checkpoint key: " + r.key);
    }
    const Weight* pe = findw(e, "pe_table");
    if (!pe || pe->shape.size() != 2 || pe->shape[1] != kD) return fail(VX_ERR_STATE, "missing pe_table [rows,1024]");
    e->pe = pe->f32;
    e->pe_rows = (int)pe->shape[0];
    if (e->pe_rows < e->max_ctx) return fail(VX_ERR_CAPACITY, "pe_table has fewer rows than max_ctx");
  }
  if (e->has_vocos)
    for (const Req& r : vocos_keys()) {
      const Weight* w = findw(e, r.key);
      if (!w) return fail(VX_ERR_STATE, "missing checkpoint key: " + r.key);
      if (w->shape != r.shape) return fail(VX_ERR_ARG, "shape mismatch for checkpoint key: " + r.key);
    }

  if (e->has_valle) {
  // AdaLN (weight, bias) = project_layer(stage_embedding) for 25 norms x 7 stages, in fp32
  // (modules/transformer.py:96-101).  Layout ada_all[norm][stage][2048].
  if (!e->ada_all) CK(cudaMalloc(&e->ada_all, (size_t)25 * 7 * 2 * kD * sizeof(float)));
  {
    float* stage_rows = nullptr;  // [7][1024] gathered stage embeddings
    CK(cudaMalloc(&stage_rows, 7 * kD * sizeof(float)));
    for (int i = 0; i < 7; ++i)
      CK(cudaMemcpyAsync(stage_rows + i * kD, vec(e, "nar_stage_embeddings." + std::to_string(i) + ".word_embeddings.weight"),
                         kD * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int n = 0; n < 25; ++n) {
      std::string p = (n < 24) ? "nar_decoder.layers." + std::to_string(n / 2) + (n % 2 ? ".norm2" : ".norm1") : "nar_decoder.norm";
      launch_gemv(stage_rows, kD, vec(e, p + ".project_layer.weight"), 0, vec(e, p + ".project_layer.bias"),
                  e->ada_all + (size_t)n * 7 * 2 * kD, 2 * kD, 7, 2 * kD, kD, EPI_NONE, nullptr, nullptr, st);
    }
    CK(cudaStreamSynchronize(st));
    CK(cudaFree(stage_rows));
  }
  // bf16 mode: every GEMM matrix gets a tile image (tcgen05 operand); the AR matrices and the tied embedding
  // tables also keep a row-major bf16 copy (decode GEMV / gathers).  The fp32 originals are released.
  if (e->mode == VX_BF16) {
    for (auto& kv : e->w) {
      Weight& w = kv.second;
      if (!is_gemm_matrix(kv.first) || w.img || !w.f32) continue;
      const int64_t rows = w.shape[0];
      const int K = (int)w.shape[1];
      const int64_t rows_pad = (rows + 127) / 128 * 128;
      CK(cudaMalloc(&w.img, (size_t)rows_pad * K * sizeof(bf16)));
      launch_pack_tile_image(w.f32, rows, K, w.img, st);
      const bool need_row_major = e->force_simt || (kv.first.rfind("nar_decoder.", 0) != 0 && kv.first.rfind("nar_predict_layers.", 0) != 0);
      if (need_row_major) {
        CK(cudaMalloc(&w.b16, w.numel * sizeof(bf16)));
        f32_to_bf16_kernel<<<296, 256, 0, st>>>(w.f32, w.b16, w.numel);
      }
    }
    CK(cudaStreamSynchronize(st));
    for (auto& kv : e->w) {
      Weight& w = kv.second;
      if (w.img && w.f32) {
        cudaFree(w.f32);
        w.f32 = nullptr;
      }
    }
  }
  const int wt = e->mode;
  (void)wt;
  for (int which = 0; which < 2; ++which)
    for (int i = 0; i < kL; ++i) {
      const std::string p = std::string(which ? "nar" : "ar") + "_decoder.layers." + std::to_string(i);
      LayerW& L = which ? e->nar[i] : e->ar[i];
      L.qkv_w = mat(e, p + ".self_attn.in_proj_weight");
      L.qkv_b = vec(e, p + ".self_attn.in_proj_bias");
      L.out_w = mat(e, p + ".self_attn.out_proj.weight");
      L.out_b = vec(e, p + ".self_attn.out_proj.bias");
      L.ff1_w = mat(e, p + ".linear1.weight");
      L.ff1_b = vec(e, p + ".linear1.bias");
      L.ff2_w = mat(e, p + ".linear2.weight");
      L.ff2_b = vec(e, p + ".linear2.bias");
      L.qkv_i = matimg(e, p + ".self_attn.in_proj_weight");
      L.out_i = matimg(e, p + ".self_attn.out_proj.weight");
      L.ff1_i = matimg(e, p + ".linear1.weight");
      L.ff2_i = matimg(e, p + ".linear2.weight");
      if (which) {
        L.n1_g = vec(e, p + ".norm1.norm.weight");
        L.n1_b = vec(e, p + ".norm1.norm.bias");
        L.n2_g = vec(e, p + ".norm2.norm.weight");
        L.n2_b = vec(e, p + ".norm2.norm.bias");
        L.ada1 = e->ada_all + (size_t)(2 * i) * 7 * 2 * kD;
        L.ada2 = e->ada_all + (size_t)(2 * i + 1) * 7 * 2 * kD;
      } else {
        L.n1_g = vec(e, p + ".norm1.weight");
        L.n1_b = vec(e, p + ".norm1.bias");
        L.n2_g = vec(e, p + ".norm2.weight");
        L.n2_b = vec(e, p + ".norm2.bias");
        L.ada1 = L.ada2 = nullptr;
      }
    }
  e->ar_norm_g = vec(e, "ar_decoder.norm.weight");
  e->ar_norm_b = vec(e, "ar_decoder.norm.bias");
  e->nar_norm_g = vec(e, "nar_decoder.norm.norm.weight");
  e->nar_norm_b = vec(e, "nar_decoder.norm.norm.bias");
  e->nar_norm_ada = e->ada_all + (size_t)24 * 7 * 2 * kD;
  e->ar_predict = mat(e, "ar_predict_layer.weight");
  if (e->decode_fused && e->mode == VX_BF16) {
    // fused decode variant: hidden unit j's FFN2 column as a contiguous row (W2^T [4096][1024] per AR layer)
    if (!e->ff2_t_blob) CK(cudaMalloc(&e->ff2_t_blob, (size_t)kL * kFF * kD * sizeof(bf16)));
    for (int i = 0; i < kL; ++i) {
      bf16* dst = (bf16*)e->ff2_t_blob + (size_t)i * kFF * kD;
      transpose_bf16_kernel<<<dim3(kFF / 32, kD / 32), dim3(32, 8), 0, st>>>((const bf16*)e->ar[i].ff2_w, dst, kD, kFF);
      e->ar[i].ff2_t = dst;
    }
    CK(cudaStreamSynchronize(st));
  }
  for (int j = 0; j < 7; ++j) {
    e->nar_predict[j] = mat(e, "nar_predict_layers." + std::to_string(j) + ".weight");
    e->nar_predict_img[j] = matimg(e, "nar_predict_layers." + std::to_string(j) + ".weight");
  }
  for (int j = 0; j < 8; ++j) {
    const std::string k = "nar_audio_embeddings." + std::to_string(j) + ".word_embeddings.weight";
    const Weight* w = findw(e, k);
    e->nar_emb.bf[j] = (e->mode == VX_BF16 && w->b16) ? 1 : 0;
    e->nar_emb.p[j] = e->nar_emb.bf[j] ? (const void*)w->b16 : (const void*)w->f32;
  }
  }
  if (e->has_vocos && !e->dft_basis) {
    // inverse real DFT as a contraction: frame[n] = sum_k spec_re[k]*Bc[n][k] + spec_im[k]*Bs[n][k]
    // (torch.fft.irfft(n=1280, norm="backward"): imaginary parts of DC and Nyquist are ignored).
    const int N = 1280, KB = 641, LD = 1296;
    std::vector<float> h((size_t)N * LD, 0.f);
    for (int n = 0; n < N; ++n)
      for (int k = 0; k < KB; ++k) {
        const double c = (k == 0 || k == KB - 1) ? 1.0 : 2.0;
        const double ang = 2.0 * M_PI * (double)((int64_t)k * n % N) / (double)N;
        h[(size_t)n * LD + k] = (float)(c * std::cos(ang) / N);
        h[(size_t)n * LD + KB + k] = (k == 0 || k == KB - 1) ? 0.f : (float)(-c * std::sin(ang) / N);
      }
    CK(cudaMalloc(&e->dft_basis, h.size() * sizeof(float)));
    CK(cudaMemcpy(e->dft_basis, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    // Conv1d(128->384,k=7) weight [co][ci][j] -> im2col layout [co][j*128+ci]
    const Weight* cw = findw(e, "vocos.backbone.embed.weight");
    std::vector<float> src(cw->numel), dst(cw->numel);
    CK(cudaMemcpy(src.data(), cw->f32, cw->numel * sizeof(float), cudaMemcpyDeviceToHost));
    for (int co = 0; co < 384; ++co)
      for (int ci = 0; ci < 128; ++ci)
        for (int j = 0; j < 7; ++j) dst[(size_t)co * 896 + j * 128 + ci] = src[((size_t)co * 128 + ci) * 7 + j];
    Weight packed;
    packed.shape = {384, 896};
    packed.numel = cw->numel;
    CK(cudaMalloc(&packed.f32, packed.numel * sizeof(float)));
    CK(cudaMemcpy(packed.f32, dst.data(), packed.numel * sizeof(float), cudaMemcpyHostToDevice));
    e->w["vocos.__embed_im2col"] = packed;
  }
  if (e->has_vocos && e->mode == VX_BF16 && !e->force_simt && !e->vtc.ready) {
    // tensor-core Vocos path: split-3 weight images (bf16 x 3; vocos.cu) of the im2col'ed embed conv, the 16 point-wise
    // linears, the head (1282 rows padded to 1408) and the inverse-DFT basis (K 1282 -> 1344)
    auto img3 = [&](const float* w, int64_t rows, int K_src, int64_t ld, int K, void** out) -> int {
      const int64_t rows_pad = (rows + 127) / 128 * 128;
      CK(cudaMalloc(out, (size_t)rows_pad * 3 * K * sizeof(bf16)));
      launch_pack_split3_weight(w, rows, K_src, ld, K, *out, st);
      return VX_OK;
    };
    int r;
    if ((r = img3(vec(e, "vocos.__embed_im2col"), 384, 896, 896, 896, &e->vtc.embed))) return r;
    for (int i = 0; i < 8; ++i) {
      const std::string p = "vocos.backbone.convnext." + std::to_string(i);
      if ((r = img3(vec(e, p + ".pwconv1.weight"), 1152, 384, 384, 384, &e->vtc.pw1[i]))) return r;
      if ((r = img3(vec(e, p + ".pwconv2.weight"), 384, 1152, 1152, 1152, &e->vtc.pw2[i]))) return r;
      std::vector<float> w(384 * 7), wt(7 * 384);
      CK(cudaMemcpy(w.data(), vec(e, p + ".dwconv.weight"), w.size() * 4, cudaMemcpyDeviceToHost));
      for (int c = 0; c < 384; ++c)
        for (int j = 0; j < 7; ++j) wt[j * 384 + c] = w[c * 7 + j];
      CK(cudaMalloc(&e->vtc.dwT[i], wt.size() * 4));
      CK(cudaMemcpy(e->vtc.dwT[i], wt.data(), wt.size() * 4, cudaMemcpyHostToDevice));
    }
    if ((r = img3(vec(e, "vocos.head.out.weight"), 1282, 384, 384, 384, &e->vtc.head))) return r;
    CK(cudaMalloc(&e->vtc.head_bias, 1408 * 4));
    CK(cudaMemset(e->vtc.head_bias, 0, 1408 * 4));
    CK(cudaMemcpy(e->vtc.head_bias, vec(e, "vocos.head.out.bias"), 1282 * 4, cudaMemcpyDeviceToDevice));
    if ((r = img3(e->dft_basis, 1280, 1282, 1296, 1344, &e->vtc.dft))) return r;
    CK(cudaStreamSynchronize(st));
    CK(cudaGetLastError());
    e->vtc.ready = true;
  }
  e->finalized = true;
  return VX_OK;
}

int vx_set_batch(vx_engine* e, int B, const int32_t* text_ids, const int32_t* text_lens, const int32_t* lang_ids,
                 int Smax, const int32_t* prompt_codes, const int32_t* prompt_lens, int Pmax, int mem, void* stream) {
  if (!e || !e->finalized || !e->has_valle) return fail(VX_ERR_STATE, "vx_set_batch: VALL-E X weights not finalized");
  if (B < 1 || B > e->max_batch || !text_ids || !text_lens || !lang_ids || !prompt_lens || Smax < 1 || Pmax < 0 ||
      (Pmax > 0 && !prompt_codes))
    return fail(VX_ERR_ARG, "vx_set_batch: bad arguments");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  e->hS.assign(B, 0);
  e->hP.assign(B, 0);
  if (mem == VX_HOST) {
    for (int b = 0; b < B; ++b) { e->hS[b] = text_lens[b]; e->hP[b] = prompt_lens[b]; }
  } else {
    CK(cudaMemcpyAsync(e->hS.data(), text_lens, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(e->hP.data(), prompt_lens, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  }
  e->h_ar_row0.assign(B, 0);
  int rows = 0, maxlen = 0;
  for (int b = 0; b < B; ++b) {
    if (e->hS[b] < 1 || e->hS[b] > Smax) return fail(VX_ERR_ARG, "vx_set_batch: text_lens out of range (x_lens > 0 required, models/vallex.py:493)");
    if (e->hP[b] < 0 || e->hP[b] > Pmax) return fail(VX_ERR_ARG, "vx_set_batch: prompt_lens out of range");
    const int len = e->hS[b] + e->hP[b] + 1;
    if (len >= e->max_ctx) return fail(VX_ERR_CAPACITY, "vx_set_batch: S+P+1 exceeds max_ctx");
    e->h_ar_row0[b] = rows;
    rows += (e->attn_t5 && e->mode == VX_BF16 && !e->force_simt) ? (len + 127) / 128 * 128 : len;  // tcgen05 attention: key tiles = image blocks
    maxlen = std::max(maxlen, len);
  }
  e->B = B; e->Smax = Smax; e->Pmax = Pmax; e->ar_rows = rows; e->ar_maxlen = maxlen;
  e->ymax = e->max_ctx + 1;
  e->ctx_cap = e->max_ctx;
  ENS(e->d_text, (size_t)B * Smax * 4);
  ENS(e->d_lang, (size_t)B * Smax * 4);
  ENS(e->d_prompt, (size_t)B * std::max(Pmax, 1) * kNQ * 4);
  const int MB = e->max_batch;   // per-utterance state is sized for max_batch: vx_ar_fork grows the batch in place
  ENS(e->d_S, MB * 4); ENS(e->d_P, MB * 4); ENS(e->d_ar_row0, MB * 4); ENS(e->d_nar_row0, MB * 4);
  ENS(e->d_T, MB * 4); ENS(e->d_ctx, MB * 4); ENS(e->d_budget, MB * 4); ENS(e->d_fin, MB * 4);
  ENS(e->d_slp, MB * 4); ENS(e->d_nact, 4);
  ENS(e->d_y, (size_t)MB * e->ymax * 4);
  int r;
  if ((r = copy_in(e->d_text.p, text_ids, (size_t)B * Smax * 4, mem, st))) return r;
  if ((r = copy_in(e->d_lang.p, lang_ids, (size_t)B * Smax * 4, mem, st))) return r;
  if (Pmax > 0 && (r = copy_in(e->d_prompt.p, prompt_codes, (size_t)B * Pmax * kNQ * 4, mem, st))) return r;
  CK(cudaMemcpyAsync(e->d_S.p, e->hS.data(), B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->d_P.p, e->hP.data(), B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->d_ar_row0.p, e->h_ar_row0.data(), B * 4, cudaMemcpyHostToDevice, st));
  ENS(e->d_flag, 16);
  CK(cudaMemsetAsync(e->d_flag.p, 0, 16, st));
  check_range_kernel<<<dim3(4, B), 256, 0, st>>>((const int32_t*)e->d_text.p, Smax, (const int32_t*)e->d_S.p, 1, 0, 0, kTextVocab, 1, (int32_t*)e->d_flag.p);
  check_range_kernel<<<dim3(4, B), 256, 0, st>>>((const int32_t*)e->d_lang.p, Smax, (const int32_t*)e->d_S.p, 1, 0, 0, 3, 2, (int32_t*)e->d_flag.p);
  if (Pmax > 0)
    check_range_kernel<<<dim3(8, B), 256, 0, st>>>((const int32_t*)e->d_prompt.p, (int64_t)Pmax * kNQ, (const int32_t*)e->d_P.p, kNQ, 0, 0, kVocab, 4, (int32_t*)e->d_flag.p);
  CK(cudaMemcpyAsync(e->h_pinned, e->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));  // hS/hP/h_ar_row0 are reused by later calls
  if (e->h_pinned[0] != 0) {
    e->batch_set = false;
    const int f = e->h_pinned[0];
    return fail(VX_ERR_ARG, std::string("vx_set_batch: index out of range in") + ((f & 1) ? " text_ids (0..2047)" : "") +
                            ((f & 2) ? " lang_ids (0..2)" : "") + ((f & 4) ? " prompt_codes (0..1023)" : ""));
  }
  Batch& bt = e->bt;
  bt.B = B; bt.Smax = Smax; bt.Pmax = std::max(Pmax, 1);
  bt.text_ids = (const int32_t*)e->d_text.p; bt.lang_ids = (const int32_t*)e->d_lang.p;
  bt.prompt = (const int32_t*)e->d_prompt.p;
  bt.S = (const int32_t*)e->d_S.p; bt.P = (const int32_t*)e->d_P.p;
  bt.ar_row0 = (const int32_t*)e->d_ar_row0.p; bt.nar_row0 = (const int32_t*)e->d_nar_row0.p;
  bt.T = (int32_t*)e->d_T.p; bt.ctx = (int32_t*)e->d_ctx.p; bt.y = (int32_t*)e->d_y.p; bt.ymax = e->ymax;
  bt.budget = (int32_t*)e->d_budget.p; bt.finished = (int32_t*)e->d_fin.p;
  bt.sum_logprobs = (float*)e->d_slp.p; bt.n_active = (int32_t*)e->d_nact.p;
  e->Pmax = Pmax;
  e->batch_set = true; e->prefilled = false; e->decoded = false; e->forked = false;
  return VX_OK;
}

// The AR prefill over the utterances e->bt describes (B of them, R packed rows): embeddings, 12 masked layers with KV scatter,
// final LayerNorm of each utterance's last row, first-step logits into `logits_out` [B][1025].
static int prefill_core(vx_engine* e, int B, int64_t R, int maxlen, const int32_t* h_last, float* logits_out, cudaStream_t st) {
  const int wt = e->mode;
  { int _r = ensure_full_buffers(e, R); if (_r) return _r; }
  const Batch& bt = e->bt;
  float* x = (float*)e->x.p;
  if (e->attn_t5 && e->mode == VX_BF16 && !e->force_simt) CK(cudaMemsetAsync(x, 0, (size_t)e->full_rows * kD * 4, st));  // pad rows stay finite
  launch_init_y(bt, st);
  launch_embed_text(bt, vec(e, "ar_text_embedding.word_embeddings.weight"), vec(e, "ar_language_embedding.word_embeddings.weight"),
                    vec(e, "ar_text_position.alpha"), e->pe, bt.ar_row0, x, st);
  launch_embed_audio_prefill(bt, vec(e, "ar_audio_embedding.word_embeddings.weight"), vec(e, "ar_audio_position.alpha"), e->pe, x, st);
  e->launches += 3;
  for (int l = 0; l < kL; ++l) full_layer(e, e->ar[l], R, bt.ar_row0, 1, maxlen, nullptr, nullptr, l, st);
  memcpy(e->h_pinned, h_last, B * 4);
  CK(cudaMemcpyAsync(e->lastmap.p, e->h_pinned, B * 4, cudaMemcpyHostToDevice, st));
  launch_layernorm(x, (const int32_t*)e->lastmap.p, B, kD, e->ar_norm_g, e->ar_norm_b, nullptr, nullptr, 1e-5f, (float*)e->xsn.p, st);
  e->launches++;
  linear(e, (float*)e->xsn.p, kD, e->ar_predict, wt, nullptr, logits_out, kARLogits, B, kARLogits, kD, EPI_NONE, st);
  return VX_OK;
}

int vx_ar_prefill(vx_engine* e, void* stream) {
  if (!e || !e->batch_set) return fail(VX_ERR_STATE, "vx_ar_prefill: no batch set");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int B = e->B;
  const size_t kvb = kv_esize(e);
  e->kv_layer_elems = (size_t)e->max_batch * kH * e->ctx_cap * kDH;
  ENS(e->kc, e->kv_layer_elems * kL * kvb); ENS(e->vc, e->kv_layer_elems * kL * kvb);
  ENS(e->xs, (size_t)e->max_batch * kD * 4); ENS(e->xsn, (size_t)e->max_batch * kD * 4); ENS(e->logits, (size_t)e->max_batch * kARLogits * 4);
  ENS(e->lastmap, e->max_batch * 4);
  ENS(e->d_pending, e->max_batch * 4); ENS(e->d_zero, 16);
  CK(cudaMemsetAsync(e->d_pending.p, 0, e->max_batch * 4, st));
  CK(cudaMemsetAsync(e->d_zero.p, 0, 16, st));
  CK(cudaEventRecord(e->ev0, st));
  std::vector<int32_t> last(B);
  for (int b = 0; b < B; ++b) last[b] = e->h_ar_row0[b] + e->hS[b] + e->hP[b];
  e->kv_slot_off = 0;
  { int _r = prefill_core(e, B, e->ar_rows, e->ar_maxlen, last.data(), (float*)e->logits.p, st); if (_r) return _r; }
  CK(cudaMemsetAsync(e->d_T.p, 0, B * 4, st));
  CK(cudaMemsetAsync(e->d_fin.p, 0, B * 4, st));
  CK(cudaMemsetAsync(e->d_slp.p, 0, B * 4, st));
  CK(cudaEventRecord(e->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&e->ms_prefill, e->ev0, e->ev1));
  CK(cudaGetLastError());
  e->hfin.assign(B, 0);
  e->prefilled = true; e->decoded = false;
  return VX_OK;
}

// Continuous batching: the utterance in `slot` has finished; a new one takes its place while the other slots keep
// their decode state.  The new utterance is prefilled ALONE (batch-1 view of the slot's rows of every per-utterance array, KV
// scatter into the slot's cache region, first-step logits into the slot's row) and marked pending, so the next
// vx_ar_decode_ex(resume = 1) samples its first token while the continuing slots carry on from the token they sampled last.
int vx_slot_prefill(vx_engine* e, int slot, const int32_t* text_ids, int S, const int32_t* lang_ids, const int32_t* prompt_codes, int P,
                    int mem, void* stream) {
  if (!e || !e->batch_set || !e->prefilled) return fail(VX_ERR_STATE, "vx_slot_prefill: call after vx_set_batch + vx_ar_prefill");
  if (e->forked) return fail(VX_ERR_STATE, "vx_slot_prefill: the batch holds the beams of one utterance");
  if (slot < 0 || slot >= e->B || !text_ids || !lang_ids || S < 1 || P < 0 || (P > 0 && !prompt_codes))
    return fail(VX_ERR_ARG, "vx_slot_prefill: bad arguments");
  if (S > e->Smax || P > e->Pmax) return fail(VX_ERR_CAPACITY, "vx_slot_prefill: S / P exceed the capacity vx_set_batch reserved (Smax, Pmax)");
  if (S + P + 1 >= e->max_ctx) return fail(VX_ERR_CAPACITY, "vx_slot_prefill: S+P+1 exceeds max_ctx");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int Pm = std::max(e->Pmax, 1);
  int32_t* d_text = (int32_t*)e->d_text.p + (size_t)slot * e->Smax;
  int32_t* d_lang = (int32_t*)e->d_lang.p + (size_t)slot * e->Smax;
  int32_t* d_prm = (int32_t*)e->d_prompt.p + (size_t)slot * Pm * kNQ;
  int r;
  if ((r = copy_in(d_text, text_ids, (size_t)S * 4, mem, st))) return r;
  if ((r = copy_in(d_lang, lang_ids, (size_t)S * 4, mem, st))) return r;
  if (P > 0 && (r = copy_in(d_prm, prompt_codes, (size_t)P * kNQ * 4, mem, st))) return r;
  e->h_pinned[0] = S; e->h_pinned[1] = P; e->h_pinned[2] = 1; e->h_pinned[3] = 0;
  CK(cudaMemcpyAsync((int32_t*)e->d_S.p + slot, e->h_pinned, 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync((int32_t*)e->d_P.p + slot, e->h_pinned + 1, 4, cudaMemcpyHostToDevice, st));
  ENS(e->d_flag, 16);
  CK(cudaMemsetAsync(e->d_flag.p, 0, 16, st));
  check_range_kernel<<<dim3(4, 1), 256, 0, st>>>(d_text, 0, nullptr, 1, S, 0, kTextVocab, 1, (int32_t*)e->d_flag.p);
  check_range_kernel<<<dim3(4, 1), 256, 0, st>>>(d_lang, 0, nullptr, 1, S, 0, 3, 2, (int32_t*)e->d_flag.p);
  if (P > 0) check_range_kernel<<<dim3(8, 1), 256, 0, st>>>(d_prm, 0, nullptr, 1, (int64_t)P * kNQ, 0, kVocab, 4, (int32_t*)e->d_flag.p);
  CK(cudaMemcpyAsync(e->h_pinned + 8, e->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (e->h_pinned[8] != 0) return fail(VX_ERR_ARG, "vx_slot_prefill: index out of range in text_ids / lang_ids / prompt_codes");
  // batch-1 view of the slot
  const Batch full = e->bt;
  Batch v = full;
  v.B = 1;
  v.text_ids = d_text; v.lang_ids = d_lang; v.prompt = d_prm;
  v.S = full.S + slot; v.P = full.P + slot; v.ar_row0 = (const int32_t*)e->d_zero.p; v.nar_row0 = (const int32_t*)e->d_zero.p;
  v.T = full.T + slot; v.ctx = full.ctx + slot; v.y = full.y + (size_t)slot * full.ymax;
  v.budget = full.budget + slot; v.finished = full.finished + slot; v.sum_logprobs = full.sum_logprobs + slot;
  const int len = S + P + 1;
  const int64_t R = (e->attn_t5 && e->mode == VX_BF16 && !e->force_simt) ? (len + 127) / 128 * 128 : len;
  const int32_t last = S + P;
  CK(cudaEventRecord(e->ev0, st));
  e->bt = v;
  e->kv_slot_off = (size_t)slot * kH * e->ctx_cap * kDH;
  r = prefill_core(e, 1, R, len, &last, (float*)e->logits.p + (size_t)slot * kARLogits, st);
  e->bt = full;
  e->kv_slot_off = 0;
  if (r) return r;
  CK(cudaMemsetAsync((int32_t*)e->d_T.p + slot, 0, 4, st));
  CK(cudaMemsetAsync((int32_t*)e->d_fin.p + slot, 0, 4, st));
  CK(cudaMemsetAsync((float*)e->d_slp.p + slot, 0, 4, st));
  CK(cudaMemcpyAsync((int32_t*)e->d_pending.p + slot, e->h_pinned + 2, 4, cudaMemcpyHostToDevice, st));
  CK(cudaEventRecord(e->ev1, st));
  CK(cudaStreamSynchronize(st));
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e->ev0, e->ev1));
  e->ms_prefill += ms;
  CK(cudaGetLastError());
  e->hS[slot] = S; e->hP[slot] = P;
  if ((int)e->hfin.size() == e->B) e->hfin[slot] = 0;
  if ((int)e->hT.size() == e->B) e->hT[slot] = 0;
  return VX_OK;
}

int vx_ar_fork(vx_engine* e, int n_beams, void* stream) {
  if (!e || !e->prefilled || e->decoded) return fail(VX_ERR_STATE, "vx_ar_fork: call right after vx_ar_prefill");
  if (e->B != 1) return fail(VX_ERR_STATE, "vx_ar_fork: the batch must hold exactly one utterance");
  if (n_beams < 1 || n_beams > e->max_batch) return fail(VX_ERR_CAPACITY, "vx_ar_fork: n_beams exceeds max_batch");
  if (n_beams == 1) return VX_OK;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int n = e->hS[0] + e->hP[0] + 1;   // keys in the cache after the prefill
  if (e->mode == VX_BF16)
    fork_kv_kernel<bf16><<<dim3(kL * kH, n_beams - 1), 256, 0, st>>>((bf16*)e->kc.p, (bf16*)e->vc.p, e->kv_layer_elems, e->ctx_cap, n);
  else
    fork_kv_kernel<float><<<dim3(kL * kH, n_beams - 1), 256, 0, st>>>((float*)e->kc.p, (float*)e->vc.p, e->kv_layer_elems, e->ctx_cap, n);
  fork_state_kernel<<<n_beams - 1, 256, 0, st>>>(e->bt, n_beams, e->hP[0] + 1, (float*)e->logits.p, (float*)e->xs.p, (int32_t*)e->d_S.p,
                                                (int32_t*)e->d_P.p, (int32_t*)e->d_ar_row0.p);
  e->launches += 2;
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  e->hS.assign(n_beams, e->hS[0]);
  e->hP.assign(n_beams, e->hP[0]);
  e->h_ar_row0.assign(n_beams, e->h_ar_row0[0]);
  e->B = n_beams;
  e->bt.B = n_beams;
  e->forked = true;
  return VX_OK;
}

int vx_ar_select_beam(vx_engine* e, float length_penalty, int return_worst, int* out_pick, void* stream) {
  if (!e || !e->decoded) return fail(VX_ERR_STATE, "vx_ar_select_beam: call after vx_ar_decode");
  if (!e->forked && e->B != 1) return fail(VX_ERR_STATE, "vx_ar_select_beam: the batch does not hold the beams of one utterance (vx_ar_fork)");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  ENS(e->d_flag, 16);
  select_beam_kernel<<<1, 256, 0, st>>>(e->bt, e->B, length_penalty, return_worst, (int32_t*)e->d_flag.p);
  e->launches++;
  CK(cudaMemcpyAsync(e->h_pinned, e->d_flag.p, 8, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaGetLastError());
  if (out_pick) *out_pick = e->h_pinned[0];
  e->hS.resize(1); e->hP.resize(1); e->h_ar_row0.resize(1);
  e->hT.assign(1, e->h_pinned[1]);
  e->B = 1;
  e->bt.B = 1;
  e->forked = false;
  return VX_OK;
}

int vx_ar_budget(vx_engine* e, int max_new, int* out_max_steps) {
  if (!e || !e->batch_set || !out_max_steps) return fail(VX_ERR_STATE, "vx_ar_budget: no batch set");
  int m = 0;
  for (int b = 0; b < e->B; ++b) {
    int bud = 16 * e->hS[b];  // models/vallex.py:577
    if (max_new > 0) bud = std::min(bud, max_new);
    bud = std::min(bud, e->max_ctx - (e->hS[b] + e->hP[b] + 1) - 1);
    m = std::max(m, bud);
  }
  *out_max_steps = m;
  return VX_OK;
}

// One decode forward pass for the tokens just appended (all utterances at once).
static int decode_forward(vx_engine* e, int nsplit, cudaStream_t st) {
  const Batch& bt = e->bt;
  const int B = e->B, wt = e->mode;
  const size_t kvb = kv_esize(e);
  float* xs = (float*)e->xs.p; float* xsn = (float*)e->xsn.p; float* qkvs = (float*)e->qkvs.p;
  float* atts = (float*)e->atts.p; float* ffs = (float*)e->ffs.p;
  launch_embed_audio_step(bt, vec(e, "ar_audio_embedding.word_embeddings.weight"), vec(e, "ar_audio_position.alpha"), e->pe, xs, st);
  e->launches++;
  const bool small = B <= 8;
  for (int l = 0; l < kL; ++l) {
    const LayerW& L = e->ar[l];
    if (small) {
      launch_gemv(xs, kD, L.qkv_w, wt, L.qkv_b, qkvs, 3 * kD, B, 3 * kD, kD, EPI_NONE, L.n1_g, L.n1_b, st);
      e->launches++;
    } else {
      launch_layernorm(xs, nullptr, B, kD, L.n1_g, L.n1_b, nullptr, nullptr, 1e-5f, xsn, st);
      launch_gemm(xsn, kD, L.qkv_w, wt, L.qkv_b, qkvs, 3 * kD, B, 3 * kD, kD, EPI_NONE, nullptr, st);
      e->launches += 2;
    }
    launch_attn_decode(bt, qkvs, (char*)e->kc.p + (size_t)l * e->kv_layer_elems * kvb,
                       (char*)e->vc.p + (size_t)l * e->kv_layer_elems * kvb, wt, e->ctx_cap, atts, (float*)e->part.p,
                       (int32_t*)e->tick.p, nsplit, st);
    e->launches++;
    linear(e, atts, kD, L.out_w, wt, L.out_b, xs, kD, B, kD, kD, EPI_RESIDUAL, st);
    if (small) {
      launch_gemv(xs, kD, L.ff1_w, wt, L.ff1_b, ffs, kFF, B, kFF, kD, EPI_RELU, L.n2_g, L.n2_b, st);
      e->launches++;
    } else {
      launch_layernorm(xs, nullptr, B, kD, L.n2_g, L.n2_b, nullptr, nullptr, 1e-5f, xsn, st);
      launch_gemm(xsn, kD, L.ff1_w, wt, L.ff1_b, ffs, kFF, B, kFF, kD, EPI_RELU, nullptr, st);
      e->launches += 2;
    }
    linear(e, ffs, kFF, L.ff2_w, wt, L.ff2_b, xs, kD, B, kD, kFF, EPI_RESIDUAL, st);
  }
  if (small) {
    launch_gemv(xs, kD, e->ar_predict, wt, nullptr, (float*)e->logits.p, kARLogits, B, kARLogits, kD, EPI_NONE,
                e->ar_norm_g, e->ar_norm_b, st);
    e->launches++;
  } else {
    launch_layernorm(xs, nullptr, B, kD, e->ar_norm_g, e->ar_norm_b, nullptr, nullptr, 1e-5f, xsn, st);
    launch_gemm(xsn, kD, e->ar_predict, wt, nullptr, (float*)e->logits.p, kARLogits, B, kARLogits, kD, EPI_NONE, nullptr, st);
    e->launches += 2;
  }
  return VX_OK;
}

int vx_ar_decode(vx_engine* e, int max_new, int top_k, float top_p, float temperature, uint64_t seed, int ignore_eos,
                 const int32_t* forced_tokens, int32_t* out_tokens, int32_t* out_lens, float* out_sum_logprobs, int mem,
                 void* stream) {
  return vx_ar_decode_ex(e, max_new, top_k, top_p, temperature, seed, ignore_eos, forced_tokens, out_tokens, out_lens, out_sum_logprobs,
                         0, 0, nullptr, mem, stream);
}

int vx_ar_decode_ex(vx_engine* e, int max_new, int top_k, float top_p, float temperature, uint64_t seed, int ignore_eos,
                    const int32_t* forced_tokens, int32_t* out_tokens, int32_t* out_lens, float* out_sum_logprobs, int resume,
                    int stop_on_finish, int32_t* out_finished, int mem, void* stream) {
  if (!e || !e->prefilled) return fail(VX_ERR_STATE, "vx_ar_decode: call vx_ar_prefill first");
  if (resume && !e->decoded) return fail(VX_ERR_STATE, "vx_ar_decode_ex: resume needs a previous decode launch on this batch");
  if ((resume || stop_on_finish) && (e->decode_graph || e->B > 32)) return fail(VX_ERR_STATE, "vx_ar_decode_ex: resume / stop_on_finish need the persistent decode kernel");
  if (!resume) e->hfin.assign(e->B, 0);
  if (!(temperature > 0.f)) return fail(VX_ERR_ARG, "vx_ar_decode: temperature must be > 0");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int B = e->B;
  int max_steps = 0;
  vx_ar_budget(e, max_new, &max_steps);
  if (max_steps < 1) return fail(VX_ERR_CAPACITY, "vx_ar_decode: no room for new tokens (max_ctx too small)");
  std::vector<int32_t> bud(B);
  int final_ctx = 0;
  e->ctx_clamped = false;
  for (int b = 0; b < B; ++b) {
    int v = 16 * e->hS[b];
    if (max_new > 0) v = std::min(v, max_new);
    const int room = e->max_ctx - (e->hS[b] + e->hP[b] + 1) - 1;
    if (room < v) { v = room; e->ctx_clamped = true; }   // the KV cache, not EOS / 16*S / max_new, will end this utterance
    bud[b] = v;
    final_ctx = std::max(final_ctx, e->hS[b] + e->hP[b] + 1 + v);
  }
  int nsplit = std::max(1, (2 * kSMs + B * kH - 1) / (B * kH));
  nsplit = std::max(nsplit, (final_ctx + 2047) / 2048);
  nsplit = std::min(nsplit, 64);
  ENS(e->qkvs, (size_t)B * 3 * kD * 4); ENS(e->atts, (size_t)B * kD * 4); ENS(e->ffs, (size_t)B * kFF * 4);
  ENS(e->part, (size_t)B * kH * nsplit * (kDH + 2) * 4); ENS(e->tick, ((size_t)B * kH + 64) * 4);   // + FFN2 fold tickets of the batched kernel
  ENS(e->out_tok, (size_t)B * max_steps * 4);
  if (forced_tokens) {
    ENS(e->forced, (size_t)B * max_steps * 4);
    int r = copy_in(e->forced.p, forced_tokens, (size_t)B * max_steps * 4, mem, st);
    if (r) return r;
  }
  memcpy(e->h_pinned, bud.data(), B * 4);
  int n_live = 0;
  for (int b = 0; b < B; ++b) n_live += e->hfin[b] ? 0 : 1;
  if (n_live == 0) return fail(VX_ERR_STATE, "vx_ar_decode_ex: every slot has finished (refill one with vx_slot_prefill)");
  e->h_pinned[B] = n_live;
  CK(cudaMemcpyAsync(e->d_budget.p, e->h_pinned, B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->d_nact.p, e->h_pinned + B, 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemsetAsync(e->tick.p, 0, ((size_t)B * kH + 64) * 4, st));
  CK(cudaMemsetAsync(e->out_tok.p, 0, (size_t)B * max_steps * 4, st));
  CK(cudaStreamSynchronize(st));  // h_pinned reused below
  const int32_t* forced_dev = forced_tokens ? (const int32_t*)e->forced.p : nullptr;

  if (!e->decode_graph && B <= 32) {
    // ---- persistent kernel: the whole token loop in one cooperative launch -------------------------------
    int ns = std::max(1, e->num_sms / (B * kH));
    if (getenv("VX_DECODE_NSPLIT")) ns = std::max(1, atoi(getenv("VX_DECODE_NSPLIT")));
    ns = std::max(ns, (final_ctx + 2047) / 2048);
    ns = std::min(ns, 64);
    ENS(e->part, (size_t)B * kH * ns * (kDH + 2) * 4);
    ENS(e->p_pf, (size_t)4 * B * kD * 4);
    ENS(e->p_bar, 64);
    CK(cudaMemsetAsync(e->p_bar.p, 0, 64, st));
    { const int32_t big = 0x7fffffff; CK(cudaMemcpyAsync((char*)e->p_bar.p + 40, &big, 4, cudaMemcpyHostToDevice, st)); }
    PParams pp{};
    pp.bt = e->bt;
    for (int l = 0; l < kL; ++l) {
      const LayerW& L = e->ar[l];
      pp.layer[l] = PLayer{L.qkv_w, L.out_w, L.ff1_w, L.ff2_w, L.ff2_t, L.qkv_b, L.out_b, L.ff1_b, L.ff2_b, L.n1_g, L.n1_b, L.n2_g, L.n2_b};
    }
    // 3-barrier variant: one attention unit per CTA, bf16 weights, B <= 8; anything else takes the 5-barrier kernel
    const char* mm_env = getenv("VX_DECODE_MMA");   // VX_DECODE_MMA=0: CUDA-core GEMV kernel instead of the tensor-core batched one
    const bool batched = e->mode == VX_BF16 && B > 4 && e->num_sms >= 128 && e->num_sms % 4 == 0 && !(mm_env && mm_env[0] == '0');
    pp.fused = (!batched && e->decode_fused && e->mode == VX_BF16 && B <= e->decode_fused_maxb && B * kH * ns <= e->num_sms &&
                (kFF + e->num_sms - 1) / e->num_sms <= 64 && e->ar[0].ff2_t) ? 1 : 0;
    {   // batch 1: head groups as 4-CTA clusters, two grid barriers per layer (decode_cluster.cu; default); VX_DECODE_CLUSTER=0 keeps
        // the 3-barrier kernel, as does a device that cannot hold the 16 head clusters at once
      const char* ce = getenv("VX_DECODE_CLUSTER");
      if (e->decode_cluster_grid < 0) e->decode_cluster_grid = decode_cluster_grid(e->num_sms);
      pp.cluster = (pp.fused && B == 1 && e->decode_cluster_grid > 0 && !(ce && ce[0] == '0')) ? 1 : 0;
    }
    if (pp.fused) {
      ENS(e->p_xb, (size_t)B * kD * 4);
      ENS(e->p_acc, (size_t)4 * B * kD * 8);
      CK(cudaMemsetAsync(e->p_acc.p, 0, (size_t)4 * B * kD * 8, st));
      CK(cudaMemsetAsync(e->tick.p, 0, (size_t)B * kH * 4, st));
      pp.xb = (float*)e->p_xb.p;
      pp.acc_out = (long long*)e->p_acc.p;
      pp.acc_ffn = (long long*)e->p_acc.p + (size_t)2 * B * kD;
    }
    if (batched) {   // batched tensor-core kernel: balanced warp-level attention phase
      pp.att_maxp = e->num_sms + 2;   // pieces of one (utterance, head): <= warps/16 + 2
      pp.kv_l2_mb = getenv("VX_KV_L2_MB") ? atoi(getenv("VX_KV_L2_MB")) : 0;   // measured: no gain (the 3 GB a step streams do not stay in L2), off by default
      ENS(e->part, (size_t)B * kH * pp.att_maxp * (kDH + 4) * 4);
    }
    pp.norm_g = e->ar_norm_g; pp.norm_b = e->ar_norm_b; pp.predict_w = e->ar_predict;
    pp.audio_emb = vec(e, "ar_audio_embedding.word_embeddings.weight"); pp.alpha = vec(e, "ar_audio_position.alpha"); pp.pe = e->pe;
    pp.kc = e->kc.p; pp.vc = e->vc.p; pp.kv_layer_elems = e->kv_layer_elems; pp.ctx_cap = e->ctx_cap;
    pp.xa = (float*)e->xs.p; pp.pf = (float*)e->p_pf.p; pp.qkv = (float*)e->qkvs.p; pp.att = (float*)e->atts.p;
    pp.hid = (float*)e->ffs.p; pp.logits = (float*)e->logits.p; pp.part = (float*)e->part.p; pp.tickets = (int32_t*)e->tick.p;
    pp.bar = (unsigned int*)e->p_bar.p; pp.nsplit = ns;
    pp.top_k = top_k; pp.top_p = top_p; pp.temperature = temperature; pp.seed = seed; pp.ignore_eos = ignore_eos;
    pp.forced = forced_dev; pp.forced_stride = max_steps; pp.out_tokens = (int32_t*)e->out_tok.p; pp.out_stride = max_steps;
    pp.max_steps = max_steps; pp.steps_done = (int32_t*)((char*)e->p_bar.p + 32);
    pp.resume = resume ? 1 : 0; pp.pending = (const int32_t*)e->d_pending.p;
    pp.stop_below = (stop_on_finish && n_live > 1) ? n_live - 1 : 0;
    pp.dbg = getenv("VX_DECODE_DBG") ? atoi(getenv("VX_DECODE_DBG")) : 0;
    if (getenv("VX_DECODE_PROF")) {
      ENS(e->p_prof, 3 * 4096 * 8);
      CK(cudaMemsetAsync(e->p_prof.p, 0, 3 * 4096 * 8, st));
      pp.prof = (long long*)e->p_prof.p;
      pp.prof_step0 = std::max(0, std::min(max_steps - 6, atoi(getenv("VX_DECODE_PROF"))));
    }
    {
      const int mr = B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : B <= 8 ? 8 : B <= 16 ? 16 : 32;
      if (pp.att_maxp > 0) e->decode_kernel = "ar_decode_mma_kernel<" + std::to_string(mr < 8 ? 8 : mr) + ">";
      else if (pp.cluster) e->decode_kernel = "ar_decode_cluster_kernel";
      else if (pp.fused) e->decode_kernel = "ar_decode_fused_kernel<" + std::to_string(mr) + ">";
      else e->decode_kernel = std::string("ar_decode_persistent_kernel<") + (e->mode == VX_BF16 ? "bf16," : "float,") + std::to_string(mr) + ">";
    }
    CK(cudaEventRecord(e->ev0, st));
    const int rc = launch_decode_persistent(&pp, e->mode, B, e->num_sms, st);
    if (rc != 0) return fail(VX_ERR_CUDA, std::string("persistent decode launch: ") + cudaGetErrorString((cudaError_t)rc));
    e->launches++;
    CK(cudaEventRecord(e->ev1, st));
    CK(cudaMemcpyAsync(e->h_pinned, pp.steps_done, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->ms_decode, e->ev0, e->ev1));
    e->decode_steps = e->h_pinned[0];
    CK(cudaGetLastError());
    if (pp.dbg & 32) {
      int32_t probe[2];
      CK(cudaMemcpy(probe, (char*)e->p_bar.p + 40, 8, cudaMemcpyDeviceToHost));
      fprintf(stderr, "[vx decode probe] first non-finite code %d (step*1000 + layer*10 + site; 9e8+h = split combine), bits 0x%08x\n", probe[0], (unsigned)probe[1]);
    }
    int r2;
    if (out_tokens && (r2 = copy_out(out_tokens, e->out_tok.p, (size_t)B * max_steps * 4, mem, st))) return r2;
    if (out_lens && (r2 = copy_out(out_lens, e->d_T.p, B * 4, mem, st))) return r2;
    if (out_sum_logprobs && (r2 = copy_out(out_sum_logprobs, e->d_slp.p, B * 4, mem, st))) return r2;
    e->hT.assign(B, 0);
    CK(cudaMemcpyAsync(e->hT.data(), e->d_T.p, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(e->hfin.data(), e->d_fin.p, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemsetAsync(e->d_pending.p, 0, e->max_batch * 4, st));   // every live slot has now sampled the logits it had
    if (out_finished && (r2 = copy_out(out_finished, e->d_fin.p, B * 4, mem, st))) return r2;
    CK(cudaStreamSynchronize(st));
    e->decoded = true;
    return VX_OK;
  }

  // capture one iteration [sample -> forward] into a CUDA graph; kernels read all step state from the device
  e->decode_kernel = "cuda-graph step (sample_kernel + gemv/attn_decode kernels)";
  cudaGraph_t graph = nullptr;
  cudaGraphExec_t exec = nullptr;
  const int64_t launches_before = e->launches;
  CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
  launch_sample(e->bt, (const float*)e->logits.p, top_k, top_p, temperature, seed, 0, ignore_eos, forced_dev, max_steps,
                (int32_t*)e->out_tok.p, max_steps, st);
  e->launches++;
  decode_forward(e, nsplit, st);
  CK(cudaStreamEndCapture(st, &graph));
  CK(cudaGraphInstantiate(&exec, graph, 0));
  const int64_t per_iter = e->launches - launches_before;
  e->launches = launches_before;

  CK(cudaEventRecord(e->ev0, st));
  int steps = 0;
  const bool can_stop_early = !ignore_eos;
  for (int s = 0; s < max_steps; ++s) {
    if (s == max_steps - 1) {  // last token: sample only, the forward pass would be discarded
      launch_sample(e->bt, (const float*)e->logits.p, top_k, top_p, temperature, seed, 0, ignore_eos, forced_dev,
                    max_steps, (int32_t*)e->out_tok.p, max_steps, st);
      e->launches++;
      steps++;
      break;
    }
    CK(cudaGraphLaunch(exec, st));
    e->launches += per_iter;
    steps++;
    if (can_stop_early && (s % 32) == 31) {
      CK(cudaMemcpyAsync(e->h_pinned, e->d_nact.p, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (e->h_pinned[0] <= 0) break;
    }
  }
  CK(cudaEventRecord(e->ev1, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&e->ms_decode, e->ev0, e->ev1));
  e->decode_steps = steps;
  CK(cudaGraphExecDestroy(exec));
  CK(cudaGraphDestroy(graph));
  CK(cudaGetLastError());
  int r;
  if (out_tokens && (r = copy_out(out_tokens, e->out_tok.p, (size_t)B * max_steps * 4, mem, st))) return r;
  if (out_lens && (r = copy_out(out_lens, e->d_T.p, B * 4, mem, st))) return r;
  if (out_sum_logprobs && (r = copy_out(out_sum_logprobs, e->d_slp.p, B * 4, mem, st))) return r;
  e->hT.assign(B, 0);
  CK(cudaMemcpyAsync(e->hT.data(), e->d_T.p, B * 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  e->decoded = true;
  return VX_OK;
}

int vx_nar(vx_engine* e, const int32_t* cb0, const int32_t* t_lens, int Tmax, int32_t* out_codes, int mem,
           void* stream) {
  if (!e || !e->batch_set) return fail(VX_ERR_STATE, "vx_nar: no batch set");
  if (!out_codes || Tmax < 0) return fail(VX_ERR_ARG, "vx_nar: bad arguments");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int B = e->B, wt = e->mode;
  const Batch& bt = e->bt;
  if (cb0) {
    if (!t_lens) return fail(VX_ERR_ARG, "vx_nar: t_lens required with cb0");
    ENS(e->tl_dev, B * 4); ENS(e->cb0_dev, (size_t)B * std::max(Tmax, 1) * 4);
    int r;
    if ((r = copy_in(e->tl_dev.p, t_lens, B * 4, mem, st))) return r;
    if ((r = copy_in(e->cb0_dev.p, cb0, (size_t)B * Tmax * 4, mem, st))) return r;
    e->hT.assign(B, 0);
    CK(cudaMemcpyAsync(e->hT.data(), e->tl_dev.p, B * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (int b = 0; b < B; ++b)
      if (e->hT[b] < 0 || e->hT[b] > Tmax || e->hS[b] + e->hP[b] + e->hT[b] + 1 > e->max_ctx)
        return fail(VX_ERR_CAPACITY, "vx_nar: t_lens out of range for max_ctx");
    if (Tmax > 0) {
      ENS(e->d_flag, 16);
      CK(cudaMemsetAsync(e->d_flag.p, 0, 16, st));
      check_range_kernel<<<dim3(8, B), 256, 0, st>>>((const int32_t*)e->cb0_dev.p, Tmax, (const int32_t*)e->tl_dev.p, 1, 0, 0, kVocab, 1, (int32_t*)e->d_flag.p);
      CK(cudaMemcpyAsync(e->h_pinned, e->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
      CK(cudaStreamSynchronize(st));
      if (e->h_pinned[0] != 0) return fail(VX_ERR_ARG, "vx_nar: cb0 token out of range (0..1023)");
    }
    launch_init_y(bt, st);
    if (Tmax > 0) write_cb0_kernel<<<dim3((Tmax + 255) / 256, B), 256, 0, st>>>(bt, (const int32_t*)e->cb0_dev.p, (const int32_t*)e->tl_dev.p, Tmax);
    else CK(cudaMemsetAsync(e->d_T.p, 0, B * 4, st));
    e->launches += 2;
  } else if (!e->decoded) {
    return fail(VX_ERR_STATE, "vx_nar: no codebook-0 tokens (pass cb0 or run vx_ar_decode)");
  }
  std::vector<int32_t> nrow0(B), yrow0(B), trow0(B);
  int64_t R = 0, YR = 0, TR = 0;
  int maxlen = 0, maxpt = 0, maxT = 0;
  for (int b = 0; b < B; ++b) {
    if (e->hT[b] > Tmax) return fail(VX_ERR_ARG, "vx_nar: Tmax smaller than a generated length");
    nrow0[b] = (int32_t)R; yrow0[b] = (int32_t)YR; trow0[b] = (int32_t)TR;
    const int pt = e->hP[b] + e->hT[b];
    R += (e->attn_t5 && e->mode == VX_BF16 && !e->force_simt) ? (e->hS[b] + pt + 127) / 128 * 128 : e->hS[b] + pt;
    YR += pt; TR += e->hT[b];
    maxlen = std::max(maxlen, e->hS[b] + pt); maxpt = std::max(maxpt, pt); maxT = std::max(maxT, e->hT[b]);
  }
  // utterance chunks for the layer stack (bf16 + tcgen05 attention + QKV-heads epilogue only; otherwise one chunk)
  std::vector<int> chunk_b{0};
  const std::vector<int32_t> nrow0_h = nrow0;
  {
    const bool can_chunk = e->mode == VX_BF16 && !e->force_simt && e->attn_t5 && e->attn_t5_fuse;
    const char* ce = getenv("VX_NAR_CHUNK_ROWS");
    const int64_t chunk_rows = ce ? atoll(ce) : 23040;   // 180 row tiles x 4 column tiles of the N = 1024 GEMMs = 720 = 4.86 waves
    if (can_chunk && chunk_rows > 0) {
      int64_t start = 0;
      for (int b = 1; b < B; ++b)
        if ((int64_t)nrow0[b] + ((b + 1 < B ? (int64_t)nrow0[b + 1] : R) - nrow0[b]) - start > chunk_rows && nrow0[b] > start) { chunk_b.push_back(b); start = nrow0[b]; }
    }
    chunk_b.push_back(B);
  }
  std::vector<int32_t> tgt((size_t)std::max<int64_t>(TR, 1));
  for (int b = 0; b < B; ++b)
    for (int t = 0; t < e->hT[b]; ++t) tgt[trow0[b] + t] = nrow0[b] + e->hS[b] + e->hP[b] + t;
  ENS(e->yrow0, B * 4); ENS(e->trow0, B * 4); ENS(e->tgtmap, tgt.size() * 4);
  CK(cudaMemcpyAsync(e->d_nar_row0.p, nrow0.data(), B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->yrow0.p, yrow0.data(), B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->trow0.p, trow0.data(), B * 4, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(e->tgtmap.p, tgt.data(), tgt.size() * 4, cudaMemcpyHostToDevice, st));
  { int _r = ensure_full_buffers(e, R); if (_r) return _r; }
  ENS(e->xn, (std::max<int64_t>(TR, 1) + 127) / 128 * 128 * kD * 4);
  ENS(e->t_img, (std::max<int64_t>(TR, 1) + 127) / 128 * 128 * kD * 2);
  ENS(e->yemb, std::max<int64_t>(YR, 1) * kD * 4); ENS(e->nlog, std::max<int64_t>(TR, 1) * kVocab * 4);
  ENS(e->codes, (size_t)B * std::max(Tmax, 1) * kNQ * 4);
  CK(cudaStreamSynchronize(st));  // host vectors above go out of scope; x/xn may have been reallocated
  CK(cudaEventRecord(e->ev0, st));
  float* x = (float*)e->x.p; float* xn = (float*)e->xn.p; float* yemb = (float*)e->yemb.p;
  const int32_t* d_yrow0 = (const int32_t*)e->yrow0.p; const int32_t* d_trow0 = (const int32_t*)e->trow0.p;
  CK(cudaMemsetAsync(e->codes.p, 0, (size_t)B * std::max(Tmax, 1) * kNQ * 4, st));
  launch_nar_copy_cb0(bt, (int32_t*)e->codes.p, Tmax, st);
  launch_nar_embed_init(bt, e->nar_emb, yemb, d_yrow0, maxpt, st);
  e->launches += 2;
  if (TR > 0) {
    for (int stage = 0; stage < kNQ - 1; ++stage) {
      if (e->attn_t5 && e->mode == VX_BF16 && !e->force_simt) CK(cudaMemsetAsync(x, 0, (size_t)e->full_rows * kD * 4, st));
      launch_embed_text(bt, vec(e, "nar_text_embedding.word_embeddings.weight"), vec(e, "nar_language_embedding.word_embeddings.weight"),
                        vec(e, "nar_text_position.alpha"), e->pe, bt.nar_row0, x, st);
      launch_nar_build_x(bt, yemb, d_yrow0, vec(e, "nar_audio_position.alpha"), e->pe, x, maxpt, st);
      e->launches += 2;
      // The 12 layers run chunk by chunk over the utterances (independent: models/vallex.py:491 has batch 1): a chunk's residual
      // stream and operand images stay L2-resident from kernel to kernel, which a 64-utterance pass (0.5 GB of residual stream
      // alone) does not; chunks are sized by rows so the N = 1024 GEMMs still fill whole waves of 148 CTAs.
      for (size_t c = 0; c + 1 < chunk_b.size(); ++c) {
        const int cb0 = chunk_b[c], cb1 = chunk_b[c + 1];
        const int64_t cr0 = nrow0_h[cb0], cr1 = (cb1 < B) ? nrow0_h[cb1] : R;
        for (int l = 0; l < kL; ++l) {
          const LayerW& L = e->nar[l];
          const int rc = full_layer(e, L, cr1 - cr0, bt.nar_row0, 0, maxlen, L.ada1 + (size_t)stage * 2 * kD, L.ada2 + (size_t)stage * 2 * kD, -1, st,
                                    cr0, cb0, cb1 - cb0);
          if (rc) return rc;
        }
      }
      const float* af = e->nar_norm_ada + (size_t)stage * 2 * kD;
      if (e->mode == VX_BF16 && !e->force_simt) {
        launch_layernorm_image(x, (const int32_t*)e->tgtmap.p, TR, e->nar_norm_g, e->nar_norm_b, af, af + kD, 1e-5f, e->t_img.p, st);
        launch_gemm_tc(e->t_img.p, e->nar_predict_img[stage], nullptr, e->nlog.p, kVocab, TR, kVocab, kD, TC_EPI_F32, st);
        e->launches += 2;
      } else {
        launch_layernorm(x, (const int32_t*)e->tgtmap.p, TR, kD, e->nar_norm_g, e->nar_norm_b, af, af + kD, 1e-5f, xn, st);
        e->launches++;
        linear(e, xn, kD, e->nar_predict[stage], wt, nullptr, (float*)e->nlog.p, kVocab, TR, kVocab, kD, EPI_NONE, st);
      }
      launch_nar_argmax_update(bt, (const float*)e->nlog.p, d_trow0, stage, e->nar_emb, yemb, d_yrow0,
                               (int32_t*)e->codes.p, Tmax, st);
      e->launches++;
    }
  }
  CK(cudaEventRecord(e->ev1, st));
  int r;
  if ((r = copy_out(out_codes, e->codes.p, (size_t)B * Tmax * kNQ * 4, mem, st))) return r;
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&e->ms_nar, e->ev0, e->ev1));
  CK(cudaGetLastError());
  return VX_OK;
}

int vx_vocos_decode(vx_engine* e, const int32_t* codes, int B, int T, int bandwidth_id, float* wave, int mem,
                    void* stream) {
  if (!e || !e->finalized || !e->has_vocos) return fail(VX_ERR_STATE, "vx_vocos_decode: Vocos weights not loaded");
  if (!codes || !wave || B < 1 || T < 1 || bandwidth_id < 0 || bandwidth_id > 3) return fail(VX_ERR_ARG, "vx_vocos_decode: bad arguments");
  CK(cudaSetDevice(e->device));
  cudaStream_t st = pick(e, stream);
  const int64_t R = (int64_t)B * T;
  ENS(e->v_codes, (size_t)kNQ * R * 4); ENS(e->v_feat, R * 128 * 4); ENS(e->v_col, R * 896 * 4);
  ENS(e->v_x, R * 384 * 4); ENS(e->v_h, R * 384 * 4); ENS(e->v_ff, R * 1152 * 4); ENS(e->v_head, R * 1282 * 4);
  ENS(e->v_spec, R * 1296 * 4); ENS(e->v_frames, R * 1280 * 4); ENS(e->v_wave, R * 320 * 4);
  CK(cudaEventRecord(e->ev0, st));
  int r;
  if ((r = copy_in(e->v_codes.p, codes, (size_t)kNQ * R * 4, mem, st))) return r;
  ENS(e->d_flag, 16);
  CK(cudaMemsetAsync(e->d_flag.p, 0, 16, st));   // checked after the pass (the gather clamps, so a bad code cannot fault)
  check_range_kernel<<<dim3(148, 1), 256, 0, st>>>((const int32_t*)e->v_codes.p, 0, nullptr, 1, (int64_t)kNQ * R, 0, 1024, 1, (int32_t*)e->d_flag.p);
  auto V = [&](const std::string& k) { return vec(e, "vocos." + k); };
  float* feat = (float*)e->v_feat.p; float* col = (float*)e->v_col.p; float* x = (float*)e->v_x.p;
  float* h = (float*)e->v_h.p; float* ff = (float*)e->v_ff.p; float* head = (float*)e->v_head.p;
  float* spec = (float*)e->v_spec.p; float* frames = (float*)e->v_frames.p; float* wv = (float*)e->v_wave.p;
  launch_vocos_features((const int32_t*)e->v_codes.p, B, T, V("feature_extractor.codebook_weights"), feat, st);
  if (e->vtc.ready) {
    // ---- tensor-core path: every contraction on the tcgen05 GEMM with bf16 x 3 split operands (vocos.cu) ------------------
    const int64_t Rp = (R + 127) / 128 * 128;
    ENS(e->v_imgA, (size_t)Rp * 4032 * 2); ENS(e->v_imgB, (size_t)Rp * 3456 * 2);
    ENS(e->v_head, (size_t)R * 1408 * 4);
    head = (float*)e->v_head.p;
    void *ia = e->v_imgA.p, *ib = e->v_imgB.p;
    launch_vocos_im2col_split3(feat, B, T, ia, st);
    launch_gemm_tc(ia, e->vtc.embed, V("backbone.embed.bias"), x, 384, R, 384, 3 * 896, TC_EPI_F32, st);
    launch_layernorm(x, nullptr, R, 384, nullptr, nullptr, V("backbone.norm.scale.weight") + bandwidth_id * 384,
                     V("backbone.norm.shift.weight") + bandwidth_id * 384, 1e-6f, x, st);
    e->launches += 4;
    for (int i = 0; i < 8; ++i) {
      const std::string p = "backbone.convnext." + std::to_string(i);
      launch_vocos_dwconv_ln_split3(x, B, T, e->vtc.dwT[i], V(p + ".dwconv.bias"), V(p + ".norm.scale.weight") + bandwidth_id * 384,
                                    V(p + ".norm.shift.weight") + bandwidth_id * 384, ia, st);
      launch_gemm_tc(ia, e->vtc.pw1[i], V(p + ".pwconv1.bias"), ib, 0, R, 1152, 3 * 384, TC_EPI_GELU_SPLIT3, st);
      launch_gemm_tc(ib, e->vtc.pw2[i], V(p + ".pwconv2.bias"), x, 384, R, 384, 3 * 1152, TC_EPI_F32_SCALE_RESIDUAL, st, V(p + ".gamma"));
      e->launches += 3;
    }
    launch_vocos_ln_split3(x, R, V("backbone.final_layer_norm.weight"), V("backbone.final_layer_norm.bias"), ia, st);
    launch_gemm_tc(ia, e->vtc.head, e->vtc.head_bias, head, 1408, R, 1408, 3 * 384, TC_EPI_F32, st);
    launch_vocos_spectrum_split3(head, 1408, R, ia, st);
    launch_gemm_tc(ia, e->vtc.dft, nullptr, frames, 1280, R, 1280, 3 * 1344, TC_EPI_F32, st);
    launch_vocos_overlap_add(frames, B, T, V("head.istft.window"), wv, st);
    e->launches += 5;
    if ((r = copy_out(wave, wv, (size_t)R * 320 * 4, mem, st))) return r;
    CK(cudaEventRecord(e->ev1, st));
    CK(cudaMemcpyAsync(e->h_pinned, e->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    CK(cudaEventElapsedTime(&e->ms_vocos, e->ev0, e->ev1));
    CK(cudaGetLastError());
    if (e->h_pinned[0] != 0) return fail(VX_ERR_ARG, "vx_vocos_decode: code out of range (0..1023)");
    return VX_OK;
  }
  launch_vocos_im2col7(feat, B, T, 128, col, st);
  launch_gemm(col, 896, V("__embed_im2col"), 0, V("backbone.embed.bias"), x, 384, R, 384, 896, EPI_NONE, nullptr, st);
  launch_layernorm(x, nullptr, R, 384, nullptr, nullptr, V("backbone.norm.scale.weight") + bandwidth_id * 384,
                   V("backbone.norm.shift.weight") + bandwidth_id * 384, 1e-6f, x, st);
  e->launches += 4;
  for (int i = 0; i < 8; ++i) {
    const std::string p = "backbone.convnext." + std::to_string(i);
    launch_vocos_dwconv7(x, B, T, 384, V(p + ".dwconv.weight"), V(p + ".dwconv.bias"), h, st);
    launch_layernorm(h, nullptr, R, 384, nullptr, nullptr, V(p + ".norm.scale.weight") + bandwidth_id * 384,
                     V(p + ".norm.shift.weight") + bandwidth_id * 384, 1e-6f, h, st);
    launch_gemm(h, 384, V(p + ".pwconv1.weight"), 0, V(p + ".pwconv1.bias"), ff, 1152, R, 1152, 384, EPI_GELU, nullptr, st);
    launch_gemm(ff, 1152, V(p + ".pwconv2.weight"), 0, V(p + ".pwconv2.bias"), x, 384, R, 384, 1152, EPI_SCALE_RESIDUAL,
                V(p + ".gamma"), st);
    e->launches += 4;
  }
  launch_layernorm(x, nullptr, R, 384, V("backbone.final_layer_norm.weight"), V("backbone.final_layer_norm.bias"), nullptr,
                   nullptr, 1e-6f, h, st);
  launch_gemm(h, 384, V("head.out.weight"), 0, V("head.out.bias"), head, 1282, R, 1282, 384, EPI_NONE, nullptr, st);
  launch_vocos_spectrum(head, R, spec, 1296, st);
  launch_gemm(spec, 1296, e->dft_basis, 0, nullptr, frames, 1280, R, 1280, 1296, EPI_NONE, nullptr, st);
  launch_vocos_overlap_add(frames, B, T, V("head.istft.window"), wv, st);
  e->launches += 5;
  if ((r = copy_out(wave, wv, (size_t)R * 320 * 4, mem, st))) return r;
  CK(cudaEventRecord(e->ev1, st));
  CK(cudaMemcpyAsync(e->h_pinned, e->d_flag.p, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  CK(cudaEventElapsedTime(&e->ms_vocos, e->ev0, e->ev1));
  CK(cudaGetLastError());
  if (e->h_pinned[0] != 0) return fail(VX_ERR_ARG, "vx_vocos_decode: code out of range (0..1023)");
  return VX_OK;
}

int vx_last_timing(vx_engine* e, float* ms_prefill, float* ms_decode, float* ms_nar, float* ms_vocos,
                   int64_t* decode_steps, int64_t* kernel_launches) {
  if (!e) return fail(VX_ERR_ARG, "vx_last_timing: null engine");
  if (ms_prefill) *ms_prefill = e->ms_prefill;
  if (ms_decode) *ms_decode = e->ms_decode;
  if (ms_nar) *ms_nar = e->ms_nar;
  if (ms_vocos) *ms_vocos = e->ms_vocos;
  if (decode_steps) *decode_steps = e->decode_steps;
  if (kernel_launches) *kernel_launches = e->launches;
  return VX_OK;
}

const char* vx_last_decode_kernel(vx_engine* e) { return e ? e->decode_kernel.c_str() : ""; }
int vx_ar_ctx_clamped(vx_engine* e) { return (e && e->ctx_clamped) ? 1 : 0; }

int vx_debug_read(vx_engine* e, const char* what, float* dst, int64_t max_elems, int64_t* out_elems) {
  if (!e || !what || !dst) return fail(VX_ERR_ARG, "vx_debug_read: bad arguments");
  CK(cudaSetDevice(e->device));
  const std::string k(what);
  const void* src = nullptr;
  int64_t n = 0;
  if (k == "ar_logits") { src = e->logits.p; n = (int64_t)e->B * kARLogits; }
  else if (k == "nar_logits") { src = e->nlog.p; int64_t tr = 0; for (int t : e->hT) tr += t; n = tr * kVocab; }
  else if (k == "ar_prefill_x") { src = e->x.p; n = (int64_t)e->ar_rows * kD; }
  else if (k == "ada_all") { src = e->ada_all; n = 25 * 7 * 2 * kD; }
  else if (k == "decode_prof") { src = e->p_prof.p; n = 3 * 4096 * 2; }  // int64 stamps viewed as float pairs
  else return fail(VX_ERR_ARG, "vx_debug_read: unknown item " + k);
  if (!src) return fail(VX_ERR_STATE, "vx_debug_read: item not available yet");
  n = std::min(n, max_elems);
  CK(cudaDeviceSynchronize());
  CK(cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost));
  if (out_elems) *out_elems = n;
  return VX_OK;
}

// ---- isolated kernel tests (tests/test_gpu_kernels.py): host fp32 in/out, run on the current device ----
__global__ void unpack_image_kernel(const uint8_t* __restrict__ img, int64_t rows, int K, float* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * K) return;
  const int64_t r = i / K;
  const int k = (int)(i % K);
  out[i] = __bfloat162float(*reinterpret_cast<const bf16*>(img + tc::tile_image_offset(r, k, K)));
}
__global__ void bf16_to_f32_kernel(const bf16* __restrict__ in, float* __restrict__ out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = __bfloat162float(in[i]);
}

int vx_debug_gemm_tc(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K, int epi) {
  if (!A || !W || !C || M < 1 || N % 128 || K % 64) return fail(VX_ERR_ARG, "vx_debug_gemm_tc: bad arguments");
  const int64_t Mp = (M + 127) / 128 * 128;
  float *dA = nullptr, *dW = nullptr, *dB = nullptr, *dC = nullptr;
  void *iA = nullptr, *iW = nullptr, *o16 = nullptr;
  CK(cudaMalloc(&dA, M * K * 4)); CK(cudaMalloc(&dW, (size_t)N * K * 4)); CK(cudaMalloc(&dC, Mp * N * 4));
  CK(cudaMalloc(&iA, Mp * K * 2)); CK(cudaMalloc(&iW, (size_t)N * K * 2)); CK(cudaMalloc(&o16, Mp * N * 2));
  CK(cudaMemcpy(dA, A, M * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dW, W, (size_t)N * K * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dC, C, M * N * 4, cudaMemcpyHostToDevice));  // residual input for TC_EPI_F32_RESIDUAL
  if (bias) { CK(cudaMalloc(&dB, N * 4)); CK(cudaMemcpy(dB, bias, N * 4, cudaMemcpyHostToDevice)); }
  launch_pack_tile_image(dA, M, K, iA, 0);
  launch_pack_tile_image(dW, N, K, iW, 0);
  if (epi == TC_EPI_BF16_ROW) {
    launch_gemm_tc(iA, iW, dB, o16, N, M, N, K, epi, 0);
    bf16_to_f32_kernel<<<(unsigned)((M * N + 255) / 256), 256>>>((const bf16*)o16, dC, (size_t)M * N);
  } else if (epi == TC_EPI_BF16_IMAGE_RELU) {
    launch_gemm_tc(iA, iW, dB, o16, N, M, N, K, epi, 0);
    unpack_image_kernel<<<(unsigned)((M * N + 255) / 256), 256>>>((const uint8_t*)o16, M, N, dC);
  } else {
    launch_gemm_tc(iA, iW, dB, dC, N, M, N, K, epi, 0);
  }
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(C, dC, M * N * 4, cudaMemcpyDeviceToHost));
  cudaFree(dA); cudaFree(dW); cudaFree(dC); cudaFree(iA); cudaFree(iW); cudaFree(o16);
  if (dB) cudaFree(dB);
  return VX_OK;
}

int vx_debug_attn_tc(const float* qkv, float* out, int S, int len, int mode) {
  if (!qkv || !out || S < 1 || len <= S - (mode == 0 ? 1 : 0) || (mode != 0 && mode != 1)) return fail(VX_ERR_ARG, "vx_debug_attn_tc: bad arguments");
  const int64_t Lp = (len + 127) / 128 * 128;
  float *dq = nullptr, *dout = nullptr;
  bf16* q16 = nullptr;
  void* img = nullptr;
  int32_t* meta = nullptr;  // S, P, T, row0
  CK(cudaMalloc(&dq, (size_t)len * 3 * kD * 4)); CK(cudaMalloc(&q16, (size_t)len * 3 * kD * 2));
  CK(cudaMalloc(&img, Lp * kD * 2)); CK(cudaMalloc(&dout, (size_t)len * kD * 4)); CK(cudaMalloc(&meta, 16));
  CK(cudaMemset(img, 0, Lp * kD * 2));
  const int32_t h[4] = {S, mode == 1 ? len - S - 1 : 0, mode == 1 ? 0 : len - S, 0};
  CK(cudaMemcpy(meta, h, 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dq, qkv, (size_t)len * 3 * kD * 4, cudaMemcpyHostToDevice));
  f32_to_bf16_kernel<<<296, 256>>>(dq, q16, (size_t)len * 3 * kD);
  Batch bt{};
  bt.B = 1; bt.S = meta; bt.P = meta + 1; bt.T = meta + 2;
  launch_attn_tc(bt, q16, meta + 3, mode, len, img, 0);
  unpack_image_kernel<<<(unsigned)(((int64_t)len * kD + 255) / 256), 256>>>((const uint8_t*)img, len, kD, dout);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, (size_t)len * kD * 4, cudaMemcpyDeviceToHost));
  cudaFree(dq); cudaFree(q16); cudaFree(img); cudaFree(dout); cudaFree(meta);
  return VX_OK;
}

int vx_debug_attn_t5(const float* qkv, float* out, int S, int len, int mode) {
  if (!qkv || !out || S < 1 || len <= S - (mode == 0 ? 1 : 0) || (mode != 0 && mode != 1)) return fail(VX_ERR_ARG, "vx_debug_attn_t5: bad arguments");
  const int64_t Lp = (len + 127) / 128 * 128;
  float *dq = nullptr, *dout = nullptr;
  bf16* q16 = nullptr;
  void *img = nullptr, *qi = nullptr, *ki = nullptr, *vi = nullptr;
  int32_t* meta = nullptr;  // S, P, T, row0
  CK(cudaMalloc(&dq, (size_t)Lp * 3 * kD * 4)); CK(cudaMalloc(&q16, (size_t)Lp * 3 * kD * 2));
  CK(cudaMalloc(&img, Lp * kD * 2)); CK(cudaMalloc(&dout, (size_t)len * kD * 4)); CK(cudaMalloc(&meta, 16));
  CK(cudaMalloc(&qi, Lp * kD * 2)); CK(cudaMalloc(&ki, Lp * kD * 2)); CK(cudaMalloc(&vi, Lp * kD * 2));
  CK(cudaMemset(img, 0, Lp * kD * 2));
  CK(cudaMemset(dq, 0, (size_t)Lp * 3 * kD * 4));
  const int32_t h[4] = {S, mode == 1 ? len - S - 1 : 0, mode == 1 ? 0 : len - S, 0};
  CK(cudaMemcpy(meta, h, 16, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dq, qkv, (size_t)len * 3 * kD * 4, cudaMemcpyHostToDevice));
  f32_to_bf16_kernel<<<296, 256>>>(dq, q16, (size_t)Lp * 3 * kD);
  Batch bt{};
  bt.B = 1; bt.S = meta; bt.P = meta + 1; bt.T = meta + 2;
  launch_pack_qkv_heads(q16, Lp, qi, ki, vi, 0);
  launch_attn_t5(bt, qi, ki, vi, Lp, meta + 3, mode, len, img, 0);
  unpack_image_kernel<<<(unsigned)(((int64_t)len * kD + 255) / 256), 256>>>((const uint8_t*)img, len, kD, dout);
  CK(cudaDeviceSynchronize());
  CK(cudaGetLastError());
  CK(cudaMemcpy(out, dout, (size_t)len * kD * 4, cudaMemcpyDeviceToHost));
  cudaFree(dq); cudaFree(q16); cudaFree(img); cudaFree(dout); cudaFree(meta); cudaFree(qi); cudaFree(ki); cudaFree(vi);
  return VX_OK;
}

}  // extern "C"
