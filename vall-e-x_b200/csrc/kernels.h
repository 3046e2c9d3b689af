// Host-side launchers of the vallex_b200 kernels.  wt = VX_F32 (0) / VX_BF16 (1) selects the storage
// type of GEMM weights and of the KV cache; activations on the residual stream are always fp32.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace vx {

// Device-resident description of the utterances in flight (all arrays have B entries).
struct Batch {
  int B;
  int Smax, Pmax;            // strides of text_ids/lang_ids and prompt_codes
  const int32_t* text_ids;   // [B,Smax]
  const int32_t* lang_ids;   // [B,Smax]
  const int32_t* prompt;     // [B,Pmax,8]
  const int32_t* S;          // [B] phoneme count
  const int32_t* P;          // [B] prompt frames
  const int32_t* ar_row0;    // [B] first packed row of utterance b in the AR prefill matrix
  const int32_t* nar_row0;   // [B] first packed row in the NAR matrix
  int32_t* T;                // [B] generated frames (written by the AR decode, read by the NAR)
  int32_t* ctx;              // [B] keys currently in the KV cache
  int32_t* y;                // [B,ymax] BOS + prompt codebook 0 + generated tokens
  int ymax;
  int32_t* budget;           // [B] max new tokens for this utterance
  int32_t* finished;         // [B]
  float* sum_logprobs;       // [B]
  int32_t* n_active;         // [1] utterances still decoding (polled by the host)
};

// The 8 NAR codebook embedding tables; tables 2..7 are tied to the predict layers and are bf16 in bf16 mode.
struct EmbTables {
  const void* p[8];
  int bf[8];
};

enum Epilogue { EPI_NONE = 0, EPI_RELU = 1, EPI_RESIDUAL = 2, EPI_GELU = 3, EPI_SCALE_RESIDUAL = 4 };

// ---- embeddings / positions (models/vallex.py:497-507,529-531,605-632; modules/embedding.py:93-97) ----
void launch_embed_text(const Batch& bt, const float* text_emb, const float* lang_emb, const float* alpha,
                       const float* pe, const int32_t* row0, float* x, cudaStream_t st);
void launch_embed_audio_prefill(const Batch& bt, const float* audio_emb, const float* alpha, const float* pe,
                                float* x, cudaStream_t st);
void launch_embed_audio_step(const Batch& bt, const float* audio_emb, const float* alpha, const float* pe,
                             float* xstep, cudaStream_t st);
void launch_init_y(const Batch& bt, cudaStream_t st);
void launch_nar_embed_init(const Batch& bt, const EmbTables& tab, float* yemb, const int32_t* yrow0, int max_pt,
                           cudaStream_t st);
void launch_nar_build_x(const Batch& bt, const float* yemb, const int32_t* yrow0, const float* alpha,
                        const float* pe, float* x, int max_pt, cudaStream_t st);
void launch_nar_argmax_update(const Batch& bt, const float* logits, const int32_t* trow0, int stage,
                              const EmbTables& tab, float* yemb, const int32_t* yrow0, int32_t* codes, int Tmax,
                              cudaStream_t st);
void launch_nar_copy_cb0(const Batch& bt, int32_t* codes, int Tmax, cudaStream_t st);

// ---- LayerNorm / AdaLN (modules/transformer.py:57-74,93-108) ----
// out[r] = [ada_w *] (LN(x[row_map ? row_map[r] : r]) * gamma + beta) [+ ada_b]; fp32 in/out.
void launch_layernorm(const float* x, const int32_t* row_map, int64_t rows, int dim, const float* gamma,
                      const float* beta, const float* ada_w, const float* ada_b, float eps, float* out,
                      cudaStream_t st);

// ---- dense contractions: C[M,N] = A[M,K] * W[N,K]^T (+bias) with epilogue ----
// W is fp32 or bf16 (wt); A, C fp32.  EPI_RESIDUAL: C += result (C holds the residual stream);
// EPI_SCALE_RESIDUAL: C += scale[n] * result (Vocos gamma).
void launch_gemm(const float* A, int64_t lda, const void* W, int wt, const float* bias, float* C, int64_t ldc,
                 int64_t M, int N, int K, int epi, const float* scale, cudaStream_t st);
// Small-M (decode) path; optionally applies LayerNorm(gamma,beta,eps=1e-5) to each A row first.
void launch_gemv(const float* A, int64_t lda, const void* W, int wt, const float* bias, float* C, int64_t ldc,
                 int M, int N, int K, int epi, const float* ln_gamma, const float* ln_beta, cudaStream_t st);

// ---- attention ----
// mode 0: bidirectional (NAR); mode 1: AR prefill mask (models/vallex.py:534-549).
// qkv packed [rows,3072] fp32; out [rows,1024] fp32.
void launch_attn_full(const Batch& bt, const float* qkv, const int32_t* row0, int mode, int max_len, float* out,
                      cudaStream_t st);
// Scatter K/V of the prefill rows into the cache [b][h][pos][64] (type kvt).
void launch_kv_scatter(const Batch& bt, const float* qkv, void* kcache, void* vcache, int kvt, int max_len,
                       int ctx_cap, cudaStream_t st);
// Decode: append this step's k,v (qkv_step [B,3072]) at position ctx[b], attend over ctx[b]+1 keys.
void launch_attn_decode(const Batch& bt, const float* qkv_step, void* kcache, void* vcache, int kvt, int ctx_cap,
                        float* out, float* partial, int32_t* tickets, int nsplit, cudaStream_t st);

// ---- bf16 tensor-core path (tcgen05 GEMM, mma.sync attention); operands are "tile images" (tc_common.cuh) ----
enum TcEpilogue { TC_EPI_BF16_ROW = 0, TC_EPI_F32_RESIDUAL = 1, TC_EPI_BF16_IMAGE_RELU = 2, TC_EPI_F32 = 3,
                  TC_EPI_QKV_HEADS = 4, TC_EPI_F32_SCALE_RESIDUAL = 5, TC_EPI_GELU_SPLIT3 = 6 };
                  // 5: out[r][n] += aux[n] * (acc + bias[n]) (Vocos gamma residual); 6: GELU(acc + bias) -> split-3 image [rows, 3N]   // 4: N = 3072 QKV projection written as the per-head operand images of attention_t5.cu
                                            //    (out = [q_img | k_img | vt_img], ldc = padded row count R)
void launch_gemm_tc(const void* A_img, const void* W_img, const float* bias, void* out, int64_t ldc, int64_t M, int N,
                    int K, int epi, cudaStream_t st, const float* aux = nullptr);
// fp32 weight [rows, K_src] (row stride ld) -> split-3 image [pad128(rows), 3K] = [hi | lo | hi]; K % 64 == 0, K >= K_src
void launch_pack_split3_weight(const float* src, int64_t rows, int K_src, int64_t ld, int K, void* img, cudaStream_t st);
void launch_pack_tile_image(const float* src, int64_t rows, int K, void* img, cudaStream_t st);
void launch_layernorm_image(const float* x, const int32_t* row_map, int64_t rows, const float* gamma,
                            const float* beta, const float* ada_w, const float* ada_b, float eps, void* img,
                            cudaStream_t st);
void launch_attn_tc(const Batch& bt, const void* qkv_bf16, const int32_t* row0, int mode, int max_len, void* out_img,
                    cudaStream_t st);
void launch_kv_scatter_bf16(const Batch& bt, const void* qkv_bf16, void* kcache, void* vcache, int max_len, int ctx_cap,
                            cudaStream_t st);

// tcgen05/TMEM attention (attention_t5.cu): per-head operand images packed from the row-major bf16 QKV
void launch_pack_qkv_heads(const void* qkv_bf16, int64_t R, void* q_img, void* k_img, void* vt_img, cudaStream_t st);
void launch_attn_t5(const Batch& bt, const void* q_img, const void* k_img, const void* vt_img, int64_t R,
                    const int32_t* row0, int mode, int max_len, void* out_img, cudaStream_t st);

// ---- sampling (models/vallex.py:791-853, 569-578) ----
void launch_sample(const Batch& bt, const float* logits, int top_k, float top_p, float temperature, uint64_t seed,
                   int step, int ignore_eos, const int32_t* forced, int forced_stride, int32_t* out_tokens,
                   int out_stride, cudaStream_t st);

// ---- Vocos (third-party `vocos` package; see oracle/vocos_oracle.py) ----
void launch_vocos_features(const int32_t* codes, int B, int T, const float* table, float* feats, cudaStream_t st);
void launch_vocos_im2col7(const float* in, int B, int T, int C, float* out, cudaStream_t st);
void launch_vocos_dwconv7(const float* x, int B, int T, int C, const float* w, const float* bias, float* out,
                          cudaStream_t st);
void launch_vocos_spectrum(const float* head_out, int64_t rows, float* spec, int ldspec, cudaStream_t st);
// tensor-core Vocos path (bf16 x 3 split operands; activation images are [rows_pad, 3K] = [hi | hi | lo]):
void launch_vocos_im2col_split3(const float* feats, int B, int T, void* img, cudaStream_t st);                 // K = 896
void launch_vocos_dwconv_ln_split3(const float* x, int B, int T, const float* dw_wT, const float* dw_b, const float* scale,
                                   const float* shift, void* img, cudaStream_t st);                              // K = 384
void launch_vocos_ln_split3(const float* x, int64_t rows, const float* gamma, const float* beta, void* img, cudaStream_t st);  // K = 384
void launch_vocos_spectrum_split3(const float* head_out, int ldh, int64_t rows, void* img, cudaStream_t st);     // K = 1344
void launch_vocos_overlap_add(const float* frames, int B, int T, const float* window, float* wave, cudaStream_t st);

}  // namespace vx
