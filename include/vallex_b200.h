/*
 * vallex_b200.h -- C ABI of the B200-native VALL-E-X inference engine (libvallex_b200.so).
 *
 * The reference (Plachtaa/VALL-E-X) has no FFI of its own: its boundary is the Python API
 * `utils.generation.generate_audio` / `models.vallex.VALLE.inference` plus two file formats
 * (checkpoint state_dict keys, presets/*.npz).  The Python host layer in
 * vall-e-x_b200/vallex_b200/ mirrors that API and binds the entry points below with ctypes
 * (see INTEGRATION.md for the stub a reference maintainer would add).  Each entry point names
 * the reference code it replaces (paths relative to the reference tree).
 *
 * Conventions: plain pointers and sizes only; every call returns 0 on success and a negative
 * code on failure with a thread-local message in vx_last_error(); `mem` says where the caller's
 * data buffers live (VX_HOST: pageable or pinned host memory, copies are enqueued on the stream
 * inside the call; VX_DEVICE: device pointers on the engine's device); `stream` is a
 * cudaStream_t (NULL = the engine's own stream).  An engine handle is not thread-safe; distinct
 * handles are independent.  One engine per GPU; there is no CPU fallback.
 */
#ifndef VALLEX_B200_H_
#define VALLEX_B200_H_

#include <stdint.h>

#if defined(__GNUC__)
#define VX_API __attribute__((visibility("default")))
#else
#define VX_API
#endif

#ifdef __cplusplus
extern "C" {
#endif

#define VX_OK 0
#define VX_ERR_ARG (-1)
#define VX_ERR_CUDA (-2)
#define VX_ERR_STATE (-3)
#define VX_ERR_CAPACITY (-4)

/* dtype_mode */
#define VX_F32 0  /* fp32 weights, fp32 KV, fp32 math: the parity mode (bit-exact greedy tokens vs the fp32 reference) */
#define VX_BF16 1 /* bf16 GEMM weights + bf16 KV cache, fp32 accumulate: the performance mode */

/* mem */
#define VX_HOST 0
#define VX_DEVICE 1

typedef struct vx_engine vx_engine;

/* ABI/version string, e.g. "vallex_b200 0.1 sm_100a". */
VX_API const char* vx_version(void);
VX_API const char* vx_last_error(void);

/* Replaces: model construction `VALLE(N_DIM, NUM_HEAD, NUM_LAYERS, ...)` + `.to(device)`
 * (utils/generation.py:67-78).  Architecture constants are fixed (macros.py:1-6).
 * max_batch: utterances in flight; max_ctx: max S+P+1+T rows of one utterance. */
VX_API int vx_create(int device, int max_batch, int max_ctx, int dtype_mode, vx_engine** out);
VX_API void vx_destroy(vx_engine* e);

/* Replaces: `model.load_state_dict(checkpoint["model"], strict=True)` (utils/generation.py:79-83) and
 * `Vocos.from_pretrained(...)` weights (utils/generation.py:89).  `key` is the reference state-dict key
 * (SURVEY.md 8(b)); Vocos keys are prefixed "vocos." (e.g. "vocos.backbone.embed.weight"); the fp32
 * sinusoid table is passed as key "pe_table" [rows,1024] (modules/embedding.py:68-91 is evaluated with
 * torch on the host so the table is bit-identical).  `data` is host fp32, row-major, `shape[ndim]`.
 * vx_finalize_weights: strict check (every key present, shapes match), packs to the kernel layouts and
 * precomputes the 7x25 AdaLN (weight,bias) vectors (modules/transformer.py:93-108). */
VX_API int vx_load_weight(vx_engine* e, const char* key, const float* data, const int64_t* shape, int ndim);
/* Same with `mem` saying where `data` lives: VX_DEVICE takes a device pointer on the engine's GPU (e.g. the NCCL receive
 * buffer of the multi-GPU start-up broadcast, vallex_b200/dist.py) -- the tensor is copied and packed on the device and
 * never visits the host. */
VX_API int vx_load_weight_mem(vx_engine* e, const char* key, const float* data, const int64_t* shape, int ndim, int mem);
VX_API int vx_finalize_weights(vx_engine* e);

/* Replaces: the argument marshalling of `VALLE.inference` (models/vallex.py:458-527): B independent
 * utterances (the reference asserts B == 1, models/vallex.py:491; best_of replicas are B copies).
 * text_ids [B,Smax] phoneme ids (prompt phonemes then new phonemes), text_lens [B],
 * lang_ids [B,Smax] per-token language id in the MODEL numbering en=0 zh=1 ja=2
 * (models/vallex.py:439-443,499-505: prompt part = prompt_language, rest = text_language),
 * prompt_codes [B,Pmax,8] EnCodec tokens, prompt_lens [B] (0 allowed: no prompt). */
VX_API int vx_set_batch(vx_engine* e, int B, const int32_t* text_ids, const int32_t* text_lens, const int32_t* lang_ids,
                 int Smax, const int32_t* prompt_codes, const int32_t* prompt_lens, int Pmax, int mem, void* stream);

/* Replaces: the first iteration of the AR loop (kv_cache is None; models/vallex.py:528-568 with
 * modules/transformer.py:447-473, modules/activation.py:142-167): embeds text+language+PE and
 * BOS+prompt codebook-0, runs the 12 pre-LN layers under the text/audio mask (models/vallex.py:534-549),
 * fills the KV cache and leaves the first-step logits on the device. */
VX_API int vx_ar_prefill(vx_engine* e, void* stream);

/* Replaces: the rest of the AR loop (models/vallex.py:528-598) incl. `topk_sampling`
 * (models/vallex.py:791-853) and the stop rule: per utterance, stop at EOS (1024) or after
 * min(max_new, 16*S_b) tokens (max_new <= 0: 16*S_b only; models/vallex.py:574-578).
 * top_k <= 0: no top-k filter (the API default -100); top_p >= 1: off; seed: Philox stream for the
 * multinomial draw.  ignore_eos != 0: never stop early (benchmark config 3).  forced_tokens
 * (nullable) [B,max_steps]: teacher forcing -- the token fed back at step t is forced_tokens[b][t]
 * while out_tokens still records the sampled one.
 * out_tokens [B,max_steps] (max_steps = the largest per-utterance budget, returned in *out_max_steps
 * by vx_ar_budget), out_lens [B] = T_b, out_sum_logprobs [B] nullable (models/vallex.py:572). */
VX_API int vx_ar_budget(vx_engine* e, int max_new, int* out_max_steps);

/* Continuous batching (engine extension; the reference decodes the sentences of a long text one after another,
 * utils/generation.py:199-224).  vx_ar_decode_ex is vx_ar_decode with three more controls: `stop_on_finish` != 0 returns as soon
 * as ONE more utterance has finished (EOS / budget) instead of waiting for the slowest one, `out_finished` [B] reports which
 * slots are done, and `resume` != 0 continues the paused batch: slots that kept decoding carry on from the token they sampled
 * last, slots refilled by vx_slot_prefill sample their first token.  out_tokens / out_lens always describe the slot's
 * CURRENT utterance (tokens [0, T_b)).  vx_slot_prefill puts a new utterance (its own text / languages / prompt; S <= Smax and
 * P <= Pmax of the vx_set_batch that sized the batch) into a finished slot: it is prefilled alone into the slot's KV region. */
VX_API int vx_ar_decode_ex(vx_engine* e, int max_new, int top_k, float top_p, float temperature, uint64_t seed, int ignore_eos,
                    const int32_t* forced_tokens, int32_t* out_tokens, int32_t* out_lens, float* out_sum_logprobs, int resume,
                    int stop_on_finish, int32_t* out_finished, int mem, void* stream);
VX_API int vx_slot_prefill(vx_engine* e, int slot, const int32_t* text_ids, int S, const int32_t* lang_ids,
                    const int32_t* prompt_codes, int P, int mem, void* stream);

/* Replaces: `x.repeat(best_of, 1, 1); y.repeat(best_of, 1)` (models/vallex.py:525-527), where the reference pushes best_of
 * copies of the utterance through the first (full-sequence) iteration.  Here the utterance is prefilled ONCE
 * (vx_set_batch with B = 1, vx_ar_prefill) and vx_ar_fork replicates its decode state -- KV rows of every layer, token buffer,
 * first-step logits -- on the device for beams 1..n_beams-1 (n_beams <= max_batch); vx_ar_decode then runs the beams as a
 * batch, each with its own Philox stream.  vx_ar_select_beam replaces models/vallex.py:583-594: avg = sum_logprobs /
 * len ** length_penalty (len = BOS + prompt + generated non-EOS tokens), first argmax (return_worst: first argmin), computed on the
 * device; the chosen beam becomes the batch's only utterance so vx_nar(cb0 = NULL) continues with it.  *out_pick = beam index. */
VX_API int vx_ar_fork(vx_engine* e, int n_beams, void* stream);
VX_API int vx_ar_select_beam(vx_engine* e, float length_penalty, int return_worst, int* out_pick, void* stream);
VX_API int vx_ar_decode(vx_engine* e, int max_new, int top_k, float top_p, float temperature, uint64_t seed,
                 int ignore_eos, const int32_t* forced_tokens, int32_t* out_tokens, int32_t* out_lens,
                 float* out_sum_logprobs, int mem, void* stream);

/* Replaces: the NAR stage loop (models/vallex.py:604-686, prefix_mode 1) with
 * modules/transformer.py:265-312,402-445 (AdaLN encoder) and modules/activation.py:566-585.
 * cb0 [B,Tmax] generated codebook-0 tokens (NULL: use the tokens vx_ar_decode left on the device),
 * t_lens [B].  out_codes [B,Tmax,8] (column 0 = cb0). */
VX_API int vx_nar(vx_engine* e, const int32_t* cb0, const int32_t* t_lens, int Tmax, int32_t* out_codes, int mem,
           void* stream);

/* Replaces: `vocos.codes_to_features(frames)` + `vocos.decode(features, bandwidth_id)`
 * (utils/generation.py:148-150).  codes [8,B,T]; wave [B, 320*T] fp32. */
VX_API int vx_vocos_decode(vx_engine* e, const int32_t* codes, int B, int T, int bandwidth_id, float* wave, int mem,
                    void* stream);

/* ---- Prompt making: EnCodec 24 kHz encoder + residual vector quantiser ---------------------------------
 * Replaces: `AudioTokenizer(device)` = `EncodecModel.encodec_model_24khz()` + `set_target_bandwidth(6.0)` +
 * `remove_encodec_weight_norm` (data/tokenizer.py:63-88) and `AudioTokenizer.encode` / `tokenize_audio`
 * (data/tokenizer.py:98-111) as called by `make_prompt` (utils/prompt_making.py:57-84).  Its own handle, like the
 * reference's separate `codec` global (utils/generation.py:41-45,86-87).  sm_100 devices only; no CPU fallback. */
typedef struct vx_encodec vx_encodec;
VX_API int vx_encodec_create(int device, vx_encodec** out);
VX_API void vx_encodec_destroy(vx_encodec* e);
/* `key` is a key of the `encodec` package's `EncodecModel.state_dict()`: `encoder.model.<i>...conv.conv.{weight,bias}` (or
 * `weight_g` + `weight_v` when weight norm was not removed), `encoder.model.13.lstm.{weight,bias}_{ih,hh}_l{0,1}`,
 * `quantizer.vq.layers.<q>._codebook.embed` [1024,128] for q = 0.. (at least the 8 tables of 6 kbps).  Host fp32, row-major.
 * Decoder keys are not on this path and are rejected by the Python mirror, not here.  vx_encodec_finalize is strict: every
 * encoder tensor present with the right shape. */
VX_API int vx_encodec_load_weight(vx_encodec* e, const char* key, const float* data, const int64_t* shape, int ndim);
VX_API int vx_encodec_finalize(vx_encodec* e);
VX_API int vx_encodec_num_codebooks(vx_encodec* e);
/* T for an n-sample waveform: ceil at each of the four strided convolutions (2, 4, 5, 8) = the length of `encode`'s codes. */
VX_API int64_t vx_encodec_frames(int64_t n_samples);
/* Replaces: `codec.encode(wav)` for ONE mono 24 kHz waveform (the reference passes [1,1,N], data/tokenizer.py:104-110;
 * `segment=None` and `normalize=False` for this model, so there is one frame and no scale).  wav [n_samples] fp32;
 * n_q codebooks (8 at the reference's 6 kbps); out_codes [n_q, T] int32, T = vx_encodec_frames(n_samples).
 * Non-finite samples -> VX_ERR_ARG. */
VX_API int vx_encodec_encode(vx_encodec* e, const float* wav, int64_t n_samples, int n_q, int32_t* out_codes, int mem, void* stream);
/* parity hooks: the encoder output [128, T] of the last vx_encodec_encode (host fp32); its CUDA-event time, frames, launches */
VX_API int vx_encodec_embeddings(vx_encodec* e, float* dst, int64_t max_elems, int64_t* out_elems);
VX_API int vx_encodec_last_timing(vx_encodec* e, float* ms, int64_t* frames, int64_t* kernel_launches);

/* Measurement hooks (bench.py): CUDA-event durations (ms) of the last prefill / decode / nar / vocos
 * call on the engine's stream, the number of decode steps and kernels launched by the last call, and
 * debug access to device-side intermediates for the layer-by-layer parity tests. */
VX_API int vx_last_timing(vx_engine* e, float* ms_prefill, float* ms_decode, float* ms_nar, float* ms_vocos,
                   int64_t* decode_steps, int64_t* kernel_launches);
VX_API int vx_debug_read(vx_engine* e, const char* what, float* dst, int64_t max_elems, int64_t* out_elems);
/* Name of the kernel the last vx_ar_decode actually launched (e.g. "ar_decode_cluster_kernel" = batch 1, "ar_decode_fused_kernel<2>",
 * "ar_decode_mma_kernel<32>"):
 * the engine picks the decode kernel from dtype mode, batch size and context, so bench.py reports this string instead of
 * guessing from the environment.  Valid until the next vx_ar_decode / vx_destroy. */
VX_API const char* vx_last_decode_kernel(vx_engine* e);
/* 1 if the last vx_ar_decode had to cut some utterance's budget below min(max_new, 16*S) because S+P+1+budget would not fit
 * max_ctx (the reference has no such limit: it would run to its 16*S cap, models/vallex.py:577).  The Python mirror raises
 * instead of returning a silently truncated result. */
VX_API int vx_ar_ctx_clamped(vx_engine* e);

/* Isolated kernel checks (tests/test_gpu_kernels.py); host fp32 buffers, current CUDA device.
 * vx_debug_gemm_tc: C[M,N] = A[M,K] W[N,K]^T (+bias) through the tcgen05 GEMM (epi 0: bf16 row-major, 1: fp32 += into
 * the given C, 2: ReLU -> bf16 tile image, 3: fp32).  vx_debug_attn_tc: one utterance, qkv [len,3072] -> out [len,1024]
 * through the bf16 tensor-core attention (mode 0 bidirectional, 1 AR prefill mask with S text rows). */
VX_API int vx_debug_gemm_tc(const float* A, const float* W, const float* bias, float* C, int64_t M, int N, int K, int epi);
VX_API int vx_debug_attn_tc(const float* qkv, float* out, int S, int len, int mode);
/* same contract through the tcgen05/TMEM attention kernel (attention_t5.cu) */
VX_API int vx_debug_attn_t5(const float* qkv, float* out, int S, int len, int mode);

#ifdef __cplusplus
}
#endif
#endif /* VALLEX_B200_H_ */
