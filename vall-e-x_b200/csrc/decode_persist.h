// Parameter block of the persistent AR decode kernel (decode_persist.cu), filled by engine.cu.
#pragma once
#include <stdint.h>

#include "common.cuh"
#include "kernels.h"

namespace vx {

struct PLayer {
  const void *qkv_w, *out_w, *ff1_w, *ff2_w;
  const void* ff2_t;          // bf16 [4096][1024]: FFN2 weight transposed (fused variant only, else null)
  const float *qkv_b, *out_b, *ff1_b, *ff2_b, *n1_g, *n1_b, *n2_g, *n2_b;
};

struct PParams {
  Batch bt;
  PLayer layer[kL];
  const float *norm_g, *norm_b;
  const void* predict_w;
  const float *audio_emb, *alpha, *pe;
  void *kc, *vc;              // [layer][b][h][ctx_cap][64]
  size_t kv_layer_elems;
  int ctx_cap;
  // scratch (global, L2 resident)
  float* xa;                  // [B][1024] residual stream after the attention block
  float* pf;                  // [4][B][1024] FFN2 split-K partials of the previous layer
  float* qkv;                 // [B][3072]
  float* att;                 // [B][1024]
  float* hid;                 // [B][4096]
  float* logits;              // [B][1025]
  float* xb;                  // [B][1024] fused variant: residual stream after the attention block (xa holds the layer input)
  long long* acc_out;         // [2][B][1024] fused variant: out-proj sums, 2^-30 fixed point, by layer parity
  long long* acc_ffn;         // [2][B][1024] fused variant: FFN2 sums
  int fused;                  // 0: 5-barrier kernel, 1: 3-barrier kernel (default for B <= 2)
  int cluster;                // 1: batch-1 kernel with head groups as 4-CTA clusters (decode_cluster.cu; uses the fused variant's buffers)
  float* part;                // [B][16][nsplit][66]  (balanced batched attention: [B][16][att_maxp][68])
  int kv_l2_mb;               // batched kernel: MB of the next layer's KV rows prefetched into L2 during the GEMV phases (0: off)
  int att_maxp;               // balanced batched attention (decode_mma.cu): max pieces per (utterance, head); 0 = CTA-per-head units
  int32_t* tickets;           // [B*16]
  unsigned int* bar;          // grid barrier counter (zeroed by the host)
  int nsplit;
  // sampling
  int top_k; float top_p; float temperature; uint64_t seed; int ignore_eos;
  const int32_t* forced; int forced_stride; int32_t* out_tokens; int out_stride;
  int max_steps;
  int resume;                 // 1: this launch continues a paused batch (vx_ar_decode_resume): at step 0 only slots with pending[b] != 0 sample
  const int32_t* pending;     // [B] 1: the slot's logits come from a fresh prefill and have not been sampled yet
  int stop_below;             // > 0: leave the loop as soon as n_active <= stop_below (a slot finished; the host refills it)
  int32_t* steps_done;        // [1]
  long long* prof;            // nullable: [3 CTAs][4096] (tag, clock64) pairs of 4 steps (VX_DECODE_PROF=<first step>)
  int prof_step0;
  int dbg;                    // VX_DECODE_DBG bit mask: bisecting aid (1 no attention, 2 fold the key splits in the out-proj phase instead of by ticket, 4 no GEMV, 8 no LN staging, 16 no sampling)
};


// wt: 0 fp32 / 1 bf16 (weights and KV cache).  B <= 32.  Returns a cudaError_t as int.
int launch_decode_persistent(const void* params, int wt, int B, int grid, cudaStream_t st);

// launchers of the kernel variants (one translation unit each); mr = padded batch rows
cudaError_t launch_decode_fused(const PParams& p, int mr, int grid, cudaStream_t st);     // decode_fused.cu
cudaError_t launch_decode_mma(const PParams& p, int mr, int grid, cudaStream_t st);       // decode_mma.cu
cudaError_t launch_decode_cluster(const PParams& p, int num_sms, cudaStream_t st, int* grid_out);   // decode_cluster.cu (B = 1)
int decode_cluster_grid(int num_sms);   // CTAs the cluster kernel would run with on the current device; 0: not launchable
constexpr int kDecodeHS = 64;   // max hidden units per CTA of the fused kernels (kHS in decode_fused_common.cuh)

}  // namespace vx
