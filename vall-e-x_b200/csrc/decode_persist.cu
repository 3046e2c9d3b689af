// Persistent AR decode kernel: the whole token loop of models/vallex.py:528-598 in ONE cooperative launch.
//
// One CTA per SM (512 threads), all resident; every decode step walks
//     S0  sample + append + embed (CTA b serves utterance b)              models/vallex.py:529-531,569-598
//     12 x [ P1 LN1 + QKV rows | P2 attention (KV append, split-K over CTAs, ticket combine)
//            | P3 out-proj + residual | P4 LN2 + FFN1 + ReLU | P5 FFN2 (split-K x4, partial sums) ]
//     PF  final LN + predict rows -> logits                               modules/transformer.py:337-347,470-471
// separated by grid-wide barriers (release/acquire counter in L2), 62 per step.  Nothing returns to the host
// between tokens: sampling, EOS/budget bookkeeping and the all-finished early exit live on the device.
//
// Data path (HBM bound): each CTA owns a fixed, contiguous slice of the rows of every weight matrix
// (row-major, so a slice is one contiguous 40-60 KB range); a warp streams one row at a time with 128-bit
// coalesced loads (4-8 in flight per lane), activations (fp32, B x 1024) sit in shared memory, the reduction
// is a warp shuffle.  While a phase runs, the CTA's slice for the NEXT phase is pulled into L2 with
// cp.async.bulk.prefetch.L2 so HBM keeps streaming across the barriers.  The KV cache ([b][h][pos][64], 128 B
// rows in bf16) is read once per step with 128-bit loads; this step's key/value are appended in place.
// FFN2 is split over K so every GEMV phase has the same shape (~21-28 rows x 1024); the four partial sums are
// folded (fixed order, deterministic) by the consumer's prologue together with bias and residual.
//
// Algorithmic bytes per step: 304,412,672 + 49,152 * sum_b(L_b + 1) at bf16.
#include "decode_common.cuh"

namespace vx {
namespace {
// ---- the kernel -------------------------------------------------------------------------------------------
template <typename WT, int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_persistent_kernel(const PParams p) {
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* scratch = dsm + MR * kD;        // attention / sampling / LN-reduction scratch
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;
  const size_t wsz = sizeof(WT);

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // FFN2 split-K: CTA -> (k chunk kc, row slice)
  const int per = gridDim.x / 4;                   // CTAs per k chunk (the last chunk takes the remainder)
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  WPre<WT, MR> pre;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  // everything that does not depend on a barrier is requested BEFORE it: weight rows, their biases, LN parameters
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  // optional timeline: CTAs 0, grid/2 and grid-1 stamp every barrier of steps [prof_step0, prof_step0 + 4)
  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed -------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b) && !(step == 0 && p.resume && !ld_cg(p.pending + b))) sample_utterance(p, b, scratch);   // resumed launch: a continuing utterance sampled its token before the pause
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (p.stop_below > 0 && ld_cg(bt.n_active) <= p.stop_below) break;   // continuous batching: hand a finished slot back to the host
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      // ---- P1: LN1 + QKV rows; pull this CTA's K/V ranges of the coming attention phase into L2 ----------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        if (tid == 0) {
          const int split = u % p.nsplit, bh = u / p.nsplit, b = bh / kH, h = bh % kH;
          const int n = bt.S[b] + bt.P[b] + ld_cg(bt.T + b) + 1;
          int chunk = (n + p.nsplit - 1) / p.nsplit;
          chunk = (chunk + 31) & ~31;
          const int j0 = split * chunk, j1 = min(n - 1, j0 + chunk);
          if (j1 > j0) {
            const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * wsz;
            l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * wsz);
            l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * wsz);
          }
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, lp, scratch);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, Wo, kD, o0, o1, L.out_b, fb_prev);   // out-proj rows ride through the attention phase
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention ---------------------------------------------------------------------------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        if (sizeof(WT) == 2) attn_unit<bf16>(p, l, b, h, split, scratch);
        else attn_unit<float>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: (combine key splits) + out-proj + residual -> xa ------------------------------------------------
      stage_plain(sx, B, [&](int m, int c) {
        if (ld_cg(bt.finished + m)) return 0.f;
        return (p.dbg & 2) ? att_combined(p, m, c) : __ldcg(p.att + (size_t)m * kD + c);
      });
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wo, kD, o0, o1, sx, B, L.out_b, fb_prev, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                          [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      ln_issue(lp, L.n2_g, L.n2_b, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU -------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, nullptr);
      stage_ln<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W2, kFF, g0, g1, nullptr, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2, split-K: rows [g0,g1) x k chunk kc -> partial pf[kc] --------------------------------------
      stage_plain(sx, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W2, kFF, g0, g1, sx, B, nullptr, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
      PSTAMP(4);
      if (l + 1 < kL) {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1, p.layer[l + 1].qkv_b, nullptr);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: final LayerNorm + predict rows -> logits ---------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, lp, scratch);
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
  (void)wsz;
}


template <typename WT, int MR>
static cudaError_t persist_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;             // attention
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;     // sampling
  const size_t smem = ((size_t)MR * kD + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_persistent_kernel<WT, MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_persistent_kernel<WT, MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_persistent_kernel<WT, MR>, dim3(grid), dim3(PNT), args, smem, st);
}


}  // namespace

// wt: 0 fp32 / 1 bf16.  B <= 32.  Returns a cudaError_t as int.
int launch_decode_persistent(const void* params, int wt, int B, int grid, cudaStream_t st) {
  const PParams& p = *reinterpret_cast<const PParams*>(params);
  const int mr = B <= 1 ? 1 : B <= 2 ? 2 : B <= 4 ? 4 : B <= 8 ? 8 : B <= 16 ? 16 : 32;
  if (wt == 1 && p.att_maxp > 0) {   // tensor-core batched kernel (engine.cu decides: bf16, 5 <= B <= 32)
    return (int)launch_decode_mma(p, mr < 8 ? 8 : mr, grid, st);
  }
  if (wt == 1 && p.cluster) return (int)launch_decode_cluster(p, grid, st, nullptr);   // batch 1, head groups as clusters
  if (wt == 1 && p.fused) {  // 3-barrier variant: requires one attention unit per CTA and <= kDecodeHS hidden units per CTA
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kDecodeHS) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_fused(p, mr, grid, st);
  }
#define VX_P(WT_, MR_) return (int)persist_launch<WT_, MR_>(p, grid, st)
  if (wt == 0) {
    switch (mr) { case 1: VX_P(float, 1); case 2: VX_P(float, 2); case 4: VX_P(float, 4); case 8: VX_P(float, 8); case 16: VX_P(float, 16); default: VX_P(float, 32); }
  } else {
    switch (mr) { case 1: VX_P(bf16, 1); case 2: VX_P(bf16, 2); case 4: VX_P(bf16, 4); case 8: VX_P(bf16, 8); case 16: VX_P(bf16, 16); default: VX_P(bf16, 32); }
  }
#undef VX_P
}


}  // namespace vx
