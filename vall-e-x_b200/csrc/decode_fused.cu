// Fused persistent decode kernel: the default 3-barrier kernel (bf16, B <= 2; VX_DECODE_FUSED=1: B <= 8).
#include "decode_fused_common.cuh"

namespace vx {
namespace {
template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_fused_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* hs = dsm + MR * kD;             // [MR][kHS] FFN hidden units of this CTA
  float* scratch = hs + MR * kHS;        // attention / sampling / LN-reduction scratch
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);            // residual-stream columns this CTA writes back / accumulator columns it re-zeroes
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // this CTA's attention unit, fixed for the whole launch
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;
  int att_epoch = 0;
  long long* const accO2 = p.acc_out;   // [2][B][1024]
  long long* const accF2 = p.acc_ffn;   // [2][B][1024]
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed (x_0 -> xa) ---------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b) && !(step == 0 && p.resume && !ld_cg(p.pending + b))) sample_utterance(p, b, scratch);   // resumed launch: a continuing utterance sampled its token before the pause
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (p.stop_below > 0 && ld_cg(bt.n_active) <= p.stop_below) break;   // continuous batching: hand a finished slot back to the host
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      // ---- P1: x_l = xmid_{l-1} + FFN2_{l-1} (+bias); LN1; QKV rows --------------------------------------------------------
      if (has_unit && tid == 0) {
        const int n = bt.S[u_b] + bt.P[u_b] + ld_cg(bt.T + u_b) + 1;
        int chunk = (n + ns - 1) / ns;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)u_b * kH + u_h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
          l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
          l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * sizeof(WT));
        }
      }
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
        zero_slice(accO_prev, B, o0, o1);   // consumed by P3 of layer l-1
      }
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, Wq, kD, q0, q1, sx, B, L.qkv_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows ride through the attention phase
      if (has_unit) wo_issue(wp, Wo, u_h, wr0, wr1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention unit + head rendezvous + out-proj share ------------------------------------------------------
      if (has_unit && !ld_cg(bt.finished + u_b))
        attn_unit_fused(p, l, u_b, u_h, u_split, scratch, att_epoch, wp, Wo, wr0, wr1, accO);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by P1 of this layer
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: xmid_l = x_l + out-proj (+bias); LN2; FFN1 rows + ReLU; FFN2 share --------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1, p.layer[l + 1].qkv_b, nullptr);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: x_12 = xmid_11 + FFN2_11 (+bias); final LayerNorm; predict rows -> logits -----------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by P3 of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue<WT, MR>(pre, Wqkv0, kD, q0, q1, p.layer[0].qkv_b, nullptr);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t fused_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;             // attention
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;     // sampling
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_fused_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_fused_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_fused_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}


}  // namespace

cudaError_t launch_decode_fused(const PParams& p, int mr, int grid, cudaStream_t st) {
  switch (mr) {
    case 1: return fused_launch<1>(p, grid, st);
    case 2: return fused_launch<2>(p, grid, st);
    case 4: return fused_launch<4>(p, grid, st);
    default: return fused_launch<8>(p, grid, st);
  }
}

}  // namespace vx
