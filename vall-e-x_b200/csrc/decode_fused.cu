// Fused persistent decode kernels: the default 3-barrier kernel (bf16, B <= 2 / <= 8) and its trimmed A/B variant.
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
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
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

// one-pass LayerNorm staging (trim bit 8): sum and sum of squares in ONE block reduction (one __syncthreads less per
// staging than the two-pass form; variance = E[x^2] - mean^2, clamped at 0)
template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_x1(float* sx, int B, Val value, const LnPre& lp, float* red, float* xout, int c0, int c1) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = MR < 8 ? MR : 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v0[R], v1[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      v0[m] = (m < nb) ? value(mb + m, tid) + lp.f0 : 0.f;
      v1[m] = (m < nb) ? value(mb + m, tid + PNT) + lp.f1 : 0.f;
    }
    if (xout) {
#pragma unroll
      for (int m = 0; m < R; ++m) {
        if (m < nb) {
          if (tid >= c0 && tid < c1) xout[(size_t)(mb + m) * kD + tid] = v0[m];
          if (tid + PNT >= c0 && tid + PNT < c1) xout[(size_t)(mb + m) * kD + tid + PNT] = v1[m];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float s1 = warp_sum(v0[m] + v1[m]);
      const float s2 = warp_sum(v0[m] * v0[m] + v1[m] * v1[m]);
      if (lane == 0) { red[warp * 8 + m] = s1; red[PNW * 8 + warp * 8 + m] = s2; }
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) { s1 += red[w * 8 + m]; s2 += red[PNW * 8 + w * 8 + m]; }
        const float mean = s1 / (float)kD;
        const float var = fmaxf(s2 / (float)kD - mean * mean, 0.f);
        const float rstd = 1.0f / sqrtf(var + 1e-5f);
        sx[(mb + m) * kD + tid] = __fadd_rn(__fmul_rn((v0[m] - mean) * rstd, lp.g0), lp.b0);
        sx[(mb + m) * kD + tid + PNT] = __fadd_rn(__fmul_rn((v1[m] - mean) * rstd, lp.g1), lp.b1);
      }
    }
    __syncthreads();
  }
}

// =================================================================================================================
// Trimmed 3-barrier variant (VX_DECODE_TRIM=<bits>, experimental A/B switches on top of ar_decode_fused_kernel):
//   1  V chunks of the attention unit requested together with the K rows (one memory round trip less)
//   2  the CTA's W2^T share prefetched into registers at the start of the FFN phase
//   4  attention partials exchanged through flag-carrying stores instead of fence + ticket + spin + reload
//   8  one-pass LayerNorm statistics (one block reduction per staging instead of two)
// =================================================================================================================
__device__ void attn_unit_fused2(const PParams& p, int layer, int b, int h, int split, float* smem, int& att_epoch, uint32_t ll_epoch,
                                const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();
  if (tid < kDH) {
    sq[tid] = __ldcg(row + tid) * 0.125f;
    sknew[tid] = __ldcg(row + kD + tid);
    svnew[tid] = __ldcg(row + 2 * kD + tid);
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
  }
  // trim bit 1: this thread's V chunks (keys j0+g, +64, ...) are requested now, so they travel with the K rows
  constexpr int VP = 4;
  uint4 vraw[VP];
  {
    const int g_ = tid >> 3, dc_ = (tid & 7) * 8;
#pragma unroll
    for (int t = 0; t < VP; ++t) {
      const int j = j0 + g_ + t * (PNT / 8);
      vraw[t] = ((p.trim & 1) && j < j1 && j != pos) ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + dc_)) : make_uint4(0, 0, 0, 0);
    }
  }
  float mloc = -INFINITY;
  for (int j = j0 + tid; j < j1; j += PNT) {
    float s = 0.f;
    if (j == pos) {
#pragma unroll
      for (int d = 0; d < kDH; ++d) s = fmaf(sq[d], kv_round<bf16>(sknew[d]), s);
    } else {
      const bf16* kr = kb + (size_t)j * kDH;
#pragma unroll
      for (int d = 0; d < kDH; d += 8) {
        float w[8];
        ldkv8(kr + d, w);
#pragma unroll
        for (int i = 0; i < 8; ++i) s = fmaf(sq[d + i], w[i], s);
      }
    }
    sp[j - j0] = s;
    mloc = fmaxf(mloc, s);
  }
  const float m = block_max(mloc, red);
  float lloc = 0.f;
  for (int j = j0 + tid; j < j1; j += PNT) {
    const float pr = expf(sp[j - j0] - m);
    sp[j - j0] = pr;
    lloc += pr;
  }
  const float l = block_sum(lloc, red);
  constexpr int NG = PNT / 8;
  const int g = tid >> 3, dc = (tid & 7) * 8;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int vt = 0;
  for (int j = j0 + g; j < j1; j += NG, ++vt) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else if ((p.trim & 1) && vt < VP) {
      uint4 r = vraw[0];
#pragma unroll
      for (int t = 1; t < VP; ++t)
        if (vt == t) r = vraw[t];
      raw_to_f32<bf16>(&r, 0, w);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
  if (p.trim & 4) {
    // trim bit 4: partials carry their own flag (value + epoch in one 8-byte store); the fold polls the data itself
    unsigned long long* pg = p.part_ll + ((size_t)b * kH + h) * p.nsplit * (kDH + 2);
    if (tid < kDH) ll_store(pg + (size_t)split * (kDH + 2) + 2 + tid, od, ll_epoch);
    if (tid == 0) {
      ll_store(pg + (size_t)split * (kDH + 2), m, ll_epoch);
      ll_store(pg + (size_t)split * (kDH + 2) + 1, l, ll_epoch);
    }
    if (tid < kDH) {
      constexpr int U = 12;
      const int ns = p.nsplit;
      float mv[U], lv[U], ov[U];
      bool ok;
      unsigned int spins = 0;
      do {
        ok = true;
#pragma unroll
        for (int s = 0; s < U; ++s) {
          if (s < ns) {
            const unsigned long long a = ll_load(pg + (size_t)s * (kDH + 2));
            const unsigned long long c = ll_load(pg + (size_t)s * (kDH + 2) + 1);
            const unsigned long long d = ll_load(pg + (size_t)s * (kDH + 2) + 2 + tid);
            ok = ok && (uint32_t)(a >> 32) == ll_epoch && (uint32_t)(c >> 32) == ll_epoch && (uint32_t)(d >> 32) == ll_epoch;
            mv[s] = __uint_as_float((uint32_t)a); lv[s] = __uint_as_float((uint32_t)c); ov[s] = __uint_as_float((uint32_t)d);
          } else {
            mv[s] = -INFINITY; lv[s] = 0.f; ov[s] = 0.f;
          }
        }
        if (++spins > (1u << 22)) __trap();
      } while (!ok);
      float M = -INFINITY;
#pragma unroll
      for (int s = 0; s < U; ++s) M = fmaxf(M, mv[s]);
      float L = 0.f, O = 0.f;
#pragma unroll
      for (int s = 0; s < U; ++s) {
        const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
        L += w * lv[s];
        O += w * ov[s];
      }
      sq[tid] = O / L;
    }
    __syncthreads();
  } else {
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  __syncthreads();
  // ---- rendezvous of the head's key splits (all co-resident: one unit per CTA) ---------------------------------
  att_epoch += p.nsplit;
  if (tid == 0) {
    int32_t* tk = p.tickets + b * kH + h;
    __threadfence();
    atomicAdd(tk, 1);
    if (p.nsplit > 1) {
      unsigned int spins = 0;
      int v;
      do {
        asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(tk) : "memory");
        if (++spins > (1u << 24)) __trap();
      } while (v - att_epoch < 0);
    }
  }
  __syncthreads();
  // ---- fold the splits (fixed order) -> normalised head output in shared memory ----------------------------------
  if (tid < kDH) sq[tid] = att_combined(p, b, h * kDH + tid);
  __syncthreads();
  }
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns --------------------------------------------------------
  const int grp = lane >> 3, sub = lane & 7;
  float xv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
  for (int pass = 0; pass * 64 < r1 - r0; ++pass) {
    const int nrow = r0 + pass * 64 + warp * 4 + grp;
    const bool ok = nrow < r1;
    uint4 raw;
    if (pass < 2) raw = wp.r[pass];
    else raw = ok ? ldw_raw(Wo + (size_t)nrow * kD + h * kDH + sub * 8) : make_uint4(0, 0, 0, 0);
    float w[8];
    raw_to_f32<bf16>(&raw, 0, w);
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    if (ok && sub == 0) fix_add(accO + (size_t)b * kD + nrow, s);
  }
}

struct W2Pre { uint32_t w[32]; };   // columns (2*tid, 2*tid+1) of the first 32 hidden units' rows of W2^T
__device__ __forceinline__ void ffn2_issue(W2Pre& wp2, const bf16* __restrict__ W2T, int f0, int f1) {
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + threadIdx.x;
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    wp2.w[i] = 0;
    if (i < f1 - f0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(wp2.w[i]) : "l"(base + (size_t)i * (kD / 2)));
  }
}
template <int MR>
__device__ __forceinline__ void ffn2_share_pre(const W2Pre& wp2, const bf16* __restrict__ W2T, int f0, int f1, const float* __restrict__ hs,
                                               int B, long long* accF) {
  const int tid = threadIdx.x;
  float acc[MR][2];
#pragma unroll
  for (int m = 0; m < MR; ++m) { acc[m][0] = 0.f; acc[m][1] = 0.f; }
#pragma unroll
  for (int i = 0; i < 32; ++i) {
    const float wa = __uint_as_float(wp2.w[i] << 16), wb = __uint_as_float(wp2.w[i] & 0xffff0000u);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float hv = hs[m * kHS + i];   // zero beyond f1-f0
      acc[m][0] = fmaf(wa, hv, acc[m][0]);
      acc[m][1] = fmaf(wb, hv, acc[m][1]);
    }
  }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + tid;
  for (int j = 32; j < f1 - f0; ++j) {   // grids smaller than 128 CTAs only
    uint32_t w;
    asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w) : "l"(base + (size_t)j * (kD / 2)));
    const float wa = __uint_as_float(w << 16), wb = __uint_as_float(w & 0xffff0000u);
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float hv = hs[m * kHS + j];
      acc[m][0] = fmaf(wa, hv, acc[m][0]);
      acc[m][1] = fmaf(wb, hv, acc[m][1]);
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) {
    if (m < B) {
      fix_add(accF + (size_t)m * kD + 2 * tid, acc[m][0]);
      fix_add(accF + (size_t)m * kD + 2 * tid + 1, acc[m][1]);
    }
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_fused2_kernel(const PParams p) {
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
  W2Pre wp2;
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
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
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
        if (p.trim & 8) {
          stage_ln_x1<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
        } else {
          stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
        }
      } else {
        if (p.trim & 8) {
          stage_ln_x1<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                         lp, scratch, p.xa, o0, o1);
        } else {
          stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                         lp, scratch, p.xa, o0, o1);
        }
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
        attn_unit_fused2(p, l, u_b, u_h, u_split, scratch, att_epoch, (uint32_t)(step * kL + l + 1), wp, Wo, wr0, wr1, accO);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by P1 of this layer
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: xmid_l = x_l + out-proj (+bias); LN2; FFN1 rows + ReLU; FFN2 share --------------------------------------
      if (p.trim & 2) ffn2_issue(wp2, W2T, f0, f1);   // trim bit 2: the W2^T share rides under staging + LN2 + FFN1
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      if (p.trim & 8) {
        stage_ln_x1<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                       lp, scratch, p.xb, o0, o1);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                       lp, scratch, p.xb, o0, o1);
      }
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      if (p.trim & 2) ffn2_share_pre<MR>(wp2, W2T, f0, f1, hs, B, accF);
      else ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
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
    if (p.trim & 8) {
      stage_ln_x1<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                     lp, scratch, nullptr, 0, 0);
    } else {
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                     lp, scratch, nullptr, 0, 0);
    }
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
static cudaError_t fused2_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_fused2_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_fused2_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_fused2_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
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
cudaError_t launch_decode_fused2(const PParams& p, int mr, int grid, cudaStream_t st) {
  switch (mr) {
    case 1: return fused2_launch<1>(p, grid, st);
    case 2: return fused2_launch<2>(p, grid, st);
    case 4: return fused2_launch<4>(p, grid, st);
    default: return fused2_launch<8>(p, grid, st);
  }
}

}  // namespace vx
