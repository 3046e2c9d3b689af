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
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
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

// =================================================================================================================
// Ring variant (bf16 weights, B <= 8): the CTA's weight-row slice of every GEMV phase -- plus that phase's LayerNorm
// vectors and biases -- is staged into shared memory by the TMA engine (cp.async.bulk -> mbarrier, 1-D bulk copies of
// the contiguous slice) TWO phases ahead through a 2-slot ring, so HBM streams continuously across the grid barriers
// and no phase waits on DRAM latency; the dot products read weights from shared memory with conflict-free 128-bit loads.
// =================================================================================================================

constexpr int RW_ROWS = 32;
constexpr int RW_BYTES = RW_ROWS * 2048;             // weight rows (K = 1024 bf16)
constexpr int RLN_OFF = RW_BYTES;                    // gamma | beta | previous FFN2 bias (3 x 4 KB)
constexpr int RBIAS_OFF = RLN_OFF + 3 * 4096;        // bias | bias2 (2 x 256 B)
constexpr int RSTAGE_BYTES = RBIAS_OFF + 512;        // 78,336 B per ring slot

struct RSlice {
  const bf16* W; size_t ldw; int r0, r1, N;
  const float *bias, *bias2, *lg, *lb, *lf;
};

// thread 0: arm the slot's mbarrier and launch the bulk copies of slice `s`
__device__ __forceinline__ void ring_issue(uint8_t* slot, uint64_t* bar, const RSlice& s) {
  using namespace tc;
  const int rows = s.r1 - s.r0;
  const int a0 = s.r0 & ~3;
  int nb = ((s.r1 - a0) + 3) & ~3;
  if (a0 + nb > s.N) nb = s.N - a0;
  uint32_t bytes = (uint32_t)rows * 2048u;
  if (s.lg) bytes += 8192u;
  if (s.lf) bytes += 4096u;
  if (s.bias) bytes += (uint32_t)nb * 4u;
  if (s.bias2) bytes += (uint32_t)nb * 4u;
  fence_proxy_async();  // the slot was last read through the generic proxy
  mbar_expect_tx(bar, bytes);
  if (s.ldw == (size_t)kD) {
    bulk_g2s(slot, s.W + (size_t)s.r0 * kD, (uint32_t)rows * 2048u, bar);
  } else {
    for (int r = 0; r < rows; ++r) bulk_g2s(slot + r * 2048, s.W + (size_t)(s.r0 + r) * s.ldw, 2048u, bar);
  }
  if (s.lg) { bulk_g2s(slot + RLN_OFF, s.lg, 4096u, bar); bulk_g2s(slot + RLN_OFF + 4096, s.lb, 4096u, bar); }
  if (s.lf) bulk_g2s(slot + RLN_OFF + 8192, s.lf, 4096u, bar);
  if (s.bias) bulk_g2s(slot + RBIAS_OFF, s.bias + a0, (uint32_t)nb * 4u, bar);
  if (s.bias2) bulk_g2s(slot + RBIAS_OFF + 256, s.bias2 + a0, (uint32_t)nb * 4u, bar);
}

template <int MR>
__device__ __forceinline__ void ring_row_dot(const uint8_t* wrow, const float* __restrict__ sx, float (&acc)[MR]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = 0.f;
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const uint4 r = *reinterpret_cast<const uint4*>(wrow + it * 512 + lane * 16);
    const uint32_t w4[4] = {r.x, r.y, r.z, r.w};
    float w[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) { w[2 * i] = __uint_as_float(w4[i] << 16); w[2 * i + 1] = __uint_as_float(w4[i] & 0xffff0000u); }
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
      const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
      float a = acc[m], c = 0.f;  // two chains halve the FMA dependency depth
      a = fmaf(w[0], x0.x, a); c = fmaf(w[1], x0.y, c); a = fmaf(w[2], x0.z, a); c = fmaf(w[3], x0.w, c);
      a = fmaf(w[4], x1.x, a); c = fmaf(w[5], x1.y, c); a = fmaf(w[6], x1.z, a); c = fmaf(w[7], x1.w, c);
      acc[m] = a + c;
    }
  }
#pragma unroll
  for (int m = 0; m < MR; ++m) acc[m] = warp_sum(acc[m]);
}

// rows of the slice from the ring slot; addend(n, m) evaluated by lane m before the dots; emit(n, m, v) by lane m.
template <int MR, typename Add, typename Emit>
__device__ __forceinline__ void ring_finish(const uint8_t* slot, const RSlice& s, const float* __restrict__ sx, int B,
                                            Add addend, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int a0 = s.r0 & ~3;
  const float* sb = reinterpret_cast<const float*>(slot + RBIAS_OFF);
  const float* sb2 = reinterpret_cast<const float*>(slot + RBIAS_OFF + 256);
  for (int n = s.r0 + warp; n < s.r1; n += PNW) {
    float ad = (lane < B) ? addend(n, lane) : 0.f;
    if (s.bias) ad += sb[n - a0];
    if (s.bias2) ad += sb2[n - a0];
    float acc[MR];
    ring_row_dot<MR>(slot + (size_t)(n - s.r0) * 2048, sx, acc);
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m)
      if (lane == m) mine = acc[m];
    if (lane < B) emit(n, lane, mine + ad);
  }
}

// LayerNorm staging with the parameters taken from the ring slot
template <int MR, typename Val>
__device__ __forceinline__ void ring_stage_ln(float* sx, int B, Val value, const uint8_t* slot, bool has_f, float* red) {
  const float* lnp = reinterpret_cast<const float*>(slot + RLN_OFF);
  LnPre lp;
  const int t = threadIdx.x;
  lp.g0 = lnp[t]; lp.g1 = lnp[t + PNT]; lp.b0 = lnp[1024 + t]; lp.b1 = lnp[1024 + t + PNT];
  lp.f0 = has_f ? lnp[2048 + t] : 0.f; lp.f1 = has_f ? lnp[2048 + t + PNT] : 0.f;
  stage_ln<MR>(sx, B, value, lp, red);
}

// out-proj input: softmax-combine the key splits.  (m, l) of every (b, head, split) first go to shared memory, then
// each thread folds its columns with all split loads in flight at once.
template <int MR>
__device__ __forceinline__ void stage_att(float* sx, const PParams& p, int B, float* sml /*[B*16*nsplit*2]*/) {
  const int tid = threadIdx.x;
  const int npairs = B * kH * p.nsplit;
  for (int i = tid; i < npairs * 2; i += PNT) {
    const int pair = i >> 1, which = i & 1;
    sml[i] = __ldcg(p.part + (size_t)pair * (kDH + 2) + which);
  }
  __syncthreads();
  for (int i = tid; i < B * kD; i += PNT) {
    const int b = i / kD, c = i % kD, h = c >> 6, d = c & 63;
    const float* ml = sml + (size_t)(b * kH + h) * p.nsplit * 2;
    const float* po = p.part + ((size_t)(b * kH + h) * p.nsplit) * (kDH + 2) + 2 + d;
    float o[16];
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) o[s2] = (s2 < p.nsplit) ? __ldcg(po + (size_t)s2 * (kDH + 2)) : 0.f;
    float M = -INFINITY;
    for (int s2 = 0; s2 < p.nsplit; ++s2) M = fmaxf(M, ml[2 * s2]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int s2 = 0; s2 < 16; ++s2) {
      if (s2 < p.nsplit) {
        const float ms = ml[2 * s2];
        const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
        L += w * ml[2 * s2 + 1];
        O += w * o[s2];
      }
    }
    for (int s2 = 16; s2 < p.nsplit; ++s2) {  // nsplit > 16 only for very long contexts
      const float ms = ml[2 * s2];
      const float w = (ms == -INFINITY) ? 0.f : expf(ms - M);
      L += w * ml[2 * s2 + 1];
      O += w * __ldcg(po + (size_t)s2 * (kDH + 2));
    }
    sx[i] = ld_cg(p.bt.finished + b) ? 0.f : O / L;
  }
  __syncthreads();
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_ring_kernel(const PParams p) {
  using namespace tc;
  extern __shared__ __align__(128) uint8_t rsm[];
  uint8_t* ring = rsm;                                                   // [2][RSTAGE_BYTES]
  float* sx = reinterpret_cast<float*>(rsm + 2 * RSTAGE_BYTES);          // [MR][1024]
  float* scratch = sx + MR * kD;                                         // attention / sampling / LN scratch
  uint64_t* rbar = reinterpret_cast<uint64_t*>(rsm + 2 * RSTAGE_BYTES + (size_t)MR * kD * 4 + 32768);
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int per = gridDim.x / 4;
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  // GEMV phase c of the 49-phase cycle of one step: 4 per layer (qkv, out, ff1, ff2) + the predict layer
  auto slice_of = [&](int c) {
    RSlice s{};
    if (c == 4 * kL) {
      s.W = reinterpret_cast<const bf16*>(p.predict_w); s.ldw = kD; s.r0 = pr0; s.r1 = pr1; s.N = kARLogits;
      s.lg = p.norm_g; s.lb = p.norm_b; s.lf = p.layer[kL - 1].ff2_b;
      return s;
    }
    const int l = c >> 2, kind = c & 3;
    const PLayer& L = p.layer[l];
    const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
    if (kind == 0) { s.W = (const bf16*)L.qkv_w; s.ldw = kD; s.r0 = q0; s.r1 = q1; s.N = 3 * kD; s.bias = L.qkv_b; s.lg = L.n1_g; s.lb = L.n1_b; s.lf = fb_prev; }
    else if (kind == 1) { s.W = (const bf16*)L.out_w; s.ldw = kD; s.r0 = o0; s.r1 = o1; s.N = kD; s.bias = L.out_b; s.bias2 = fb_prev; }
    else if (kind == 2) { s.W = (const bf16*)L.ff1_w; s.ldw = kD; s.r0 = f0; s.r1 = f1; s.N = kFF; s.bias = L.ff1_b; s.lg = L.n2_g; s.lb = L.n2_b; }
    else { s.W = (const bf16*)L.ff2_w + (size_t)kc * kD; s.ldw = kFF; s.r0 = g0; s.r1 = g1; s.N = kD; }
    return s;
  };
  constexpr int CYC = 4 * kL + 1;
  if (tid == 0) {
    mbar_init(&rbar[0], 1); mbar_init(&rbar[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  int g = 0;  // GEMV phases consumed so far (phase g lives in ring slot g & 1, its mbarrier parity is (g >> 1) & 1)
  if (tid == 0) {
    ring_issue(ring, &rbar[0], slice_of(0));
    ring_issue(ring + RSTAGE_BYTES, &rbar[1], slice_of(1));
  }
  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  // finish GEMV phase g: after every thread has left the slot, thread 0 refills it with phase g + 2, then the grid barrier
  auto next_phase = [&]() {
    __syncthreads();
    if (tid == 0) ring_issue(ring + (g & 1) * RSTAGE_BYTES, &rbar[g & 1], slice_of((g + 2) % CYC));
    ++g;
    grid_barrier(p.bar, epoch, prof);
  };

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0 ----
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      // ---- P1: LN1 + QKV ----
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        if (tid == 0) {  // K/V ranges of the coming attention units -> L2
          const int split = u % p.nsplit, bh = u / p.nsplit, b = bh / kH, h = bh % kH;
          const int n = bt.S[b] + bt.P[b] + ld_cg(bt.T + b) + 1;
          int chunk = (n + p.nsplit - 1) / p.nsplit;
          chunk = (chunk + 31) & ~31;
          const int j0 = split * chunk, j1 = min(n - 1, j0 + chunk);
          if (j1 > j0) {
            const size_t off = ((size_t)l * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH + (size_t)j0 * kDH) * 2;
            l2_prefetch(reinterpret_cast<const char*>(p.kc) + off, (size_t)(j1 - j0) * kDH * 2);
            l2_prefetch(reinterpret_cast<const char*>(p.vc) + off, (size_t)(j1 - j0) * kDH * 2);
          }
        }
      }
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, slot, s.lf != nullptr, scratch);
        PSTAMP(3);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
      // ---- P2: attention ----
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        attn_unit<bf16>(p, l, b, h, split, scratch);
      }
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: combine splits + out-proj + residual ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        stage_att<MR>(sx, p, B, scratch);
        PSTAMP(3);
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_finish<MR>(slot, s, sx, B, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                        [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
      // ---- P4: LN2 + FFN1 + ReLU ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_stage_ln<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, slot, false, scratch);
        PSTAMP(3);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
        PSTAMP(4);
        next_phase();
      }
      // ---- P5: FFN2 split-K ----
      {
        const RSlice s = slice_of(g % CYC);
        const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
        stage_plain(sx, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
        PSTAMP(3);
        mbar_wait(&rbar[g & 1], (g >> 1) & 1);
        PSTAMP(6);
        ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
        PSTAMP(4);
        next_phase();
      }
    }
    // ---- PF: final LN + predict ----
    {
      const RSlice s = slice_of(g % CYC);
      const uint8_t* slot = ring + (g & 1) * RSTAGE_BYTES;
      mbar_wait(&rbar[g & 1], (g >> 1) & 1);
      ring_stage_ln<MR>(sx, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, slot, true, scratch);
      ring_finish<MR>(slot, s, sx, B, [&](int, int) { return 0.f; },
                      [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
      next_phase();
    }
  }
  // drain the two bulk-copy groups still in flight before the CTA (and its shared memory) goes away
  mbar_wait(&rbar[g & 1], (g >> 1) & 1);
  mbar_wait(&rbar[(g + 1) & 1], ((g + 1) >> 1) & 1);
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t ring_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t smem = 2 * (size_t)RSTAGE_BYTES + (size_t)MR * kD * 4 + 32768 + 64;
  cudaError_t e = cudaFuncSetAttribute(ar_decode_ring_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_ring_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
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
  const char* mma_env = getenv("VX_DECODE_MMA");   // tensor-core batched variant (<= 32 rows per CTA slice); VX_DECODE_MMA=0 -> CUDA-core GEMV
  if (wt == 1 && mr >= 8 && grid >= 128 && !(mma_env && mma_env[0] == '0')) {
    return (int)launch_decode_mma(p, mr, grid, st);
  }
  if (wt == 1 && p.fused == 5) {  // experimental: 3-barrier kernel with TMA-staged weights (VX_DECODE_TMA=1)
    if (B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kDecodeHS) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_tma(p, mr, grid, st);
  }
  if (wt == 1 && p.fused == 4) {  // experimental trimmed 3-barrier variant (VX_DECODE_TRIM=<bits>)
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kDecodeHS || p.nsplit > 12) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_fused2(p, mr, grid, st);
  }
  if (wt == 1 && p.fused == 3) {  // experimental cluster variant (VX_DECODE_CLUSTER=1): B = 1, 16 clusters x 8 CTAs
    if (B != 1 || p.nsplit != kDecodeCS || (kFF + kH * kDecodeCS - 1) / (kH * kDecodeCS) > kDecodeHS) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_cluster(p, st);
  }
  if (wt == 1 && p.fused == 2) {  // experimental two-barrier variant (VX_DECODE_V4=1): same preconditions as the 3-barrier one
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kDecodeHS) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_v4(p, mr, grid, st);
  }
  if (wt == 1 && p.fused) {  // 3-barrier variant: requires one attention unit per CTA and <= kDecodeHS hidden units per CTA
    if (mr > 8 || B * kH * p.nsplit > grid || (kFF + grid - 1) / grid > kDecodeHS) return (int)cudaErrorInvalidConfiguration;
    return (int)launch_decode_fused(p, mr, grid, st);
  }
  if (wt == 1 && mr <= 8 && grid >= 128 && getenv("VX_DECODE_RING")) {  // experimental TMA-ring variant (slower: smem-bound dots)  // 28 rows/CTA max fits the 32-row ring slot
    switch (mr) {
      case 1: return (int)ring_launch<1>(p, grid, st);
      case 2: return (int)ring_launch<2>(p, grid, st);
      case 4: return (int)ring_launch<4>(p, grid, st);
      default: return (int)ring_launch<8>(p, grid, st);
    }
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
