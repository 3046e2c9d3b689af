// Experimental two-barrier decode kernels: head-group exchange through flag-carrying stores (v4) and through the
// distributed shared memory of an 8-CTA cluster (cluster variant).
#include "decode_fused_common.cuh"

namespace vx {
namespace {
// =================================================================================================================
// Two-barrier variant (bf16 weights, VX_DECODE_V4=1, experimental): 2 grid barriers per layer (26 per step).
//   PA  stage x_l, LN1, the unit's share of ITS HEAD's q/k/v rows (192 rows of Wqkv split over the head's key-split
//       CTAs) -> exchanged inside the head group through flag-carrying 8-byte stores (value + epoch in one store,
//       readers poll the data itself: one L2 round trip, no fence, no ticket) -> attention on the unit's key split ->
//       partials exchanged the same way -> fold -> out-proj share -> fixed-point accumulator      | grid barrier
//   PB  residual + LN2 + FFN1 rows + ReLU + FFN2 share -> fixed-point accumulator                  | grid barrier
// Only the two cross-head sums (out-proj, FFN2) need the whole grid; everything else is local to the <= 9 CTAs of a head.
// =================================================================================================================
// row i (0..191) of head h's q|k|v block -> row of Wqkv [3072][1024]
__device__ __forceinline__ int qkv_head_row(int h, int i) { return (i >> 6) * kD + h * kDH + (i & 63); }

// attention unit of the two-barrier variant: q/k/v arrive through qkv_ll, partials leave through part_ll
__device__ void attn_unit_v4(const PParams& p, int layer, int b, int h, int split, float* smem, uint32_t epoch,
                             const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int ns = p.nsplit;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + ns - 1) / ns;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  // ---- q/k/v of this head from the group's flag-carrying stores -------------------------------------------------
  if (tid < 3 * kDH) {
    const float v = ll_wait(p.qkv_ll + ((size_t)b * kH + h) * (3 * kDH) + tid, epoch);
    if (tid < kDH) sq[tid] = v * 0.125f;
    else smem[tid] = v;              // sknew | svnew are contiguous behind sq
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
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
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  // ---- this split's partial (m, l, o[64]) to the group ---------------------------------------------------------------
  unsigned long long* pg = p.part_ll + ((size_t)b * kH + h) * ns * (kDH + 2);
  if (tid < kDH) {
    float od = 0.f;
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
    ll_store(pg + (size_t)split * (kDH + 2) + 2 + tid, od, epoch);
  }
  if (tid == 0) {
    ll_store(pg + (size_t)split * (kDH + 2), m, epoch);
    ll_store(pg + (size_t)split * (kDH + 2) + 1, l, epoch);
  }
  // ---- fold all splits (fixed order); every load of a round is in flight at once, rounds repeat until all flags match --
  if (tid < kDH) {
    constexpr int U = 12;
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
          ok = ok && (uint32_t)(a >> 32) == epoch && (uint32_t)(c >> 32) == epoch && (uint32_t)(d >> 32) == epoch;
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
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns -------------------------------------------------------
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

// request rows map(i), i = warp + j*16 < cnt, of W [.][1024]
template <int MR, typename Map>
__device__ __forceinline__ void gemv_issue_map(WPre<bf16, MR>& pre, const bf16* __restrict__ W, int cnt, const float* __restrict__ bias, Map map) {
  const int warp = threadIdx.x >> 5;
#pragma unroll
  for (int j = 0; j < 2; ++j) {
    const int i = warp + j * PNW;
    pre.bias[j] = 0.f;
    if (i < cnt) {
      const int n = map(i);
      load_row_raw<bf16, MR>(W, kD, n, pre.r[j]);
      if (bias) pre.bias[j] = bias[n];
    }
  }
}
// consume them (and any rows beyond 2 per warp); emit(i, m, value) by lane m
template <int MR, typename Map, typename Emit>
__device__ __forceinline__ void gemv_finish_map(WPre<bf16, MR>& pre, const bf16* __restrict__ W, int cnt, const float* __restrict__ sx,
                                                int B, const float* __restrict__ bias, Map map, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int i0 = warp, i1 = warp + PNW;
  if (i0 < cnt) {  // warp-uniform
    float a0[MR], a1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      float w0[8], w1[8];
      raw_to_f32<bf16>(pre.r[0], it, w0);
      raw_to_f32<bf16>(pre.r[1], it, w1);
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float4 x0 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8]);
        const float4 x1 = *reinterpret_cast<const float4*>(&sx[m * kD + it * 256 + lane * 8 + 4]);
        const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) { a0[m] = fmaf(w0[i], xv[i], a0[m]); a1[m] = fmaf(w1[i], xv[i], a1[m]); }
      }
    }
    float m0 = 0.f, m1 = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m) {
      const float s0 = warp_sum(a0[m]), s1 = warp_sum(a1[m]);
      if (lane == m) { m0 = s0; m1 = s1; }
    }
    if (lane < B) {
      emit(i0, lane, m0 + pre.bias[0]);
      if (i1 < cnt) emit(i1, lane, m1 + pre.bias[1]);
    }
  }
  for (int i = warp + 2 * PNW; i < cnt; i += PNW) {   // more than 32 rows per CTA (fewer key splits)
    const int n = map(i);
    uint4 raw[WPre<bf16, MR>::kVec];
    load_row_raw<bf16, MR>(W, kD, n, raw);
    float acc[MR];
    row_dot<bf16, MR>(raw, sx, acc);
    float mine = 0.f;
#pragma unroll
    for (int m = 0; m < MR; ++m)
      if (lane == m) mine = acc[m];
    if (lane < B) emit(i, lane, mine + (bias ? bias[n] : 0.f));
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_v4_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [MR][1024]
  float* hs = dsm + MR * kD;             // [MR][kHS]
  float* scratch = hs + MR * kHS;
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int o0, o1, f0, f1, pr0, pr1;
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;            // out-proj rows of this unit
  const int qi0 = (u_split * 3 * kDH) / ns, qi1 = ((u_split + 1) * 3 * kDH) / ns;   // its share of the head's q|k|v rows
  const int qcnt = has_unit ? qi1 - qi0 : 0;
  auto qmap = [&](int i) { return qkv_head_row(u_h, qi0 + i); };
  long long* const accO2 = p.acc_out;
  long long* const accF2 = p.acc_ffn;
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), qcnt, p.layer[0].qkv_b, qmap);
  if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
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
      const uint32_t ll_epoch = (uint32_t)(step * kL + l + 1);
      const bool unit_live = has_unit && !ld_cg(bt.finished + u_b);
      // ---- PA ------------------------------------------------------------------------------------------------------
      if (has_unit && tid == 0) {   // pull this unit's K/V range towards L2 while the q/k/v rows are computed
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
      }
      zero_slice(accO_prev, B, o0, o1);   // consumed by PB of the previous layer (layer 0: by PB of layer 11, re-zeroed twice: harmless)
      PSTAMP(3);
      // only this utterance's row of sx matters to the unit; all B rows are computed and row u_b is emitted
      gemv_finish_map<MR>(pre, Wq, qcnt, sx, B, L.qkv_b, qmap, [&](int i, int m, float v) {
        if (m == u_b && unit_live) ll_store(p.qkv_ll + ((size_t)u_b * kH + u_h) * (3 * kDH) + qi0 + i, v, ll_epoch);
      });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows ride through the attention part
      if (unit_live) attn_unit_v4(p, l, u_b, u_h, u_split, scratch, ll_epoch, wp, Wo, wr0, wr1, accO);
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- PB ------------------------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by PA of this layer
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), qcnt, p.layer[l + 1].qkv_b, qmap);
        if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[l + 1].out_w), u_h, wr0, wr1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF --------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by PB of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), qcnt, p.layer[0].qkv_b, qmap);
    if (has_unit) wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

// =================================================================================================================
// Cluster variant (bf16 weights, B = 1, VX_DECODE_CLUSTER=1, experimental): the two-barrier layout of the v4 kernel with
// the head group mapped onto a thread-block CLUSTER: 16 clusters x 8 CTAs = 128 CTAs (8 GPCs x two 8-CTA clusters is what
// fits co-resident on the 148 SMs), cluster c = head c, rank = key split.  q/k/v of the head and the attention partials
// are written to the producer's own shared memory and read by the other ranks through distributed shared memory behind
// barrier.cluster (~0.2 us) instead of through L2 flags (~1.5-2 us per exchange).  Grid barriers as before.
// =================================================================================================================
__device__ __forceinline__ uint32_t cluster_rank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// read a float from the same shared-memory offset in cluster rank `rank`
__device__ __forceinline__ float dsmem_ld(const float* local, uint32_t rank) {
  const uint32_t la = (uint32_t)__cvta_generic_to_shared(local);
  uint32_t ra;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(la), "r"(rank));
  float v;
  asm volatile("ld.shared::cluster.f32 %0, [%1];" : "=f"(v) : "r"(ra) : "memory");
  return v;
}

constexpr int kCS = 8;                 // cluster size = key splits per head
constexpr int kQShare = 3 * kDH / kCS; // 24 of the head's 192 q|k|v rows per CTA

// xchg: [0,24) this CTA's q|k|v values, [32, 32+66) its attention partial (m, l, o[64]); read by the other ranks via DSMEM
__device__ void attn_unit_cluster(const PParams& p, int layer, int h, int split, float* smem, float* xchg,
                                  const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   (reused for the folded head output)
  float* sknew = smem + 64;
  float* svnew = smem + 128;
  float* red = smem + 192;
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int b = 0;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + kCS - 1) / kCS;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  // ---- q/k/v of this head: element i lives in rank i / 24 at xchg[i % 24] -----------------------------------------------
  cluster_sync_all();                // every rank's q|k|v share is in its xchg[0,24)
  if (tid < 3 * kDH) {
    const float v = dsmem_ld(xchg + (tid % kQShare), (uint32_t)(tid / kQShare));
    if (tid < kDH) sq[tid] = v * 0.125f;
    else smem[tid] = v;              // sknew | svnew are contiguous behind sq
  }
  __syncthreads();
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  if (split == 0 && tid < kDH) {
    store_as(kb + (size_t)pos * kDH + tid, sknew[tid]);
    store_as(vb + (size_t)pos * kDH + tid, svnew[tid]);
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
  for (int j = j0 + g; j < j1; j += NG) {
    const float pr = sp[j - j0];
    float w[8];
    if (j == pos) {
#pragma unroll
      for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
    } else {
      ldkv8(vb + (size_t)j * kDH + dc, w);
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fmaf(pr, w[i], acc[i]);
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  // ---- this split's partial into its own xchg[32..98) ------------------------------------------------------------------
  float* mypart = xchg + 32;
  if (tid < kDH) {
    float od = 0.f;
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
    mypart[2 + tid] = od;
  }
  if (tid == 0) { mypart[0] = m; mypart[1] = l; }
  cluster_sync_all();                // every rank's partial is visible cluster-wide
  // ---- fold the 8 splits (fixed order) ---------------------------------------------------------------------------------------
  if (tid < kDH) {
    float mv[kCS], lv[kCS], ov[kCS];
#pragma unroll
    for (int s = 0; s < kCS; ++s) {
      mv[s] = dsmem_ld(mypart, (uint32_t)s);
      lv[s] = dsmem_ld(mypart + 1, (uint32_t)s);
      ov[s] = dsmem_ld(mypart + 2 + tid, (uint32_t)s);
    }
    float M = -INFINITY;
#pragma unroll
    for (int s = 0; s < kCS; ++s) M = fmaxf(M, mv[s]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int s = 0; s < kCS; ++s) {
      const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
      L += w * lv[s];
      O += w * ov[s];
    }
    sq[tid] = O / L;
  }
  __syncthreads();
  // ---- out-proj share: rows [r0,r1) (128 rows) x this head's 64 columns -------------------------------------------------------
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
  // the partial / q|k|v slots are rewritten two grid barriers from now, long after every rank has read them
}

__global__ void __launch_bounds__(PNT, 1) ar_decode_cluster_kernel(const PParams p) {
  using WT = bf16;
  constexpr int MR = 1;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [1024]
  float* hs = dsm + MR * kD;             // [kHS]
  float* xchg = hs + MR * kHS;           // [128] cluster exchange slots (same offset in every CTA of the cluster)
  float* scratch = xchg + 128;
  const Batch& bt = p.bt;
  const int B = 1;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int o0, o1, f0, f1, pr0, pr1;
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int u_h = blockIdx.x / kCS, u_split = (int)cluster_rank();   // cluster c = head c (blocks 8c .. 8c+7), rank = key split
  const int wr0 = (u_split * kD) / kCS, wr1 = ((u_split + 1) * kD) / kCS;
  const int qi0 = u_split * kQShare;
  auto qmap = [&](int i) { return qkv_head_row(u_h, qi0 + i); };
  long long* const accO2 = p.acc_out;
  long long* const accF2 = p.acc_ffn;
  const size_t accN = (size_t)B * kD;

  WPre<WT, MR> pre;
  WoPre wp;
  LnPre lp;
  gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), kQShare, p.layer[0].qkv_b, qmap);
  wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  bool live = true;
  for (int step = 0; step < p.max_steps && live; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    if (blockIdx.x == 0 && !ld_cg(bt.finished)) sample_utterance(p, 0, scratch);
    zero_slice(accF2 + accN, B, o0, o1);
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
      // ---- PA ------------------------------------------------------------------------------------------------------
      if (tid == 0) {
        const int n = bt.S[0] + bt.P[0] + ld_cg(bt.T) + 1;
        int chunk = (n + kCS - 1) / kCS;
        chunk = (chunk + 31) & ~31;
        const int j0 = u_split * chunk, j1 = min(n - 1, j0 + chunk);
        if (j1 > j0) {
          const size_t off = ((size_t)l * p.kv_layer_elems + (size_t)u_h * p.ctx_cap * kDH + (size_t)j0 * kDH) * sizeof(WT);
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
      }
      zero_slice(accO_prev, B, o0, o1);
      PSTAMP(3);
      gemv_finish_map<MR>(pre, Wq, kQShare, sx, B, L.qkv_b, qmap, [&](int i, int, float v) { xchg[i] = v; });
      PSTAMP(4);
      gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      attn_unit_cluster(p, l, u_h, u_split, scratch, xchg, wp, Wo, wr0, wr1, accO);
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- PB ------------------------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);
      PSTAMP(3);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
        gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kQShare, p.layer[l + 1].qkv_b, qmap);
        wo_issue(wp, reinterpret_cast<const WT*>(p.layer[l + 1].out_w), u_h, wr0, wr1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, nullptr, nullptr);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF --------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_map<MR>(pre, reinterpret_cast<const WT*>(p.layer[0].qkv_w), kQShare, p.layer[0].qkv_b, qmap);
    wo_issue(wp, reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, wr0, wr1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
  cluster_sync_all();   // no CTA of the cluster leaves while a peer may still read its shared memory
}

static cudaError_t cluster_launch(const PParams& p, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)kD + kHS + 128 + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(kH * kCS);
  cfg.blockDim = dim3(PNT);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attrs[2];
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  attrs[1].id = cudaLaunchAttributeCooperative;
  attrs[1].val.cooperative = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 2;
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, ar_decode_cluster_kernel, &cfg);
  if (e != cudaSuccess) return e;
  if (nclusters < kH) return cudaErrorLaunchOutOfResources;   // all 16 clusters must be co-resident (grid barrier)
  return cudaLaunchKernelEx(&cfg, ar_decode_cluster_kernel, p);
}

template <int MR>
static cudaError_t v4_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_v4_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_v4_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_v4_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}


}  // namespace

cudaError_t launch_decode_v4(const PParams& p, int mr, int grid, cudaStream_t st) {
  switch (mr) {
    case 1: return v4_launch<1>(p, grid, st);
    case 2: return v4_launch<2>(p, grid, st);
    case 4: return v4_launch<4>(p, grid, st);
    default: return v4_launch<8>(p, grid, st);
  }
}
cudaError_t launch_decode_cluster(const PParams& p, cudaStream_t st) { return cluster_launch(p, st); }
static_assert(kCS == kDecodeCS && kHS == kDecodeHS, "keep kDecodeCS / kDecodeHS in decode_persist.h in step");

}  // namespace vx
