// Batch-1 decode with head groups as thread-block CLUSTERS: two grid barriers per layer instead of three.
//
// The 3-barrier kernel (decode_fused.cu) spends a whole phase -- grid barrier, staging, 0.5 us of dot products, ~5.9 us in all --
// on the QKV rows because the attention of a head needs that head's q / k / v, which 148 CTAs computed.  Here a head IS a cluster
// of 4 CTAs (16 clusters = 64 CTAs; B200 co-schedules 33 four-CTA clusters of this kernel, 15 eight-CTA ones: tools/probes/
// cluster_probe.cu): the cluster computes exactly its head's 192 QKV rows (48 per CTA, 3 per warp, prefetched across the grid
// barrier towards L2), exchanges them through distributed shared memory (barrier.cluster ~0.25 us, a DSMEM read ~0.15 us -- against
// ~1.5-2 us for any exchange through L2, which is why the two-barrier variant over L2 flags measured slower,
// profiles/r02_experiments_call1.log), runs the attention of its key quarter, folds the four partials through DSMEM and adds
// its quarter of the out-proj rows into the fixed-point accumulator.  The FFN phase is the fused kernel's (all CTAs).
//   PA  x_l = xmid_{l-1} + FFN2_{l-1}; LN1; [head clusters: QKV rows | cluster sync | attention quarter + KV append | cluster
//       sync | fold + out-proj share]                                                     | grid barrier
//   PB  xmid_l = x_l + out-proj; LN2; FFN1 rows + ReLU; FFN2 share                          | grid barrier
// 26 grid barriers per step (38 in the 3-barrier kernel).  Same arithmetic per element as the fused kernel except that a head's
// keys are split 4 ways instead of 9 (a different but equally valid order of the softmax fold).
#include <cooperative_groups.h>

#include <cstdio>
#include <cstdlib>

#include "decode_fused_common.cuh"

namespace vx {
namespace {

constexpr int kCS = 4;                 // CTAs per cluster = key splits of a head
// dynamic shared memory (floats): sx [1024] | hs [kHS] | scratch (attention: 64 + 16*64; sampling: 1032 + PSORT + PSORT/2 + PNT + 40) | weights
constexpr int kClusterScratchFloats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
constexpr int kClusterScratch0 = kD + kHS + ((kClusterScratchFloats + 3) & ~3);
constexpr int kQRows = kDH / kCS;      // q (k, v) rows of the head computed by one CTA: 16 = one per warp

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
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

// The head CTA's weights of a layer -- its 48 QKV rows (96 KB) and its 256 x 64 out-proj block (32 KB) -- are copied into shared
// memory with cp.async BEFORE the grid barrier that precedes the phase, so their HBM latency hides behind the barrier and the
// LayerNorm staging like the register prefetch of the 3-barrier kernel, without holding 64 registers across the barrier (that
// version spilled, and a spill store waits for its load: +2.7 us per phase).  Each lane later reads exactly the bytes it copied
// itself, so cp.async.wait_group + __syncwarp is all the synchronisation the consumer needs.
constexpr int kWqFloats = 3 * kQRows * kD / 2;          // bf16 [48][1024] counted in floats
constexpr int kWoFloats = (kD / kCS) * kDH / 2;         // bf16 [256][64]
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void head_issue_async(bf16* wq_s, bf16* wo_s, const bf16* __restrict__ Wq, const bf16* __restrict__ Wo, int h, int rank) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const int n = j * kD + h * kDH + rank * kQRows + warp;          // this warp's q / k / v row
#pragma unroll
    for (int it = 0; it < 4; ++it)
      cp_async16(wq_s + (size_t)(j * kQRows + warp) * kD + it * 256 + lane * 8, Wq + (size_t)n * kD + it * 256 + lane * 8);
  }
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int r = pass * 64 + warp * 4 + (lane >> 3);                // row of the CTA's 256 out-proj rows
    cp_async16(wo_s + (size_t)r * kDH + (lane & 7) * 8, Wo + (size_t)(rank * (kD / kCS) + r) * kD + h * kDH + (lane & 7) * 8);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void head_wait_async() {
  asm volatile("cp.async.wait_group 0;" ::: "memory");
  __syncwarp();
}
// -> xchg[j*16 + warp] (j = q, k, v)
__device__ __forceinline__ void qkv_finish(const bf16* wq_s, const float* __restrict__ bias, int h, int rank, const float* __restrict__ sx, float* xchg) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float a[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int it = 0; it < 4; ++it) {
    const float4 x0 = *reinterpret_cast<const float4*>(&sx[it * 256 + lane * 8]);
    const float4 x1 = *reinterpret_cast<const float4*>(&sx[it * 256 + lane * 8 + 4]);
    const float xv[8] = {x0.x, x0.y, x0.z, x0.w, x1.x, x1.y, x1.z, x1.w};
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      const uint4 raw = *reinterpret_cast<const uint4*>(wq_s + (size_t)(j * kQRows + warp) * kD + it * 256 + lane * 8);
      float w[8];
      raw_to_f32<bf16>(&raw, 0, w);
#pragma unroll
      for (int i = 0; i < 8; ++i) a[j] = fmaf(w[i], xv[i], a[j]);
    }
  }
#pragma unroll
  for (int j = 0; j < 3; ++j) {
    const float s = warp_sum(a[j]);
    if (lane == 0) xchg[j * kQRows + warp] = s + __ldg(bias + j * kD + h * kDH + rank * kQRows + warp);
  }
}

// attention of keys [j0, j1) of head h (utterance 0) with q / new k / new v in shared memory (qkvf[0..63 | 64..127 | 128..191]);
// the warp-level unit of the fused kernel (decode_fused_common.cuh attn_unit_fused).  Partial (m, l, o[64]) -> part[0..65].
constexpr int kKPG = 4;    // keys per 8-lane group per round: 16 warps x 4 groups x 8 = 512 keys per round (one round up to context 2048)
__device__ __forceinline__ void attn_quarter(const PParams& p, int layer, int h, int split, const float* qkvf, float* smem, float* part) {
  float* wm = smem;                  // [16]   warp partial max
  float* wl = smem + 16;             // [16]   warp partial sum
  float* wo = smem + 64;             // [16][64] warp partial outputs
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 3, sub = lane & 7;
  const int pos = bt.S[0] + bt.P[0] + ld_cg(bt.T);
  const int n = pos + 1;
  int chunk = (n + kCS - 1) / kCS;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + (size_t)h * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + (size_t)h * p.ctx_cap * kDH;
  uint4 kr[kKPG], vr[kKPG];
  const int wbase = j0 + warp * (4 * kKPG) + g;
#pragma unroll
  for (int i = 0; i < kKPG; ++i) {
    const int j = wbase + 4 * i;
    const bool ok = j < j1 && j != pos;
    kr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
    vr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
  }
  float q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) q[i] = qkvf[sub * 8 + i] * 0.125f;
  uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);   // the new row, bf16-rounded like the cached ones
  if (j1 == n) {   // the quarter that holds the new key (CTA-uniform): it also appends the row to the cache
    auto pk = [](float x, float y) {
      return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y)) << 16);
    };
    const float* kn = qkvf + kDH + sub * 8;
    const float* vn = qkvf + 2 * kDH + sub * 8;
    knew = make_uint4(pk(kn[0], kn[1]), pk(kn[2], kn[3]), pk(kn[4], kn[5]), pk(kn[6], kn[7]));
    vnew = make_uint4(pk(vn[0], vn[1]), pk(vn[2], vn[3]), pk(vn[4], vn[5]), pk(vn[6], vn[7]));
    if (warp == 0 && g == 0) {
      *reinterpret_cast<uint4*>(kb + (size_t)pos * kDH + sub * 8) = knew;
      *reinterpret_cast<uint4*>(vb + (size_t)pos * kDH + sub * 8) = vnew;
    }
  }
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int base = wbase; base - g < j1; base += 64 * kKPG) {   // warp-uniform trip count
    float sc[kKPG];
#pragma unroll
    for (int i = 0; i < kKPG; ++i) {
      const int j = base + 4 * i;
      const uint4 kk = (j == pos) ? knew : kr[i];
      const uint32_t u[4] = {kk.x, kk.y, kk.z, kk.w};
      float d = 0.f;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        d = fmaf(q[2 * e], __uint_as_float(u[e] << 16), d);
        d = fmaf(q[2 * e + 1], __uint_as_float(u[e] & 0xffff0000u), d);
      }
      d += __shfl_xor_sync(0xffffffffu, d, 1);
      d += __shfl_xor_sync(0xffffffffu, d, 2);
      d += __shfl_xor_sync(0xffffffffu, d, 4);
      sc[i] = (j < j1) ? d : -INFINITY;
    }
    uint4 vv[kKPG];
#pragma unroll
    for (int i = 0; i < kKPG; ++i) vv[i] = (base + 4 * i == pos) ? vnew : vr[i];
    if (base - g + 64 * kKPG < j1) {   // next round's chunks
#pragma unroll
      for (int i = 0; i < kKPG; ++i) {
        const int j = base + 64 * kKPG + 4 * i;
        const bool ok = j < j1 && j != pos;
        kr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
        vr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
      }
    }
    float bm = sc[0];
#pragma unroll
    for (int i = 1; i < kKPG; ++i) bm = fmaxf(bm, sc[i]);
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    if (bm == -INFINITY) continue;            // warp-uniform: no key of this round belongs to the warp
    const float mn = fmaxf(m, bm);
    const float scale = expf(m - mn);         // 0 on the first round (m = -inf)
    l *= scale;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= scale;
#pragma unroll
    for (int i = 0; i < kKPG; ++i) {
      const float pr = expf(sc[i] - mn);      // 0 for masked keys
      l += pr;
      const uint32_t u[4] = {vv[i].x, vv[i].y, vv[i].z, vv[i].w};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        acc[2 * e] = fmaf(pr, __uint_as_float(u[e] << 16), acc[2 * e]);
        acc[2 * e + 1] = fmaf(pr, __uint_as_float(u[e] & 0xffff0000u), acc[2 * e + 1]);
      }
    }
    m = mn;
  }
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (g == 0) {
    *reinterpret_cast<float4*>(wo + warp * kDH + sub * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(wo + warp * kDH + sub * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (sub == 0) { wm[warp] = m; wl[warp] = l; }
  }
  __syncthreads();
  if (tid < kDH) {
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < PNW; ++w) M = fmaxf(M, wm[w]);
    float L = 0.f, O = 0.f;
#pragma unroll
    for (int w = 0; w < PNW; ++w) {
      const float wt = (wm[w] == -INFINITY) ? 0.f : expf(wm[w] - M);
      L = fmaf(wt, wl[w], L);
      O = fmaf(wt, wo[w * kDH + tid], O);
    }
    part[2 + tid] = O;
    if (tid == 0) { part[0] = M; part[1] = L; }
  }
}

__global__ void __launch_bounds__(PNT, 1) ar_decode_cluster_kernel(const PParams p) {
  using WT = bf16;
  constexpr int MR = 1;
  extern __shared__ __align__(16) float dsm[];
  float* sx = dsm;                       // [1024]
  float* hs = dsm + kD;                  // [kHS] FFN hidden units of this CTA
  float* scratch = hs + kHS;             // attention / sampling / LN-reduction scratch
  bf16* wq_s = reinterpret_cast<bf16*>(dsm + kClusterScratch0);            // [48][1024] this CTA's QKV rows of the layer (head CTAs)
  bf16* wo_s = reinterpret_cast<bf16*>(dsm + kClusterScratch0 + kWqFloats); // [256][64]  its out-proj block
  __shared__ __align__(16) float xchg[3 * kQRows];   // this CTA's q | k | v rows of the head (read by the cluster through DSMEM)
  __shared__ __align__(16) float qkvf[3 * kDH];      // the head's full q | new k | new v
  __shared__ __align__(16) float part[kDH + 8];      // this CTA's attention partial (m, l, o[64]) (read by the cluster through DSMEM)
  __shared__ __align__(16) float sq[kDH];            // folded, normalised head output
  const Batch& bt = p.bt;
  const int B = 1;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  unsigned int epoch = 0;

  int o0, o1, f0, f1, pr0, pr1;
  slice(kD, o0, o1);            // residual-stream columns this CTA writes back / accumulator columns it re-zeroes
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int rank = (int)cluster_ctarank();
  const int cid = blockIdx.x / kCS;
  const bool head = cid < kH;            // cluster c < 16 = head c; the other clusters only take part in the FFN / predict phases
  const int u_h = head ? cid : 0;
  const int wr0 = rank * (kD / kCS);     // this CTA's 256 out-proj rows
  long long* const accO2 = p.acc_out;   // [2][1024]
  long long* const accF2 = p.acc_ffn;   // [2][1024]
  const size_t accN = kD;

  WPre<WT, MR> pre;
  LnPre lp;
  if (head) head_issue_async(wq_s, wo_s, reinterpret_cast<const WT*>(p.layer[0].qkv_w), reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, rank);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
  for (int i = tid; i < kHS; i += PNT) hs[i] = 0.f;

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed (x_0 -> xa) ---------------------------------------------------------------------
    if (blockIdx.x == 0 && !ld_cg(bt.finished) && !(step == 0 && p.resume && !ld_cg(p.pending))) sample_utterance(p, 0, scratch);
    zero_slice(accF2 + accN, B, o0, o1);   // acc_ffn[1] was consumed by the previous step's PF
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (p.stop_below > 0 && ld_cg(bt.n_active) <= p.stop_below) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2T = reinterpret_cast<const WT*>(L.ff2_t);
      // ---- PA: x_l = xmid_{l-1} + FFN2_{l-1} (+bias); LN1; head clusters: QKV rows, attention quarter, out-proj share --------
      if (head && tid == 0) {
        const int n = bt.S[0] + bt.P[0] + ld_cg(bt.T) + 1;
        int chunk = (n + kCS - 1) / kCS;
        chunk = (chunk + 31) & ~31;
        const int j0 = rank * chunk, j1 = min(n - 1, j0 + chunk);
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
        zero_slice(accO_prev, B, o0, o1);   // consumed by PB of layer l-1
      }
      PSTAMP(3);
      if (head) {
        head_wait_async();                 // this lane's copies of the layer's weights have landed (issued before the grid barrier)
        qkv_finish(wq_s, L.qkv_b, u_h, rank, sx, xchg);
        cluster_sync_all();                // every rank's q | k | v rows are in its xchg
        if (tid < 3 * kDH) {
          const int j = tid / kDH, d = tid % kDH;
          qkvf[tid] = dsmem_ld(xchg + j * kQRows + (d % kQRows), (uint32_t)(d / kQRows));
        }
        __syncthreads();
        PSTAMP(4);
        attn_quarter(p, l, u_h, rank, qkvf, scratch, part);
        cluster_sync_all();                // every rank's partial is in its part[]
        if (tid < kDH) {
          float mv[kCS], lv[kCS], ov[kCS];
#pragma unroll
          for (int s = 0; s < kCS; ++s) {
            mv[s] = dsmem_ld(part, (uint32_t)s);
            lv[s] = dsmem_ld(part + 1, (uint32_t)s);
            ov[s] = dsmem_ld(part + 2 + tid, (uint32_t)s);
          }
          float M = -INFINITY;
#pragma unroll
          for (int s = 0; s < kCS; ++s) M = fmaxf(M, mv[s]);
          float Ls = 0.f, Os = 0.f;
#pragma unroll
          for (int s = 0; s < kCS; ++s) {
            const float w = (mv[s] == -INFINITY) ? 0.f : expf(mv[s] - M);
            Ls += w * lv[s];
            Os += w * ov[s];
          }
          sq[tid] = Os / Ls;
        }
        __syncthreads();
        PSTAMP(5);
        // out-proj share: rows [wr0, wr0 + 256) x this head's 64 columns
        {
          const int grp = lane >> 3, sub = lane & 7;
          float xv[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) xv[i] = sq[sub * 8 + i];
#pragma unroll
          for (int pass = 0; pass < 4; ++pass) {
            const int nrow = wr0 + pass * 64 + warp * 4 + grp;
            const uint4 raw = *reinterpret_cast<const uint4*>(wo_s + (size_t)(pass * 64 + warp * 4 + grp) * kDH + sub * 8);
            float w[8];
            raw_to_f32<bf16>(&raw, 0, w);
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 8; ++i) s = fmaf(w[i], xv[i], s);
            s += __shfl_xor_sync(0xffffffffu, s, 1);
            s += __shfl_xor_sync(0xffffffffu, s, 2);
            s += __shfl_xor_sync(0xffffffffu, s, 4);
            if (sub == 0) fix_add(accO + nrow, s);
          }
        }
        gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);   // FFN1 rows land while the CTA waits at the grid barrier
      } else {
        gemv_issue<WT, MR>(pre, W1, kD, f0, f1, L.ff1_b, nullptr);
      }
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      if (tid == 0) l2_prefetch(W2T + (size_t)f0 * kD, (size_t)(f1 - f0) * kD * sizeof(WT));
      PSTAMP(6);
      grid_barrier(p.bar, epoch, prof);
      // ---- PB: xmid_l = x_l + out-proj (+bias); LN2; FFN1 rows + ReLU; FFN2 share --------------------------------------
      if (l > 0) zero_slice(accF_prev, B, o0, o1);   // consumed by PA of this layer
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      PSTAMP(3);
      // the next layer's head weights go out now (their shared-memory buffers were last read in PA): issued after ffn2_share they
      // queued behind its 1024 atomics and made the head CTAs the stragglers of this phase (2.7 vs 1.5 us)
      if (head && l + 1 < kL)
        head_issue_async(wq_s, wo_s, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), reinterpret_cast<const WT*>(p.layer[l + 1].out_w), u_h, rank);
      gemv_finish<WT, MR>(pre, W1, kD, f0, f1, sx, B, L.ff1_b, nullptr, [&](int, int) { return 0.f; },
                          [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();
      PSTAMP(4);
      ffn2_share<MR>(W2T, f0, f1, hs, B, accF);
      if (l + 1 < kL) {
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
    zero_slice(accO2 + accN, B, o0, o1);   // acc_out[1] was consumed by PB of layer 11
    gemv_finish<WT, MR>(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1, sx, B, nullptr, nullptr,
                        [&](int, int) { return 0.f; },
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    if (head) head_issue_async(wq_s, wo_s, reinterpret_cast<const WT*>(p.layer[0].qkv_w), reinterpret_cast<const WT*>(p.layer[0].out_w), u_h, rank);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
  asm volatile("cp.async.wait_all;" ::: "memory");
  cluster_sync_all();   // no CTA of a cluster leaves while a peer may still read its shared memory
}

}  // namespace

// Number of CTAs the cluster kernel should be launched with on this device (0: the 16 head clusters cannot be co-resident).
static size_t cluster_smem_bytes() { return ((size_t)kClusterScratch0 + kWqFloats + kWoFloats) * sizeof(float); }

static cudaError_t cluster_config(int num_sms, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attrs, int* nclusters_out) {
  const size_t smem = cluster_smem_bytes();
  cudaError_t e = cudaFuncSetAttribute(ar_decode_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  cfg = cudaLaunchConfig_t{};
  cfg.blockDim = dim3(PNT);
  cfg.dynamicSmemBytes = smem;
  attrs[0].id = cudaLaunchAttributeClusterDimension;
  attrs[0].val.clusterDim.x = kCS; attrs[0].val.clusterDim.y = 1; attrs[0].val.clusterDim.z = 1;
  cfg.attrs = attrs;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(kCS * kH);
  int nclusters = 0;
  e = cudaOccupancyMaxActiveClusters(&nclusters, ar_decode_cluster_kernel, &cfg);
  if (e != cudaSuccess) return e;
  nclusters = std::min(nclusters, num_sms / kCS);
  // every CTA spins on the grid barrier: all of them must be resident at once.  The kernel is launched with exactly the number of
  // clusters the device can hold (33 on a B200) and needs the 16 head clusters plus enough CTAs for <= kHS hidden units each.
  if (nclusters < kH || (kFF + nclusters * kCS - 1) / (nclusters * kCS) > kHS) return cudaErrorLaunchOutOfResources;
  cfg.gridDim = dim3(kCS * nclusters);
  *nclusters_out = nclusters;
  return cudaSuccess;
}

// CTAs the cluster kernel runs with on the current device, 0 if its 16 head clusters cannot be co-resident (engine.cu then keeps
// the 3-barrier kernel)
int decode_cluster_grid(int num_sms) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[1];
  int n = 0;
  if (cluster_config(num_sms, cfg, attrs, &n) != cudaSuccess) { cudaGetLastError(); return 0; }
  return n * kCS;
}

cudaError_t launch_decode_cluster(const PParams& p, int num_sms, cudaStream_t st, int* grid_out) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attrs[1];
  int n = 0;
  cudaError_t e = cluster_config(num_sms, cfg, attrs, &n);
  if (e != cudaSuccess) return e;
  cfg.stream = st;
  if (grid_out) *grid_out = kCS * n;
  // Every CTA spins on the grid barrier, so all of them must be resident at once.  A cooperative launch makes the runtime check that
  // (and fail instead of hanging when the GPU is shared); where the driver does not combine the cooperative attribute with cluster
  // dimensions the occupancy query above is the guarantee (an otherwise idle device holds exactly `n` clusters of this kernel).
  static int coop_ok = -1;   // -1 unknown, 0 rejected once, 1 accepted
  if (coop_ok != 0) {
    cudaLaunchAttribute a2[2];
    a2[0] = attrs[0];
    a2[1].id = cudaLaunchAttributeCooperative;
    a2[1].val.cooperative = 1;
    cudaLaunchConfig_t c2 = cfg;
    c2.attrs = a2;
    c2.numAttrs = 2;
    e = cudaLaunchKernelEx(&c2, ar_decode_cluster_kernel, p);
    if (e == cudaSuccess) {
      if (coop_ok < 0 && getenv("VX_DECODE_PROF")) fprintf(stderr, "ar_decode_cluster_kernel: cooperative + cluster launch accepted (%d CTAs)\n", kCS * n);
      coop_ok = 1;
      return e;
    }
    if (coop_ok == 1) return e;      // it worked before: a real failure
    if (getenv("VX_DECODE_PROF")) fprintf(stderr, "ar_decode_cluster_kernel: cooperative + cluster launch rejected (%s), plain cluster launch\n", cudaGetErrorString(e));
    cudaGetLastError();
    coop_ok = 0;
  }
  return cudaLaunchKernelEx(&cfg, ar_decode_cluster_kernel, p);
}

}  // namespace vx
