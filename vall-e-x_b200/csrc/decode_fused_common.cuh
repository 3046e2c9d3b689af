// Device code of the fused decode kernel (decode_fused.cu): fixed-point accumulators, staging with residual
// write-back, the out-proj share prefetch, the attention unit with head rendezvous, the FFN2 share.
#pragma once
#include "decode_common.cuh"

namespace vx {
namespace {
// =================================================================================================================
// Fused variant (bf16 weights, B <= 8, VX_DECODE_FUSED=1): 3 grid barriers per layer instead of 5 (38 per step).
//   P1  LN1 + QKV rows                                                      | grid barrier
//   P2  attention unit (b, head, key split); the splits of a head meet on a ticket (a 9-CTA rendezvous, not a
//       grid barrier), every one folds the partials and multiplies the head's 64 outputs by ITS rows of the
//       out-proj head slice, adding into a fixed-point accumulator          | grid barrier
//   P3  residual + LN2 + FFN1 rows (hidden units [f0,f1)) + ReLU, then at once those hidden units' share of FFN2
//       (rows of the TRANSPOSED W2) added into a fixed-point accumulator    | grid barrier
// Cross-CTA sums go through 64-bit integer atomics on 2^-30 fixed point: integer addition is associative, so the
// result is independent of arrival order (deterministic run to run) and carries more precision than an fp32 sum.
// The consumer of an accumulator adds it to the residual stream while staging; accumulators are double-buffered by
// layer parity and re-zeroed one phase after their consumption (always a barrier before their next use).
// =================================================================================================================
constexpr float kFixScale = 1073741824.f;  // 2^30
constexpr int kHS = 64;                    // max hidden units per CTA (grid >= 64)
__device__ __forceinline__ void fix_add(long long* p, float v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(p), (unsigned long long)__float2ll_rn(v * kFixScale));
}
__device__ __forceinline__ float fix_get(const long long* p) { return __ll2float_rn(__ldcg(p)) * (1.0f / kFixScale); }

// stage_ln + write-back: the staged residual value (before normalisation) of columns [c0,c1) goes to xout
// `after_loads()` runs once the activation loads have been issued (the cluster kernel requests its weight rows there: they land
// under the two block reductions without delaying the loads the reductions wait for)
template <int MR, typename Val, typename Mid>
__device__ __forceinline__ void stage_ln_x(float* sx, int B, Val value, const LnPre& lp, float* red, float* xout, int c0, int c1, Mid after_loads) {
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
    if (mb == 0) after_loads();
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
      const float s = warp_sum(v0[m] + v1[m]);
      if (lane == 0) red[warp * 8 + m] = s;
    }
    __syncthreads();
    float mean[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      float s = 0.f;
#pragma unroll
      for (int w = 0; w < PNW; ++w) s += red[w * 8 + m];
      mean[m] = s / (float)kD;
    }
#pragma unroll
    for (int m = 0; m < R; ++m) {
      const float d0 = v0[m] - mean[m], d1 = v1[m] - mean[m];
      const float s = warp_sum(d0 * d0 + d1 * d1);
      if (lane == 0) red[PNW * 8 + warp * 8 + m] = s;
    }
    __syncthreads();
#pragma unroll
    for (int m = 0; m < R; ++m) {
      if (m < nb) {
        float s = 0.f;
#pragma unroll
        for (int w = 0; w < PNW; ++w) s += red[PNW * 8 + w * 8 + m];
        const float rstd = 1.0f / sqrtf(s / (float)kD + 1e-5f);
        sx[(mb + m) * kD + tid] = __fadd_rn(__fmul_rn((v0[m] - mean[m]) * rstd, lp.g0), lp.b0);
        sx[(mb + m) * kD + tid + PNT] = __fadd_rn(__fmul_rn((v1[m] - mean[m]) * rstd, lp.g1), lp.b1);
      }
    }
    __syncthreads();
  }
}

template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_x(float* sx, int B, Val value, const LnPre& lp, float* red, float* xout, int c0, int c1) {
  stage_ln_x<MR>(sx, B, value, lp, red, xout, c0, c1, [] {});
}

__device__ __forceinline__ void zero_slice(long long* acc, int B, int c0, int c1) {
  const int w = c1 - c0;
  for (int i = threadIdx.x; i < B * w; i += PNT) acc[(size_t)(i / w) * kD + c0 + (i % w)] = 0;
}

struct WoPre { uint4 r[2]; };  // this thread's 16 bytes of the first two 64-row passes over the unit's out-proj rows
// rows [r0,r1) x columns [h*64, h*64+64) of Wo: 4 rows per warp (8 lanes x 8 elements each), 64 rows per pass
__device__ __forceinline__ void wo_issue(WoPre& wp, const bf16* __restrict__ Wo, int h, int r0, int r1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
#pragma unroll
  for (int pass = 0; pass < 2; ++pass) {
    const int n = r0 + pass * 64 + warp * 4 + (lane >> 3);
    wp.r[pass] = (n < r1) ? ldw_raw(Wo + (size_t)n * kD + h * kDH + (lane & 7) * 8) : make_uint4(0, 0, 0, 0);
  }
}

// attention unit + head rendezvous + out-proj share.  `att_epoch`: tickets this unit's head has seen so far.
// The unit's keys [j0, j1) are spread over the 16 warps in runs of 16 (4 key rows per load instruction: 8 lanes x 16 B = one
// 128-byte row, 512 contiguous bytes per instruction); a warp requests ALL of its K and V chunks of a 256-key round at once
// (one memory round trip), keeps a running (max, sum, o[64]) over its rounds, and the 16 warp partials are folded through
// shared memory in fixed order.  (The round-1 unit took one key per thread for Q.K, two block reductions, then 2-3 dependent
// V loads per 8-lane group and a 64-deep serial fold: ~3 more dependent L2 round trips and 5 more block barriers per layer.)
__device__ void attn_unit_fused(const PParams& p, int layer, int b, int h, int split, float* smem, int& att_epoch,
                                const WoPre& wp, const bf16* __restrict__ Wo, int r0, int r1, long long* accO) {
  float* sq = smem;                  // [64]   folded head output
  float* wm = smem + 64;             // [16]   warp partial max
  float* wl = smem + 80;             // [16]   warp partial sum
  float* wo = smem + 128;            // [16][64] warp partial outputs
  const Batch& bt = p.bt;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 3, sub = lane & 7;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH + sub * 8;
  bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
  // ---- first round's K / V chunks: requested before anything else -------------------------------------------------------
  uint4 kr[4], vr[4];
  const int wbase = j0 + warp * 16 + g;     // this lane group's first key of round 0
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int j = wbase + 4 * i;
    const bool ok = j < j1 && j != pos;
    kr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
    vr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
  }
  float q[8];
  {
    const float4 a = __ldcg(reinterpret_cast<const float4*>(row)), c = __ldcg(reinterpret_cast<const float4*>(row) + 1);
    q[0] = a.x * 0.125f; q[1] = a.y * 0.125f; q[2] = a.z * 0.125f; q[3] = a.w * 0.125f;
    q[4] = c.x * 0.125f; q[5] = c.y * 0.125f; q[6] = c.z * 0.125f; q[7] = c.w * 0.125f;
  }
  uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);   // the new row, bf16-rounded like the cached ones
  if (j1 == n) {   // the split that holds the new key (warp-uniform): it also appends the row to the cache
    const float4 k0 = __ldcg(reinterpret_cast<const float4*>(row + kD)), k1 = __ldcg(reinterpret_cast<const float4*>(row + kD) + 1);
    const float4 v0 = __ldcg(reinterpret_cast<const float4*>(row + 2 * kD)), v1 = __ldcg(reinterpret_cast<const float4*>(row + 2 * kD) + 1);
    auto pk = [](float x, float y) {
      return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(x)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(y)) << 16);
    };
    knew = make_uint4(pk(k0.x, k0.y), pk(k0.z, k0.w), pk(k1.x, k1.y), pk(k1.z, k1.w));
    vnew = make_uint4(pk(v0.x, v0.y), pk(v0.z, v0.w), pk(v1.x, v1.y), pk(v1.z, v1.w));
    if (warp == 0 && g == 0) {
      *reinterpret_cast<uint4*>(kb + (size_t)pos * kDH + sub * 8) = knew;
      *reinterpret_cast<uint4*>(vb + (size_t)pos * kDH + sub * 8) = vnew;
    }
  }
  float m = -INFINITY, l = 0.f;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int base = wbase; base - g < j1; base += 256) {   // warp-uniform trip count (base - g is the warp's first key of the round)
    float sc[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
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
    uint4 vv[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) vv[i] = (base + 4 * i == pos) ? vnew : vr[i];
    if (base - g + 256 < j1) {   // next round's chunks (contexts beyond 256 keys per unit)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int j = base + 256 + 4 * i;
        const bool ok = j < j1 && j != pos;
        kr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
        vr[i] = ok ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + sub * 8)) : make_uint4(0, 0, 0, 0);
      }
    }
    float bm = fmaxf(fmaxf(sc[0], sc[1]), fmaxf(sc[2], sc[3]));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
    bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
    if (bm == -INFINITY) continue;            // warp-uniform: no key of this round belongs to the warp
    const float mn = fmaxf(m, bm);
    const float scale = expf(m - mn);         // 0 on the first round (m = -inf)
    l *= scale;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] *= scale;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
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
  // fold the 4 key groups of the warp, then the 16 warps (fixed order)
  l += __shfl_xor_sync(0xffffffffu, l, 8);
  l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
    acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
  }
  if (g == 0) {   // (the previous use of this scratch is behind at least one grid barrier)
    *reinterpret_cast<float4*>(wo + warp * kDH + sub * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    *reinterpret_cast<float4*>(wo + warp * kDH + sub * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
    if (sub == 0) { wm[warp] = m; wl[warp] = l; }
  }
  __syncthreads();
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
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
    pp[2 + tid] = O;
    if (tid == 0) { pp[0] = M; pp[1] = L; }
  }
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
  // ---- out-proj share: rows [r0,r1) x this head's 64 columns --------------------------------------------------------
  const int grp = lane >> 3;
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

// hidden units [f0,f1) of this CTA (values in hs[m*kHS + j-f0]) times their rows of W2^T [4096][1024] -> accF[m][0..1023]
template <int MR>
__device__ __forceinline__ void ffn2_share(const bf16* __restrict__ W2T, int f0, int f1, const float* __restrict__ hs, int B,
                                           long long* accF) {
  const int tid = threadIdx.x;
  float acc[MR][2];
#pragma unroll
  for (int m = 0; m < MR; ++m) { acc[m][0] = 0.f; acc[m][1] = 0.f; }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(W2T + (size_t)f0 * kD) + tid;  // columns 2*tid, 2*tid+1
  for (int j0 = 0; j0 < f1 - f0; j0 += 16) {
    uint32_t w[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      w[i] = 0;
      if (j0 + i < f1 - f0) asm volatile("ld.global.nc.L1::no_allocate.u32 %0, [%1];" : "=r"(w[i]) : "l"(base + (size_t)(j0 + i) * (kD / 2)));
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float wa = __uint_as_float(w[i] << 16), wb = __uint_as_float(w[i] & 0xffff0000u);
#pragma unroll
      for (int m = 0; m < MR; ++m) {
        const float hv = hs[m * kHS + j0 + i];   // zero beyond f1-f0 (cleared by the caller)
        acc[m][0] = fmaf(wa, hv, acc[m][0]);
        acc[m][1] = fmaf(wb, hv, acc[m][1]);
      }
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



}  // namespace
}  // namespace vx
