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
template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_x(float* sx, int B, Val value, const LnPre& lp, float* red, float* xout, int c0, int c1) {
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
__device__ void attn_unit_fused(const PParams& p, int layer, int b, int h, int split, float* smem, int& att_epoch,
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
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
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
