// Tensor-core batched decode kernel (5 <= B <= 32, bf16): the five GEMV phases of the 62-barrier loop on mma.sync, row-per-warp
// activation staging, FFN2 split-K fold by ticket, balanced warp-level attention over the flattened key space.
#include "decode_common.cuh"

namespace vx {
namespace {
// =================================================================================================================
// Batched variant (bf16 weights, 8 <= B <= 32, VX_DECODE_MMA=1, experimental): the five GEMV phases of the 62-barrier
// kernel on tensor cores.  Weights are the M operand of mma.sync.m16n8k16 (a CTA's <= 32-row slice = two m-tiles),
// the batch is N, activations are split into bf16 hi + lo halves (two MMAs, fp32 accumulate: fp32-grade activations).
// The k index inside a 32-element chunk is permuted identically for A and B, so every lane loads 16 contiguous bytes
// of a weight row (64-byte runs per row) and its B fragments are 16-byte shared-memory loads from a fragment-ordered
// image written by the staging step.  16 warps = 2 m-tiles x 8 k-groups; the partial tiles are reduced through shared
// memory in fixed order.  (The CUDA-core path re-reads B x 4 KB of activations per weight row and spills at MR >= 8.)
// =================================================================================================================
struct MPre { uint4 ra[4], rb[4]; };   // weight rows (g, g+8) of this warp's m-tile: 4 chunks of 32 k of its k-group

__device__ __forceinline__ void mma_16816(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

// Fragment image of the activations (hi and lo separately): [k-group 8][chunk 4][n-tile MR/8][lane 32] x 16 B, the 16 B of
// lane (g, q) = batch row nt*8+g, columns kg*128 + ch*32 + q*8 .. +7 in order (words 0,1 = (b0,b1) of the chunk's first
// m16n8k16, words 2,3 = of its second).
template <int MR>
__device__ __forceinline__ void frag_store(uint16_t* fh, uint16_t* fl, int m, int c, float v) {
  constexpr int NT = MR / 8;
  const int kg = c >> 7, ch = (c >> 5) & 3, q = (c >> 3) & 3, e = c & 7;
  const size_t idx = (((((size_t)kg * 4 + ch) * NT + (m >> 3)) * 32) + (m & 7) * 4 + q) * 8 + e;
  const __nv_bfloat16 h = __float2bfloat16_rn(v);
  const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
  fh[idx] = __bfloat16_as_ushort(h);
  fl[idx] = __bfloat16_as_ushort(l);
}

// Row-per-warp staging: warp w stages batch rows w, w + 16 (MR = 32); a lane holds columns it*128 + lane*4 .. +3 of a row
// (8 coalesced 16-byte loads, all rows' loads in flight at once), LayerNorm statistics are warp shuffles (no block-wide
// reduction), and the normalised row goes straight into the hi/lo fragment images.  src(m) -> row pointer or nullptr (zeros).
// LayerNorm vectors of the NEXT LN phase: 8 KB copied into shared memory with cp.async BEFORE the grid barrier that precedes
// the phase (they are evicted from L2 by the 3 GB a step streams, so a plain load would put an HBM round trip behind the
// row statistics); stage_rows_frag<LN> waits for the copy.
__device__ __forceinline__ void ln_copy_async(float* lnp, const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int t = threadIdx.x;   // 512 threads x 16 B = gamma[1024] | beta[1024]
  const float* srcp = (t < PNT / 2) ? gamma + t * 4 : beta + (t - PNT / 2) * 4;
  const uint32_t dst = (uint32_t)__cvta_generic_to_shared(lnp + t * 4);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(srcp) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}

template <int MR, bool LN, typename Src>
__device__ __forceinline__ void stage_rows_frag(uint16_t* fh, uint16_t* fl, int B, Src src, const float* lnp) {
  const float* gamma = lnp;
  const float* beta = lnp + kD;
  constexpr int NT = MR / 8;
  constexpr int RPW = (MR + PNW - 1) / PNW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if constexpr (LN) {
    asm volatile("cp.async.wait_all;" ::: "memory");
    __syncthreads();   // every thread's 16 bytes of gamma | beta have landed
  }
  // one row at a time (a row = 32 registers per lane: with the prefetched weight registers live, two rows in flight spill,
  // and a spilled load serialises every following one behind it)
#pragma unroll 1
  for (int r = 0; r < RPW; ++r) {
    const int m = warp + r * PNW;
    if (m >= MR) break;   // warp-uniform (MR = 8: warps 8..15 have no row)
    const float* row = (m < B) ? src(m) : nullptr;
    float4 v[8];
#pragma unroll
    for (int it = 0; it < 8; ++it)
      v[it] = row ? __ldcg(reinterpret_cast<const float4*>(row + it * 128 + lane * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    float mean = 0.f, rstd = 1.f;
    if constexpr (LN) {
      float s = 0.f;
#pragma unroll
      for (int it = 0; it < 8; ++it) s += (v[it].x + v[it].y) + (v[it].z + v[it].w);
      mean = warp_sum(s) / (float)kD;
      float q = 0.f;
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const float d0 = v[it].x - mean, d1 = v[it].y - mean, d2 = v[it].z - mean, d3 = v[it].w - mean;
        q += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
      }
      rstd = 1.0f / sqrtf(warp_sum(q) / (float)kD + 1e-5f);
    }
#pragma unroll
    for (int it = 0; it < 8; ++it) {
      float x[4] = {v[it].x, v[it].y, v[it].z, v[it].w};
      if constexpr (LN) {
        const float4 g = *reinterpret_cast<const float4*>(gamma + it * 128 + lane * 4);
        const float4 bb = *reinterpret_cast<const float4*>(beta + it * 128 + lane * 4);
        x[0] = __fadd_rn(__fmul_rn((x[0] - mean) * rstd, g.x), bb.x);
        x[1] = __fadd_rn(__fmul_rn((x[1] - mean) * rstd, g.y), bb.y);
        x[2] = __fadd_rn(__fmul_rn((x[2] - mean) * rstd, g.z), bb.z);
        x[3] = __fadd_rn(__fmul_rn((x[3] - mean) * rstd, g.w), bb.w);
      }
      // fragment image slot of (m, c = it*128 + lane*4 .. +3): kg = it, ch = lane >> 3, q = (lane >> 1) & 3, e = (lane & 1) * 4
      const size_t idx = (((((size_t)it * 4 + (lane >> 3)) * NT + (m >> 3)) * 32) + (m & 7) * 4 + ((lane >> 1) & 3)) * 8 + (lane & 1) * 4;
      uint32_t hw[2], lw[2];
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        const __nv_bfloat16 h0 = __float2bfloat16_rn(x[2 * e]), h1 = __float2bfloat16_rn(x[2 * e + 1]);
        const __nv_bfloat16 l0 = __float2bfloat16_rn(x[2 * e] - __bfloat162float(h0));
        const __nv_bfloat16 l1 = __float2bfloat16_rn(x[2 * e + 1] - __bfloat162float(h1));
        hw[e] = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
        lw[e] = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
      }
      *reinterpret_cast<uint2*>(fh + idx) = make_uint2(hw[0], hw[1]);
      *reinterpret_cast<uint2*>(fl + idx) = make_uint2(lw[0], lw[1]);
    }
  }
  __syncthreads();
}

// request this warp's share of rows [r0,r1) (<= 32 rows) of W [.][ldw], k columns [0,1024)
__device__ __forceinline__ void gemv_issue_mma(MPre& pre, const bf16* __restrict__ W, size_t ldw, int r0, int r1) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mt = warp & 1, kg = warp >> 1, g = lane >> 2, q = lane & 3;
  const int ra = r0 + mt * 16 + g, rb = ra + 8;
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
    const int k = kg * 128 + ch * 32 + q * 8;
    pre.ra[ch] = (ra < r1) ? ldw_raw(W + (size_t)ra * ldw + k) : make_uint4(0, 0, 0, 0);
    pre.rb[ch] = (rb < r1) ? ldw_raw(W + (size_t)rb * ldw + k) : make_uint4(0, 0, 0, 0);
  }
}

// rows [r0,r1) x batch: tensor-core partial tiles, fixed-order reduction over the 8 k-groups, then
// emit(n, m, dot + addend(n, m) + bias[n] + bias2[n]).  red: [8][32][MR + 1] floats (odd stride: the transposed read is
// conflict-free).  The addends are requested BEFORE the tensor-core loop (they do not depend on it), and consecutive threads
// emit consecutive rows n of one batch entry, so the global stores are 128-byte runs.
template <int MR, typename Add, typename Emit>
__device__ __forceinline__ void gemv_finish_mma(const MPre& pre, const uint16_t* fh, const uint16_t* fl, int r0, int r1, int B,
                                                const float* __restrict__ bias, const float* __restrict__ bias2, float* red,
                                                Add addend, Emit emit) {
  constexpr int NT = MR / 8;
  constexpr int RS = MR + 1;
  constexpr int NOUT = (32 * MR + PNT - 1) / PNT;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = warp & 1, kg = warp >> 1, g = lane >> 2, q = lane & 3;
  float ad[NOUT];
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const int t = tid + i * PNT, row = t & 31, m = t >> 5, n = r0 + row;
    ad[i] = (m < B && n < r1) ? (addend(n, m) + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f)) : 0.f;
  }
  float c[NT][4];
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) { c[nt][0] = 0.f; c[nt][1] = 0.f; c[nt][2] = 0.f; c[nt][3] = 0.f; }
  const uint4* H = reinterpret_cast<const uint4*>(fh);
  const uint4* Lo = reinterpret_cast<const uint4*>(fl);
#pragma unroll
  for (int ch = 0; ch < 4; ++ch) {
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int idx = ((kg * 4 + ch) * NT + nt) * 32 + lane;
      const uint4 xh = H[idx], xl = Lo[idx];
      mma_16816(c[nt], pre.ra[ch].x, pre.rb[ch].x, pre.ra[ch].y, pre.rb[ch].y, xh.x, xh.y);
      mma_16816(c[nt], pre.ra[ch].x, pre.rb[ch].x, pre.ra[ch].y, pre.rb[ch].y, xl.x, xl.y);
      mma_16816(c[nt], pre.ra[ch].z, pre.rb[ch].z, pre.ra[ch].w, pre.rb[ch].w, xh.z, xh.w);
      mma_16816(c[nt], pre.ra[ch].z, pre.rb[ch].z, pre.ra[ch].w, pre.rb[ch].w, xl.z, xl.w);
    }
  }
#pragma unroll
  for (int nt = 0; nt < NT; ++nt) {
    const int row = mt * 16 + g, n = nt * 8 + q * 2;
    float* d0 = red + ((size_t)kg * 32 + row) * RS + n;
    float* d1 = red + ((size_t)kg * 32 + row + 8) * RS + n;
    d0[0] = c[nt][0]; d0[1] = c[nt][1];
    d1[0] = c[nt][2]; d1[1] = c[nt][3];
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < NOUT; ++i) {
    const int t = tid + i * PNT, row = t & 31, m = t >> 5, n = r0 + row;
    if (m < B && n < r1) {
      float s = 0.f;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) s += red[((size_t)k8 * 32 + row) * RS + m];
      emit(n, m, s + ad[i]);
    }
  }
  __syncthreads();
}

// =================================================================================================================
// Balanced warp-level attention phase of the batched kernel (modules/activation.py:142-167 for one new query row).
// The keys of all (utterance, head) pairs form one flattened space of 32-key blocks; every warp of the grid takes an
// equal contiguous range of it, so the KV cache streams once at HBM rate whatever the mix of context lengths (the
// CTA-per-head units before left 148 CTAs with 3.5 waves of unequal units).  A warp walks its range with 4 key rows per
// load instruction (8 lanes x 16 B = one 128-byte row; 512 contiguous bytes per instruction), keeps up to 16 x 16 B per
// lane in flight (K of block n+1 and V of block n are requested before block n's arithmetic), carries a running
// (max, sum, o[64]) per (utterance, head) segment and writes ONE partial record per segment.  The last segment of a head
// to arrive (ticket) folds the records in piece order -- deterministic whoever is last.
// =================================================================================================================
constexpr int kRec = kDH + 4;       // partial record: m, l, pad, pad, o[64]  (16-byte aligned rows)
constexpr int kPerMin = 2;          // at least 64 keys per warp (bounds the pieces per head when B is small)

struct AttnPlan {                   // per step, in registers: lane b owns utterance b
  int n, nblk, pb;                  // keys incl. the new one (0: finished / no such utterance), 32-key blocks, exclusive prefix of 16*nblk
  int tb, per;                      // total blocks, blocks per warp (warp-uniform)
};
__device__ __forceinline__ AttnPlan attn_plan(const Batch& bt) {
  const int lane = threadIdx.x & 31;
  AttnPlan pl;
  pl.n = 0;
  if (lane < bt.B && !ld_cg(bt.finished + lane)) pl.n = bt.S[lane] + bt.P[lane] + ld_cg(bt.T + lane) + 1;
  pl.nblk = (pl.n + 31) >> 5;
  int incl = kH * pl.nblk;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  pl.pb = incl - kH * pl.nblk;
  pl.tb = __shfl_sync(0xffffffffu, incl, 31);
  const int nw = gridDim.x * PNW;
  pl.per = max(kPerMin, (pl.tb + nw - 1) / nw);
  return pl;
}

__device__ __forceinline__ void unpack8(const uint4& r, float (&w)[8]) {
  const uint32_t u[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) { w[2 * i] = __uint_as_float(u[i] << 16); w[2 * i + 1] = __uint_as_float(u[i] & 0xffff0000u); }
}
__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  return (uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(a)) | ((uint32_t)__bfloat16_as_ushort(__float2bfloat16_rn(b)) << 16);
}

__device__ void attn_phase_balanced(const PParams& p, int layer, const AttnPlan& pl) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PNW + warp;
  const int g = lane >> 3, sub = lane & 7;
  if ((long long)gw * pl.per >= pl.tb) return;
  int gb = gw * pl.per;
  const int gb1 = min(pl.tb, gb + pl.per);
  while (gb < gb1) {
    // utterance of flattened block gb: the last one whose range starts at or before it
    const unsigned cand = __ballot_sync(0xffffffffu, pl.nblk > 0 && pl.pb <= gb);
    const int b = 31 - __clz(cand);
    const int nb = __shfl_sync(0xffffffffu, pl.nblk, b), pbb = __shfl_sync(0xffffffffu, pl.pb, b), n = __shfl_sync(0xffffffffu, pl.n, b);
    const int r = gb - pbb, h = r / nb, blk0 = r - h * nb;
    const int cnt = min(gb1 - gb, nb - blk0);
    const int s_first = pbb + h * nb;
    const int w_first = s_first / pl.per, npieces = (s_first + nb - 1) / pl.per - w_first + 1;
    const int piece = gw - w_first;
    const int ja = blk0 * 32, jb = min(n, (blk0 + cnt) * 32), pos = n - 1;
    gb += cnt;

    const float* qrow = p.qkv + (size_t)b * (3 * kD) + h * kDH + sub * 8;
    float q[8];
    {
      const float4 a = __ldcg(reinterpret_cast<const float4*>(qrow)), c = __ldcg(reinterpret_cast<const float4*>(qrow) + 1);
      q[0] = a.x * 0.125f; q[1] = a.y * 0.125f; q[2] = a.z * 0.125f; q[3] = a.w * 0.125f;
      q[4] = c.x * 0.125f; q[5] = c.y * 0.125f; q[6] = c.z * 0.125f; q[7] = c.w * 0.125f;
    }
    bf16* kb = reinterpret_cast<bf16*>(p.kc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
    bf16* vb = reinterpret_cast<bf16*>(p.vc) + (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
    uint4 knew = make_uint4(0, 0, 0, 0), vnew = make_uint4(0, 0, 0, 0);   // the new row, bf16-rounded like the cached ones
    if (jb == n) {
      const float4 k0 = __ldcg(reinterpret_cast<const float4*>(qrow + kD)), k1 = __ldcg(reinterpret_cast<const float4*>(qrow + kD) + 1);
      const float4 v0 = __ldcg(reinterpret_cast<const float4*>(qrow + 2 * kD)), v1 = __ldcg(reinterpret_cast<const float4*>(qrow + 2 * kD) + 1);
      knew = make_uint4(pack_bf16x2(k0.x, k0.y), pack_bf16x2(k0.z, k0.w), pack_bf16x2(k1.x, k1.y), pack_bf16x2(k1.z, k1.w));
      vnew = make_uint4(pack_bf16x2(v0.x, v0.y), pack_bf16x2(v0.z, v0.w), pack_bf16x2(v1.x, v1.y), pack_bf16x2(v1.z, v1.w));
      if (g == 0) {   // KV append (modules/activation.py:148-152 re-allocates the whole cache instead)
        *reinterpret_cast<uint4*>(kb + (size_t)pos * kDH + sub * 8) = knew;
        *reinterpret_cast<uint4*>(vb + (size_t)pos * kDH + sub * 8) = vnew;
      }
    }
    const bf16* kp = kb + (size_t)g * kDH + sub * 8;
    const bf16* vp = vb + (size_t)g * kDH + sub * 8;
    uint4 kr[8], vr[8];
    auto load8 = [&](uint4 (&dst)[8], const bf16* base, const uint4& fresh, int j0) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int j = j0 + 4 * i + g;
        dst[i] = (j < jb && j != pos) ? __ldcg(reinterpret_cast<const uint4*>(base + (size_t)(j0 + 4 * i) * kDH)) : fresh;
      }
    };
    float m = -INFINITY, l = 0.f;
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    load8(kr, kp, knew, ja);
    for (int j0 = ja; j0 < jb; j0 += 32) {
      load8(vr, vp, vnew, j0);
      float sc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float w[8];
        unpack8(kr[i], w);
        float d = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) d = fmaf(q[e], w[e], d);
        d += __shfl_xor_sync(0xffffffffu, d, 1);
        d += __shfl_xor_sync(0xffffffffu, d, 2);
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        sc[i] = (j0 + 4 * i + g < jb) ? d : -INFINITY;
      }
      if (j0 + 32 < jb) load8(kr, kp, knew, j0 + 32);
      float bm = sc[0];
#pragma unroll
      for (int i = 1; i < 8; ++i) bm = fmaxf(bm, sc[i]);
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 8));
      bm = fmaxf(bm, __shfl_xor_sync(0xffffffffu, bm, 16));
      const float mn = fmaxf(m, bm);              // finite: every block holds at least one key
      const float scale = expf(m - mn);           // 0 on the first block (m = -inf)
      l *= scale;
#pragma unroll
      for (int e = 0; e < 8; ++e) acc[e] *= scale;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float pr = expf(sc[i] - mn);        // 0 for masked keys
        l += pr;
        float w[8];
        unpack8(vr[i], w);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = fmaf(pr, w[e], acc[e]);
      }
      m = mn;
    }
    // fold the 4 key groups of the warp (fixed order)
    l += __shfl_xor_sync(0xffffffffu, l, 8);
    l += __shfl_xor_sync(0xffffffffu, l, 16);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 8);
      acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], 16);
    }
    float* att = p.att + (size_t)b * kD + h * kDH;
    if (npieces == 1) {   // the whole head in one segment
      if (g == 0) {
        const float inv = 1.0f / l;
        *reinterpret_cast<float4*>(att + sub * 8) = make_float4(acc[0] * inv, acc[1] * inv, acc[2] * inv, acc[3] * inv);
        *reinterpret_cast<float4*>(att + sub * 8 + 4) = make_float4(acc[4] * inv, acc[5] * inv, acc[6] * inv, acc[7] * inv);
      }
      continue;
    }
    float* rec0 = p.part + ((size_t)b * kH + h) * p.att_maxp * kRec;
    float* rec = rec0 + (size_t)piece * kRec;
    if (g == 0) {
      *reinterpret_cast<float4*>(rec + 4 + sub * 8) = make_float4(acc[0], acc[1], acc[2], acc[3]);
      *reinterpret_cast<float4*>(rec + 4 + sub * 8 + 4) = make_float4(acc[4], acc[5], acc[6], acc[7]);
      if (sub == 0) *reinterpret_cast<float2*>(rec) = make_float2(m, l);
    }
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence();       // the record is visible device-wide before the ticket is taken
      last = (atomicAdd(p.tickets + b * kH + h, 1) == npieces - 1);
      if (last) __threadfence();
    }
    last = __shfl_sync(0xffffffffu, last, 0);
    if (last) {
      // fold the pieces in piece order: lane owns output dims 2*lane, 2*lane+1
      float M = -INFINITY;
      for (int s = lane; s < npieces; s += 32) M = fmaxf(M, __ldcg(rec0 + (size_t)s * kRec));
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) M = fmaxf(M, __shfl_xor_sync(0xffffffffu, M, o));
      float L = 0.f, O0 = 0.f, O1 = 0.f;
      for (int s0 = 0; s0 < npieces; s0 += 4) {
        float2 ml[4], ov[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const bool ok = s0 + u < npieces;
          const float* rs = rec0 + (size_t)(s0 + u) * kRec;
          ml[u] = ok ? __ldcg(reinterpret_cast<const float2*>(rs)) : make_float2(-INFINITY, 0.f);
          ov[u] = ok ? __ldcg(reinterpret_cast<const float2*>(rs + 4) + lane) : make_float2(0.f, 0.f);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const float w = expf(ml[u].x - M);
          L = fmaf(w, ml[u].y, L);
          O0 = fmaf(w, ov[u].x, O0);
          O1 = fmaf(w, ov[u].y, O1);
        }
      }
      const float inv = 1.0f / L;
      *reinterpret_cast<float2*>(att + 2 * lane) = make_float2(O0 * inv, O1 * inv);
      if (lane == 0) p.tickets[b * kH + h] = 0;   // next use is behind two grid barriers
    }
  }
}

// ---- L2 prefetch of the NEXT layer's keys / values (issued when the attention phase of this layer is over) --------------------
// The GEMV phases leave HBM idle for ~40 us per layer while the attention phase is bandwidth-bound, so a share of the next
// attention phase's KV rows (frac16/16 of the 32-key blocks, interleaved so L2 hits and HBM reads overlap) is pulled into
// the 126 MB L2 by the TMA engine's prefetch path.  Same walk over the flattened key space as attn_phase_balanced: the warp
// that will read a block asks for it.
__device__ void kv_prefetch_balanced(const PParams& p, int layer, const AttnPlan& pl, int frac16) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int gw = blockIdx.x * PNW + warp;
  if (frac16 <= 0 || (long long)gw * pl.per >= pl.tb) return;
  int gb = gw * pl.per;
  const int gb1 = min(pl.tb, gb + pl.per);
  while (gb < gb1) {
    const unsigned cand = __ballot_sync(0xffffffffu, pl.nblk > 0 && pl.pb <= gb);
    const int b = 31 - __clz(cand);
    const int nb = __shfl_sync(0xffffffffu, pl.nblk, b), pbb = __shfl_sync(0xffffffffu, pl.pb, b), n = __shfl_sync(0xffffffffu, pl.n, b);
    const int r = gb - pbb, h = r / nb, blk0 = r - h * nb;
    const int cnt = min(gb1 - gb, nb - blk0);
    const size_t base = (size_t)layer * p.kv_layer_elems + ((size_t)b * kH + h) * p.ctx_cap * kDH;
    for (int i = lane; i < cnt; i += 32) {
      const int gbi = gb + i;
      if ((((gbi + 1) * frac16) >> 4) == ((gbi * frac16) >> 4)) continue;
      const int j0 = (blk0 + i) * 32, nk = min(32, n - 1 - j0);   // row n-1 is written by that layer's attention phase
      if (nk <= 0) continue;
      const char* kp = reinterpret_cast<const char*>(p.kc) + (base + (size_t)j0 * kDH) * sizeof(bf16);
      const char* vp = reinterpret_cast<const char*>(p.vc) + (base + (size_t)j0 * kDH) * sizeof(bf16);
      const unsigned int bytes = (unsigned int)nk * kDH * (unsigned int)sizeof(bf16);
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(kp), "r"(bytes) : "memory");
      asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(vp), "r"(bytes) : "memory");
    }
    gb += cnt;
  }
}

template <int MR>
__global__ void __launch_bounds__(PNT, 1) ar_decode_mma_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  uint16_t* fh = reinterpret_cast<uint16_t*>(dsm);        // [MR*1024] hi halves, fragment order
  uint16_t* fl = fh + MR * kD;                             // [MR*1024] lo halves
  float* scratch = dsm + MR * kD;                          // sampling scratch / partial tiles
  float* red = scratch;                                    // [8][32][MR + 1] partial tiles (GEMV phases only)
  __shared__ int s_fold;
  __shared__ __align__(16) float lnp[2 * kD];              // gamma | beta of the next LayerNorm phase
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  // FFN2 is split 4 ways over K: CTA (kc, i) multiplies rows [g0,g1) by hidden units [kc*1024, +1024)  (grid % 4 == 0)
  const int per = gridDim.x / 4;
  const int kc = (int)blockIdx.x / per, gi = (int)blockIdx.x - kc * per;
  const int g0 = (int)(((long long)gi * kD) / per), g1 = (int)(((long long)(gi + 1) * kD) / per);
  int32_t* fold_ticket = p.tickets + B * kH + gi;

  MPre pre;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
  ln_copy_async(lnp, p.layer[0].n1_g, p.layer[0].n1_b);

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;
  auto zero = [](int, int) { return 0.f; };
  auto xa_row = [&](int m) -> const float* { return p.xa + (size_t)m * kD; };

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0: sample, append, embed the new token -> xa (models/vallex.py:529-531,568-598) -------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b) && !(step == 0 && p.resume && !ld_cg(p.pending + b))) sample_utterance(p, b, scratch);   // resumed launch: a continuing utterance sampled its token before the pause
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (p.stop_below > 0 && ld_cg(bt.n_active) <= p.stop_below) break;   // continuous batching: hand a finished slot back to the host
    if (step == p.max_steps - 1) break;
    const AttnPlan plan = attn_plan(bt);   // contexts are fixed for the 12 layers of this step
    const unsigned live = __ballot_sync(0xffffffffu, plan.n > 0);   // utterances still decoding (bit b)
    // share of a layer's KV blocks to pull into L2 ahead of its attention phase: all of them while one layer's keys and values
    // (8 KB per block) fit the budget (p.kv_l2_mb, default 80 of the 126 MB), else that fraction
    const int pf16 = (p.kv_l2_mb <= 0) ? 0 : (int)min(16ll, (16ll * p.kv_l2_mb * 128) / max(plan.tb, 1));

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      // ---- P1: LN1(x_l) + QKV rows (x_l is complete in xa: the FFN2 fold of layer l-1 wrote it) ---------------------------
      stage_rows_frag<MR, true>(fh, fl, B, xa_row, lnp);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, q0, q1, B, L.qkv_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention over the flattened key space; no weight registers are held across this (long) phase ------------
      attn_phase_balanced(p, l, plan);
      gemv_issue_mma(pre, Wo, kD, o0, o1);
      PSTAMP(6);
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: out-proj + residual -> xa -------------------------------------------------------------------------------
      kv_prefetch_balanced(p, (l + 1 < kL) ? l + 1 : 0, plan, pf16);   // next layer (or layer 0 of the next step)
      stage_rows_frag<MR, false>(fh, fl, B, [&](int m) -> const float* { return ((live >> m) & 1u) ? p.att + (size_t)m * kD : nullptr; }, nullptr);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, o0, o1, B, L.out_b, nullptr, red, [&](int n, int m) { return __ldcg(p.xa + (size_t)m * kD + n); },
                          [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue_mma(pre, W1, kD, f0, f1);
      ln_copy_async(lnp, L.n2_g, L.n2_b);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU ---------------------------------------------------------------------------------------
      stage_rows_frag<MR, true>(fh, fl, B, xa_row, lnp);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, f0, f1, B, L.ff1_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue_mma(pre, W2, kFF, g0, g1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2 split-K x4; the last of the four k-quarter CTAs of a row group folds the partials (fixed order), the
      //          bias and the residual into xa, so every consumer reads ONE value per element ------------------------------
      stage_rows_frag<MR, false>(fh, fl, B, [&](int m) -> const float* { return p.hid + (size_t)m * kFF + kc * kD; }, nullptr);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, g0, g1, B, nullptr, nullptr, red, zero,
                          [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
      PSTAMP(4);
      if (l + 1 < kL) {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1);
        ln_copy_async(lnp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b);
      } else {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1);
        ln_copy_async(lnp, p.norm_g, p.norm_b);
      }
      if (tid == 0) {
        __threadfence();       // this CTA's partial tile is visible device-wide before the ticket is taken
        s_fold = (atomicAdd(fold_ticket, 1) == 3);
        if (s_fold) __threadfence();
      }
      __syncthreads();
      if (s_fold) {
        const int nr = g1 - g0;
        for (int t = tid; t < nr * B; t += PNT) {
          const int m = t / nr, n = g0 + t - m * nr;
          float s4 = 0.f;
#pragma unroll
          for (int sp = 0; sp < 4; ++sp) s4 += __ldcg(p.pf + ((size_t)sp * B + m) * kD + n);
          p.xa[(size_t)m * kD + n] = (__ldcg(p.xa + (size_t)m * kD + n) + s4) + L.ff2_b[n];
        }
        if (tid == 0) *fold_ticket = 0;   // next use is behind four grid barriers
      }
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF: final LayerNorm + predict rows -> logits ----------------------------------------------------------------------
    stage_rows_frag<MR, true>(fh, fl, B, xa_row, lnp);
    gemv_finish_mma<MR>(pre, fh, fl, pr0, pr1, B, nullptr, nullptr, red, zero,
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
    ln_copy_async(lnp, p.layer[0].n1_g, p.layer[0].n1_b);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
static cudaError_t mma_launch(const PParams& p, int grid, cudaStream_t st) {
  if (grid % 4 != 0 || p.att_maxp <= 0) return cudaErrorInvalidConfiguration;   // FFN2 k-quarters; balanced attention buffers
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t red_floats = (size_t)8 * 32 * (MR + 1);
  const size_t smem = ((size_t)MR * kD + std::max(sample_floats, red_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_mma_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_mma_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_mma_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}


}  // namespace

cudaError_t launch_decode_mma(const PParams& p, int mr, int grid, cudaStream_t st) {
  switch (mr) {
    case 8: return mma_launch<8>(p, grid, st);
    case 16: return mma_launch<16>(p, grid, st);
    default: return mma_launch<32>(p, grid, st);
  }
}

}  // namespace vx
