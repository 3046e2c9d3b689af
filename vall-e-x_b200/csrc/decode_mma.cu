// Tensor-core batched decode kernel (8 <= B <= 32): the five GEMV phases of the 62-barrier loop on mma.sync.
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

template <int MR, typename Val>
__device__ __forceinline__ void stage_ln_frag(uint16_t* fh, uint16_t* fl, int B, Val value, const LnPre& lp, float* red) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  constexpr int R = 8;
  for (int mb = 0; mb < B; mb += R) {
    const int nb = min(R, B - mb);
    float v0[R], v1[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
      v0[m] = (m < nb) ? value(mb + m, tid) + lp.f0 : 0.f;
      v1[m] = (m < nb) ? value(mb + m, tid + PNT) + lp.f1 : 0.f;
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
        frag_store<MR>(fh, fl, mb + m, tid, __fadd_rn(__fmul_rn((v0[m] - mean[m]) * rstd, lp.g0), lp.b0));
        frag_store<MR>(fh, fl, mb + m, tid + PNT, __fadd_rn(__fmul_rn((v1[m] - mean[m]) * rstd, lp.g1), lp.b1));
      }
    }
    __syncthreads();
  }
}
template <int MR, typename Val>
__device__ __forceinline__ void stage_plain_frag(uint16_t* fh, uint16_t* fl, int B, Val value) {
  for (int i = threadIdx.x; i < B * kD; i += PNT) frag_store<MR>(fh, fl, i / kD, i % kD, value(i / kD, i % kD));
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
// emit(n, m, dot + addend(n, m) + bias[n] + bias2[n]).  red: [8][32][MR] floats.
template <int MR, typename Add, typename Emit>
__device__ __forceinline__ void gemv_finish_mma(const MPre& pre, const uint16_t* fh, const uint16_t* fl, int r0, int r1, int B,
                                                const float* __restrict__ bias, const float* __restrict__ bias2, float* red,
                                                Add addend, Emit emit) {
  constexpr int NT = MR / 8;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int mt = warp & 1, kg = warp >> 1, g = lane >> 2, q = lane & 3;
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
    float* d0 = red + ((size_t)kg * 32 + row) * MR + n;
    float* d1 = red + ((size_t)kg * 32 + row + 8) * MR + n;
    d0[0] = c[nt][0]; d0[1] = c[nt][1];
    d1[0] = c[nt][2]; d1[1] = c[nt][3];
  }
  __syncthreads();
  for (int t = tid; t < 32 * MR; t += PNT) {
    const int row = t / MR, m = t % MR, n = r0 + row;
    if (n < r1 && m < B) {
      float s = 0.f;
#pragma unroll
      for (int k8 = 0; k8 < 8; ++k8) s += red[((size_t)k8 * 32 + row) * MR + m];
      emit(n, m, s + (addend(n, m) + (bias ? bias[n] : 0.f) + (bias2 ? bias2[n] : 0.f)));
    }
  }
  __syncthreads();
}

// attention unit of the batched kernel with a pipelined P.V loop (PIPE = true; VX_DECODE_MMA=2).  Same arithmetic and
// summation order as attn_unit<bf16>: a lane still walks its keys in increasing order, only the loads are issued ahead.
__device__ void attn_unit_pipe(const PParams& p, int layer, int b, int h, int split, float* smem) {
  float* sq = smem;                  // [64]
  float* sknew = smem + 64;          // [64]
  float* svnew = smem + 128;         // [64]
  float* red = smem + 192;           // [32]
  int* s_last = reinterpret_cast<int*>(smem + 224);
  float* sp = smem + 256;            // [PCHUNK]
  float* sacc = sp + PCHUNK;         // [PNT/8][64]
  const Batch& bt = p.bt;
  const int tid = threadIdx.x;
  const int pos = bt.S[b] + bt.P[b] + ld_cg(bt.T + b);
  const int n = pos + 1;
  int chunk = (n + p.nsplit - 1) / p.nsplit;
  chunk = (chunk + 31) & ~31;
  const int j0 = split * chunk, j1 = min(n, j0 + chunk);
  const float* row = p.qkv + (size_t)b * (3 * kD) + h * kDH;
  __syncthreads();  // smem reuse across units
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
  // P.V with 8 value chunks in flight per lane (a batched unit holds ALL keys of its head: ~20 keys per lane group; one
  // load per loop trip would put 20 dependent memory round trips on the critical path)
  constexpr int UV = 8;
  for (int jb = j0 + g; jb < j1; jb += NG * UV) {
    uint4 raw[UV];
    float prs[UV];
#pragma unroll
    for (int u = 0; u < UV; ++u) {
      const int j = jb + u * NG;
      const bool ok = j < j1;
      raw[u] = (ok && j != pos) ? __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)j * kDH + dc)) : make_uint4(0, 0, 0, 0);
      prs[u] = ok ? sp[j - j0] : 0.f;
    }
#pragma unroll
    for (int u = 0; u < UV; ++u) {
      const int j = jb + u * NG;
      float w[8];
      if (j == pos) {
#pragma unroll
        for (int i = 0; i < 8; ++i) w[i] = kv_round<bf16>(svnew[dc + i]);
      } else {
        raw_to_f32<bf16>(&raw[u], 0, w);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) acc[i] = fmaf(prs[u], w[i], acc[i]);
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) sacc[g * kDH + dc + i] = acc[i];
  __syncthreads();
  float od = 0.f;
  if (tid < kDH) {
    for (int q = 0; q < NG; ++q) od += sacc[q * kDH + tid];
  }
  // partial (m, l, o[64]) of this split.  The LAST split of (b, h) to arrive (ticket) folds all of them in fixed split
  // order -- deterministic whoever is last -- into att[b][h*64 ..], so the out-proj phase stages 4 KB per utterance
  // instead of every CTA gathering all partials.
  float* pp = p.part + (((size_t)b * kH + h) * p.nsplit + split) * (kDH + 2);
  if (tid < kDH) pp[2 + tid] = od;
  if (tid == 0) { pp[0] = m; pp[1] = l; }
  if (p.dbg & 2) return;   // bisecting aid: the out-proj phase folds the splits itself
  __syncthreads();
  if (tid == 0) {
    __threadfence();       // the CTA's partial is visible device-wide before the ticket is taken
    *s_last = (atomicAdd(p.tickets + b * kH + h, 1) == p.nsplit - 1);
  }
  __syncthreads();
  if (*s_last) {
    __threadfence();
    if (tid < kDH) p.att[(size_t)b * kD + h * kDH + tid] = att_combined(p, b, h * kDH + tid);
    if (tid == 0) p.tickets[b * kH + h] = 0;   // next use is behind a grid barrier
  }
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

template <int MR, bool PIPE>
__global__ void __launch_bounds__(PNT, 1) ar_decode_mma_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(16) float dsm[];
  uint16_t* fh = reinterpret_cast<uint16_t*>(dsm);        // [MR*1024] hi halves, fragment order
  uint16_t* fl = fh + MR * kD;                             // [MR*1024] lo halves
  float* scratch = dsm + MR * kD;                          // attention / sampling / LN scratch
  float* red = scratch;                                    // [8][32][MR] partial tiles (GEMV phases only)
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
  const int per = gridDim.x / 4;
  const int kc = min((int)blockIdx.x / per, 3);
  const int kc_first = kc * per;
  const int kc_n = (kc == 3) ? (int)gridDim.x - kc_first : per;
  const int g0 = (int)(((long long)(blockIdx.x - kc_first) * kD) / kc_n);
  const int g1 = (int)(((long long)(blockIdx.x - kc_first + 1) * kD) / kc_n);

  MPre pre;
  LnPre lp;
  const WT* Wqkv0 = reinterpret_cast<const WT*>(p.layer[0].qkv_w);
  gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;
  auto zero = [](int, int) { return 0.f; };

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;
    AttnPlan plan{};
    if (p.att_maxp > 0) plan = attn_plan(bt);   // contexts are fixed for the 12 layers of this step

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      (void)Wq;
      // ---- P1: LN1 + QKV rows ----------------------------------------------------------------------------------------
      for (int u = blockIdx.x; p.att_maxp == 0 && u < B * kH * p.nsplit; u += gridDim.x) {
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
      stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, lp, scratch);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, q0, q1, B, L.qkv_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      if (p.att_maxp == 0) gemv_issue_mma(pre, Wo, kD, o0, o1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention (ticket fold) -----------------------------------------------------------------------------------
      if (p.att_maxp > 0) {
        attn_phase_balanced(p, l, plan);       // no weight registers are held across this (long) phase
        gemv_issue_mma(pre, Wo, kD, o0, o1);
      } else {
        for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
          const int split = u % p.nsplit, bh = u / p.nsplit;
          const int b = bh / kH, h = bh % kH;
          if (ld_cg(bt.finished + b)) continue;
          if constexpr (PIPE) attn_unit_pipe(p, l, b, h, split, scratch);
          else attn_unit<bf16>(p, l, b, h, split, scratch);
        }
      }
      PSTAMP(6);
      grid_barrier(p.bar, epoch, prof);
      // ---- P3: out-proj + residual -> xa ---------------------------------------------------------------------------------
      stage_plain_frag<MR>(fh, fl, B, [&](int m, int c) {
        if (ld_cg(bt.finished + m)) return 0.f;
        return (p.dbg & 2) ? att_combined(p, m, c) : __ldcg(p.att + (size_t)m * kD + c);
      });
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, o0, o1, B, L.out_b, fb_prev, red, [&](int n, int m) { return stream_dyn(p, l, m, n, B); },
                          [&](int n, int m, float v) { p.xa[(size_t)m * kD + n] = v; });
      PSTAMP(4);
      gemv_issue_mma(pre, W1, kD, f0, f1);
      ln_issue(lp, L.n2_g, L.n2_b, nullptr);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P4: LN2 + FFN1 + ReLU -----------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, nullptr);
      stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, f0, f1, B, L.ff1_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.hid[(size_t)m * kFF + n] = fmaxf(v, 0.f); });
      PSTAMP(4);
      gemv_issue_mma(pre, W2, kFF, g0, g1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P5: FFN2, split-K -----------------------------------------------------------------------------------------------
      stage_plain_frag<MR>(fh, fl, B, [&](int m, int c) { return __ldcg(p.hid + (size_t)m * kFF + kc * kD + c); });
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, g0, g1, B, nullptr, nullptr, red, zero,
                          [&](int n, int m, float v) { p.pf[((size_t)kc * B + m) * kD + n] = v; });
      PSTAMP(4);
      if (l + 1 < kL) {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.layer[l + 1].qkv_w), kD, q0, q1);
        ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      } else {
        gemv_issue_mma(pre, reinterpret_cast<const WT*>(p.predict_w), kD, pr0, pr1);
        ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      }
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF ------------------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return stream_dyn(p, kL, m, c, B); }, lp, scratch);
    gemv_finish_mma<MR>(pre, fh, fl, pr0, pr1, B, nullptr, nullptr, red, zero,
                        [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    gemv_issue_mma(pre, Wqkv0, kD, q0, q1);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR, bool PIPE>
static cudaError_t mma_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t red_floats = (size_t)8 * 32 * MR;
  const size_t smem = ((size_t)MR * kD + std::max(std::max(scratch_floats, sample_floats), red_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_mma_kernel<MR, PIPE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_mma_kernel<MR, PIPE>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_mma_kernel<MR, PIPE>, dim3(grid), dim3(PNT), args, smem, st);
}


}  // namespace

cudaError_t launch_decode_mma(const PParams& p, int mr, int grid, cudaStream_t st) {
  const char* env = getenv("VX_DECODE_MMA");
  if (env && env[0] == '2') {   // experimental: pipelined P.V loads in the attention phase
    switch (mr) {
      case 8: return mma_launch<8, true>(p, grid, st);
      case 16: return mma_launch<16, true>(p, grid, st);
      default: return mma_launch<32, true>(p, grid, st);
    }
  }
  switch (mr) {
    case 8: return mma_launch<8, false>(p, grid, st);
    case 16: return mma_launch<16, false>(p, grid, st);
    default: return mma_launch<32, false>(p, grid, st);
  }
}

}  // namespace vx
