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

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wq = reinterpret_cast<const WT*>(L.qkv_w);
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      const WT* W1 = reinterpret_cast<const WT*>(L.ff1_w);
      const WT* W2 = reinterpret_cast<const WT*>(L.ff2_w) + (size_t)kc * kD;
      (void)Wq;
      // ---- P1: LN1 + QKV rows ----------------------------------------------------------------------------------------
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
      stage_ln_frag<MR>(fh, fl, B, [&](int m, int c) { return stream_dyn(p, l, m, c, B); }, lp, scratch);
      PSTAMP(3);
      gemv_finish_mma<MR>(pre, fh, fl, q0, q1, B, L.qkv_b, nullptr, red, zero,
                          [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      PSTAMP(4);
      gemv_issue_mma(pre, Wo, kD, o0, o1);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2: attention (ticket fold) -----------------------------------------------------------------------------------
      for (int u = blockIdx.x; u < B * kH * p.nsplit; u += gridDim.x) {
        const int split = u % p.nsplit, bh = u / p.nsplit;
        const int b = bh / kH, h = bh % kH;
        if (ld_cg(bt.finished + b)) continue;
        if constexpr (PIPE) attn_unit_pipe(p, l, b, h, split, scratch);
        else attn_unit<bf16>(p, l, b, h, split, scratch);
      }
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
