// Experimental 3-barrier decode kernel with TMA-staged weights (VX_DECODE_TMA=1; written after the round-1 GPU budget
// ran out: compile-checked only).
//
// Same phases, arithmetic and summation order as ar_decode_fused_kernel (decode_fused.cu); what changes is how the
// weights reach the SM.  The fused kernel "prefetches" a phase's rows with ordinary loads into registers before the
// preceding barrier; if the SM's outstanding-load capacity is exhausted those loads stall at ISSUE and the streaming
// sits on the critical path after all.  Here every weight slice of a CTA -- its QKV rows, FFN1 rows, W2^T rows and
// predict rows, each one contiguous range of global memory -- is copied by the TMA engine (cp.async.bulk, issued by
// ONE thread, completion on an mbarrier) into its own shared-memory slot, and a slot is refilled with the NEXT
// layer's slice the moment its consumer is done: the copies run a whole layer (~20 us) ahead of their use, under
// all the barriers, staging and attention of the layers in between.  The dot products read the rows from shared
// memory with conflict-free 128-bit loads.
#include "decode_fused_common.cuh"

namespace vx {
namespace {
using namespace tc;

constexpr int TQ_ROWS = 22;   // QKV rows per CTA (3072 / 148 -> 21 or 22)
constexpr int TF_ROWS = 28;   // hidden units per CTA (4096 / 148 -> 27 or 28): FFN1 rows and W2^T rows
constexpr int TP_ROWS = 8;    // predict rows per CTA (1025 / 148 -> 6 or 7)
constexpr int kRowBytes = kD * 2;
constexpr int TQ_BYTES = TQ_ROWS * kRowBytes, TF_BYTES = TF_ROWS * kRowBytes, TP_BYTES = TP_ROWS * kRowBytes;
constexpr int TSLOT_BYTES = TQ_BYTES + 2 * TF_BYTES + TP_BYTES;   // 176,128 B

struct Slot {
  uint8_t* mem;
  uint64_t* bar;
  unsigned int issued, waited;   // fills requested / consumed (thread-uniform)
};

// thread 0: arm the slot's barrier and start the bulk copies of `rows` contiguous 2 KB rows
__device__ __forceinline__ void slot_fill(Slot& s, const bf16* __restrict__ src, int rows) {
  if (threadIdx.x == 0 && rows > 0) {
    const uint32_t bytes = (uint32_t)rows * kRowBytes;
    mbar_expect_tx(s.bar, bytes);
    for (uint32_t off = 0; off < bytes; off += 32768u) {
      const uint32_t n = bytes - off < 32768u ? bytes - off : 32768u;
      bulk_g2s(s.mem + off, reinterpret_cast<const char*>(src) + off, n, s.bar);
    }
  }
  if (rows > 0) ++s.issued;
}
// all threads: wait for the oldest outstanding fill of the slot
__device__ __forceinline__ void slot_wait(Slot& s, int rows) {
  if (rows > 0) {
    mbar_wait(s.bar, s.waited & 1u);
    ++s.waited;
  }
}

// rows [r0,r1) of a matrix whose slice sits in `slot` (row r0 first): two rows per warp in one pass over x, like
// gemv_finish; emit(n, m, dot + bias[n]) by lane m
template <int MR, typename Emit>
__device__ __forceinline__ void slot_gemv(const uint8_t* slot, int r0, int r1, const float* __restrict__ sx, int B,
                                          const float* __restrict__ bias, Emit emit) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int base = warp; r0 + base < r1; base += 2 * PNW) {   // one trip when the slice has <= 32 rows
    const int i0 = base, i1 = base + PNW;
    const bool has1 = r0 + i1 < r1;
    float a0[MR], a1[MR];
#pragma unroll
    for (int m = 0; m < MR; ++m) { a0[m] = 0.f; a1[m] = 0.f; }
#pragma unroll
    for (int it = 0; it < 4; ++it) {
      const uint4 q0 = *reinterpret_cast<const uint4*>(slot + (size_t)i0 * kRowBytes + it * 512 + lane * 16);
      const uint4 q1 = has1 ? *reinterpret_cast<const uint4*>(slot + (size_t)i1 * kRowBytes + it * 512 + lane * 16) : make_uint4(0, 0, 0, 0);
      float w0[8], w1[8];
      raw_to_f32<bf16>(&q0, 0, w0);
      raw_to_f32<bf16>(&q1, 0, w1);
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
      emit(r0 + i0, lane, m0 + (bias ? bias[r0 + i0] : 0.f));
      if (has1) emit(r0 + i1, lane, m1 + (bias ? bias[r0 + i1] : 0.f));
    }
  }
}

// hidden units [f0,f1) (values in hs) times their rows of W2^T in `slot` -> accF[m][0..1023]
template <int MR>
__device__ __forceinline__ void slot_ffn2_share(const uint8_t* slot, int f0, int f1, const float* __restrict__ hs, int B,
                                                long long* accF) {
  const int tid = threadIdx.x;
  float acc[MR][2];
#pragma unroll
  for (int m = 0; m < MR; ++m) { acc[m][0] = 0.f; acc[m][1] = 0.f; }
  const uint32_t* base = reinterpret_cast<const uint32_t*>(slot) + tid;   // columns 2*tid, 2*tid+1
  for (int j = 0; j < f1 - f0; ++j) {
    const uint32_t w = base[(size_t)j * (kD / 2)];
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
__global__ void __launch_bounds__(PNT, 1) ar_decode_tma_kernel(const PParams p) {
  using WT = bf16;
  extern __shared__ __align__(128) uint8_t tsm[];
  uint8_t* wq = tsm;                                  // QKV rows of the coming P1
  uint8_t* wf1 = wq + TQ_BYTES;                       // FFN1 rows of the coming P3
  uint8_t* wf2 = wf1 + TF_BYTES;                      // W2^T rows of the coming P3
  uint8_t* wpd = wf2 + TF_BYTES;                      // predict rows of the coming PF
  uint64_t* bars = reinterpret_cast<uint64_t*>(tsm + TSLOT_BYTES);   // 4 mbarriers
  float* sx = reinterpret_cast<float*>(tsm + TSLOT_BYTES + 64);      // [MR][1024]
  float* hs = sx + MR * kD;                           // [MR][kHS]
  float* scratch = hs + MR * kHS;
  const Batch& bt = p.bt;
  const int B = bt.B;
  const int tid = threadIdx.x;
  unsigned int epoch = 0;

  int q0, q1, o0, o1, f0, f1, pr0, pr1;
  slice(3 * kD, q0, q1);
  slice(kD, o0, o1);
  slice(kFF, f0, f1);
  slice(kARLogits, pr0, pr1);
  const int ns = p.nsplit;
  const bool has_unit = (int)blockIdx.x < B * kH * ns;
  const int u_split = blockIdx.x % ns, u_bh = blockIdx.x / ns, u_b = has_unit ? u_bh / kH : 0, u_h = u_bh % kH;
  const int wr0 = (u_split * kD) / ns, wr1 = ((u_split + 1) * kD) / ns;
  int att_epoch = 0;
  long long* const accO2 = p.acc_out;
  long long* const accF2 = p.acc_ffn;
  const size_t accN = (size_t)B * kD;

  if (tid == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&bars[i], 1);
    fence_barrier_init();
  }
  for (int i = tid; i < MR * kHS; i += PNT) hs[i] = 0.f;
  __syncthreads();
  Slot sq{wq, &bars[0], 0, 0}, sf1{wf1, &bars[1], 0, 0}, sf2{wf2, &bars[2], 0, 0}, spd{wpd, &bars[3], 0, 0};
  auto Wqkv = [&](int l) { return reinterpret_cast<const WT*>(p.layer[l].qkv_w) + (size_t)q0 * kD; };
  auto Wff1 = [&](int l) { return reinterpret_cast<const WT*>(p.layer[l].ff1_w) + (size_t)f0 * kD; };
  auto Wff2 = [&](int l) { return reinterpret_cast<const WT*>(p.layer[l].ff2_t) + (size_t)f0 * kD; };
  const WT* Wpred = reinterpret_cast<const WT*>(p.predict_w) + (size_t)pr0 * kD;
  // layer 0's slices and the predict rows start streaming before the first sample is drawn
  slot_fill(sq, Wqkv(0), q1 - q0);
  slot_fill(sf1, Wff1(0), f1 - f0);
  slot_fill(sf2, Wff2(0), f1 - f0);
  slot_fill(spd, Wpred, pr1 - pr0);

  WoPre wp;
  LnPre lp;
  ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);

  long long* prof = nullptr;
  const int prof_slot = (blockIdx.x == 0) ? 0 : (blockIdx.x == gridDim.x / 2) ? 1 : (blockIdx.x == gridDim.x - 1) ? 2 : -1;

  int steps = 0;
  for (int step = 0; step < p.max_steps; ++step) {
    if (p.prof && prof_slot >= 0 && tid == 0) {
      if (step == p.prof_step0) prof = p.prof + (size_t)prof_slot * 4096;
      if (step == p.prof_step0 + 4) prof = nullptr;
    }
    // ---- S0 --------------------------------------------------------------------------------------------------------
    for (int b = blockIdx.x; b < B; b += gridDim.x)
      if (!ld_cg(bt.finished + b)) sample_utterance(p, b, scratch);
    zero_slice(accF2 + accN, B, o0, o1);
    grid_barrier(p.bar, epoch, prof);
    ++steps;
    if (!p.ignore_eos && ld_cg(bt.n_active) <= 0) break;
    if (step == p.max_steps - 1) break;

    for (int l = 0; l < kL; ++l) {
      const PLayer& L = p.layer[l];
      const int par = l & 1;
      const int ln = (l + 1 < kL) ? l + 1 : 0;   // the layer whose slices are requested next (layer 0 of the next step after layer 11)
      long long* accO = accO2 + par * accN;
      long long* accF = accF2 + par * accN;
      long long* accO_prev = accO2 + (par ^ 1) * accN;
      long long* accF_prev = accF2 + (par ^ 1) * accN;
      const float* fb_prev = (l > 0) ? p.layer[l - 1].ff2_b : nullptr;
      const WT* Wo = reinterpret_cast<const WT*>(L.out_w);
      // ---- P1 ------------------------------------------------------------------------------------------------------
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
      if (has_unit) wo_issue(wp, Wo, u_h, wr0, wr1);   // 2 x 16 B per lane: small enough for registers
      ln_load(lp, L.n1_g, L.n1_b, fb_prev);
      if (l == 0) {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c); }, lp, scratch, nullptr, 0, 0);
      } else {
        stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF_prev + (size_t)m * kD + c); },
                       lp, scratch, p.xa, o0, o1);
        zero_slice(accO_prev, B, o0, o1);
      }
      PSTAMP(3);
      slot_wait(sq, q1 - q0);
      slot_gemv<MR>(wq, q0, q1, sx, B, L.qkv_b, [&](int n, int m, float v) { p.qkv[(size_t)m * 3 * kD + n] = v; });
      __syncthreads();                       // every warp is done with the slot: refill it with the next layer's rows
      slot_fill(sq, Wqkv(ln), q1 - q0);
      PSTAMP(4);
      grid_barrier(p.bar, epoch, prof);
      // ---- P2 ------------------------------------------------------------------------------------------------------
      if (has_unit && !ld_cg(bt.finished + u_b))
        attn_unit_fused(p, l, u_b, u_h, u_split, scratch, att_epoch, wp, Wo, wr0, wr1, accO);
      if (l > 0) zero_slice(accF_prev, B, o0, o1);
      ln_issue(lp, L.n2_g, L.n2_b, L.out_b);
      grid_barrier(p.bar, epoch, prof);
      // ---- P3 ------------------------------------------------------------------------------------------------------
      ln_load(lp, L.n2_g, L.n2_b, L.out_b);
      stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xa + (size_t)m * kD + c) + fix_get(accO + (size_t)m * kD + c); },
                     lp, scratch, p.xb, o0, o1);
      PSTAMP(3);
      slot_wait(sf1, f1 - f0);
      slot_gemv<MR>(wf1, f0, f1, sx, B, L.ff1_b, [&](int n, int m, float v) { hs[m * kHS + (n - f0)] = fmaxf(v, 0.f); });
      __syncthreads();                       // hs complete; FFN1 slot free
      slot_fill(sf1, Wff1(ln), f1 - f0);
      PSTAMP(4);
      slot_wait(sf2, f1 - f0);
      slot_ffn2_share<MR>(wf2, f0, f1, hs, B, accF);
      __syncthreads();                       // W2^T slot free
      slot_fill(sf2, Wff2(ln), f1 - f0);
      if (l + 1 < kL) ln_issue(lp, p.layer[l + 1].n1_g, p.layer[l + 1].n1_b, L.ff2_b);
      else ln_issue(lp, p.norm_g, p.norm_b, L.ff2_b);
      PSTAMP(5);
      grid_barrier(p.bar, epoch, prof);
    }
    // ---- PF --------------------------------------------------------------------------------------------------------
    ln_load(lp, p.norm_g, p.norm_b, p.layer[kL - 1].ff2_b);
    stage_ln_x<MR>(sx, B, [&](int m, int c) { return __ldcg(p.xb + (size_t)m * kD + c) + fix_get(accF2 + accN + (size_t)m * kD + c); },
                   lp, scratch, nullptr, 0, 0);
    zero_slice(accO2 + accN, B, o0, o1);
    slot_wait(spd, pr1 - pr0);
    slot_gemv<MR>(wpd, pr0, pr1, sx, B, nullptr, [&](int n, int m, float v) { p.logits[(size_t)m * kARLogits + n] = v; });
    __syncthreads();
    slot_fill(spd, Wpred, pr1 - pr0);
    ln_issue(lp, p.layer[0].n1_g, p.layer[0].n1_b, nullptr);
    grid_barrier(p.bar, epoch, prof);
  }
  // drain the copies still in flight before the CTA (and its shared memory) goes away
  while (sq.waited < sq.issued) slot_wait(sq, 1);
  while (sf1.waited < sf1.issued) slot_wait(sf1, 1);
  while (sf2.waited < sf2.issued) slot_wait(sf2, 1);
  while (spd.waited < spd.issued) slot_wait(spd, 1);
  if (blockIdx.x == 0 && tid == 0) *p.steps_done = steps;
}

template <int MR>
cudaError_t tma_launch(const PParams& p, int grid, cudaStream_t st) {
  const size_t scratch_floats = 256 + PCHUNK + (PNT / 8) * kDH;
  const size_t sample_floats = 1032 + PSORT + PSORT / 2 + PNT + 32 + 8;
  const size_t smem = (size_t)TSLOT_BYTES + 64 +
                      ((size_t)MR * kD + (size_t)MR * kHS + std::max(scratch_floats, sample_floats)) * sizeof(float);
  cudaError_t e = cudaFuncSetAttribute(ar_decode_tma_kernel<MR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, ar_decode_tma_kernel<MR>, PNT, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) return cudaErrorLaunchOutOfResources;
  void* args[] = {(void*)&p};
  return cudaLaunchCooperativeKernel((void*)ar_decode_tma_kernel<MR>, dim3(grid), dim3(PNT), args, smem, st);
}

}  // namespace

// B <= 2; the slot capacities need every slice of the grid to fit (grid >= 147 CTAs)
cudaError_t launch_decode_tma(const PParams& p, int mr, int grid, cudaStream_t st) {
  if (mr > 2 || (3 * kD + grid - 1) / grid > TQ_ROWS || (kFF + grid - 1) / grid > TF_ROWS || (kARLogits + grid - 1) / grid > TP_ROWS)
    return cudaErrorInvalidConfiguration;
  return mr == 1 ? tma_launch<1>(p, grid, st) : tma_launch<2>(p, grid, st);
}

}  // namespace vx
