// Shared device helpers and model constants for the vallex_b200 kernels (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace vx {

constexpr int kD = 1024;      // N_DIM               (reference macros.py:3)
constexpr int kH = 16;        // NUM_HEAD            (macros.py:2)
constexpr int kDH = 64;       // head dim
constexpr int kFF = 4096;     // dim_feedforward     (models/vallex.py:146)
constexpr int kL = 12;        // NUM_LAYERS          (macros.py:1)
constexpr int kNQ = 8;        // NUM_QUANTIZERS      (macros.py:5)
constexpr int kVocab = 1024;  // NUM_AUDIO_TOKENS    (models/macros.py:5)
constexpr int kEOS = 1024;
constexpr int kBOS = 1025;
constexpr int kARLogits = 1025;  // ar_predict_layer rows (models/vallex.py:154-156)
constexpr int kTextVocab = 2048;
constexpr int kSMs = 148;

typedef __nv_bfloat16 bf16;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024 (multiple of 32). `red` = 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();  // protect `red` from a previous use
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  return warp_max(r);
}

// 8 consecutive elements -> fp32 registers, 128-bit global loads.
__device__ __forceinline__ void load8(const float* __restrict__ p, float (&o)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w; o[4] = b.x; o[5] = b.y; o[6] = b.z; o[7] = b.w;
}
__device__ __forceinline__ void load8(const bf16* __restrict__ p, float (&o)[8]) {
  const uint4 r = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    o[2 * i] = __uint_as_float(w[i] << 16);
    o[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ void load4(const float* __restrict__ p, float (&o)[4]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  o[0] = a.x; o[1] = a.y; o[2] = a.z; o[3] = a.w;
}
__device__ __forceinline__ void load4(const bf16* __restrict__ p, float (&o)[4]) {
  const uint2 r = __ldg(reinterpret_cast<const uint2*>(p));
  o[0] = __uint_as_float(r.x << 16); o[1] = __uint_as_float(r.x & 0xffff0000u);
  o[2] = __uint_as_float(r.y << 16); o[3] = __uint_as_float(r.y & 0xffff0000u);
}
__device__ __forceinline__ float to_f32(float v) { return v; }
__device__ __forceinline__ float to_f32(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ void store_as(float* p, float v) { *p = v; }
__device__ __forceinline__ void store_as(bf16* p, float v) { *p = __float2bfloat16_rn(v); }

}  // namespace vx
