// h16_common.cuh -- operand images and helpers of the kind::f16 split-precision planner kernel.
//
// Split: v = hi + lo with hi = fp16(v), lo = fp16(v - hi); both operands of a contraction are
// split and three products are accumulated in fp32: Wh*ah + Wh*al + Wl*ah (22-23 mantissa bits,
// the same as the 3xTF32 scheme, at half the operand bytes and half the MMA count per unit of K).
// fp16's narrow exponent is handled by exact power-of-two scales: per-layer weight scale (chosen
// when the weights are packed), per-trajectory adjoint scale (backward sweep); undone in fp32 in
// the epilogue.
//
// A operand (weights), K-major, SWIZZLE_NONE, one "unit" per (128-row block, k-step of 16):
//     unit = [2 k-chunks][128 rows][8 halfs] = 4096 B;  element (r, kk) at
//     (kk/8)*2048 + (r/8)*128 + (r%8)*16 + (kk%8)*2          descriptor: LBO 2048, SBO 128
//   the hi unit is followed by the lo unit (8 KB per block-k-step).
// B operand (activations), MN-major, SWIZZLE_NONE: N = 64 columns (trajectory r -> column r for the
// hi part, 32 + r for the lo part), K = features:
//     element (n, f) at (f/8)*1024 + (f%8)*16 + (n/8)*128 + (n%8)*2     descriptor: LBO 1024, SBO 128
//   so one thread that owns a feature and 8 consecutive trajectories stores one 16-byte vector.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "tc_common.cuh"

namespace gmpc {

constexpr int H_NB = 32;                 // trajectories per tile
constexpr uint32_t H_UNIT = 4096;        // bytes of one A unit (hi or lo)
constexpr uint32_t H_A_LBO = 2048, H_A_SBO = 128;
constexpr uint32_t H_B_LBO = 1024, H_B_SBO = 128;
constexpr uint32_t H_B_KSTEP = 2 * H_B_LBO;  // bytes of B per k-step (16 features)

// Instruction descriptor, kind::f16 with fp16 A/B, fp32 accumulate, M = 128.
__host__ __device__ constexpr uint32_t h16_idesc(int N, int a_mn_major, int b_mn_major) {
  return (1u << 4)                                  // c_format = F32
         | (0u << 7) | (0u << 10)                   // a_format = b_format = F16
         | ((uint32_t)(a_mn_major & 1) << 15) | ((uint32_t)(b_mn_major & 1) << 16) |
         ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// (a, b) -> packed hi halfs and packed lo halfs.  Saturating conversion keeps an out-of-range
// value finite (the domain of the kernel is |scaled value| < 65504).
__device__ __forceinline__ void split_h2(float a, float b, uint32_t& hi, uint32_t& lo) {
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hi) : "f"(b), "f"(a));  // low half = a
  const __half2 h = *reinterpret_cast<const __half2*>(&hi);
  const float2 hf = __half22float2(h);
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lo) : "f"(b - hf.y), "f"(a - hf.x));
}
__device__ __forceinline__ void split_h1(float a, __half& hi, __half& lo) {
  hi = __float2half_rn(a);
  lo = __float2half_rn(a - __half2float(hi));
}

// Byte offset of (column n, feature f) in a B operand buffer.
__device__ __forceinline__ uint32_t h16_b_off(int n, int f) {
  return (uint32_t)((f >> 3) * H_B_LBO + (f & 7) * 16 + (n >> 3) * H_B_SBO + (n & 7) * 2);
}
// Store one activation (trajectory r, feature f), split, into a B operand buffer.
__device__ __forceinline__ void h16_store_op(uint8_t* buf, int f, int r, float v) {
  __half hi, lo;
  split_h1(v, hi, lo);
  *reinterpret_cast<__half*>(buf + h16_b_off(r, f)) = hi;
  *reinterpret_cast<__half*>(buf + h16_b_off(H_NB + r, f)) = lo;
}

// Exact power of two s with s * v in [2^3, 2^4) for v > 0 (1 for v == 0 or non-finite).
__device__ __forceinline__ float pow2_scale_to_8(float v) {
  if (!(v > 0.f) || v > 3.0e38f) return 1.f;
  const int e = (int)((__float_as_uint(v) >> 23) & 0xFF) - 127;  // v in [2^e, 2^(e+1)) (normal v)
  int k = 3 - e;
  k = k > 100 ? 100 : (k < -100 ? -100 : k);
  return __uint_as_float((uint32_t)(k + 127) << 23);
}

// 1 / s for s an exact power of two (normal range): one integer subtract instead of a division.
__device__ __forceinline__ float pow2_recip(float s) {
  return __uint_as_float(0x7F000000u - __float_as_uint(s));
}

// max |W| of one layer -> absmax[0] (float bits compare as unsigned for non-negative values)
__global__ void h16_absmax_kernel(const float* __restrict__ W, int count, uint32_t* absmax) {
  float mx = 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < count; i += gridDim.x * blockDim.x)
    mx = fmaxf(mx, fabsf(W[i]));
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
  if ((threadIdx.x & 31) == 0 && isfinite(mx)) atomicMax(absmax, __float_as_uint(mx));
}

}  // namespace gmpc
