// diag.cuh -- measurement helpers (not on the hot path): FP32 FFMA peak microbenchmark used as
// the second roofline denominator of the CUDA-core path (MEASURED_PEAKS.json has no FP32 figure).
#pragma once
#include <cuda_runtime.h>

namespace gmpc {

__global__ void __launch_bounds__(256) ffma_peak_kernel(float* out, int iters, float seed) {
  float a[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) a[i] = seed + (float)(threadIdx.x + i);
  const float x = 1.0000001f, y = 1e-9f;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
#pragma unroll
      for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a[i];
  if (s == 123.456f) out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace gmpc

#include "tc_common.cuh"

namespace gmpc {

// tcgen05 probe: D[128][NB] = A[128][K] * B[NB][K]^T through SWIZZLE_NONE descriptors.  The host
// chooses the shared-memory image strides and the descriptor LBO/SBO independently, so the
// descriptor semantics can be pinned on hardware (tests/test_gpu_tc_probe.py).
//   A image, K-major  (a_major 0): (k/4)*a_s1 + (m/8)*a_s2 + (m%8)*16 + (k%4)*4
//   A image, MN-major (a_major 1): (m/4)*a_s1 + (k/8)*a_s2 + (k%8)*16 + (m%4)*4
//   B image, K-major             : (k/4)*b_s1 + (n/8)*b_s2 + (n%8)*16 + (k%4)*4
__global__ void __launch_bounds__(128) tc_probe_kernel(
    const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int K, int NB,
    int a_major, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_s1, uint32_t a_s2, uint32_t a_kstep,
    uint32_t b_lbo, uint32_t b_sbo, uint32_t b_s1, uint32_t b_s2, uint32_t a_bytes) {
  extern __shared__ __align__(128) uint8_t psm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* a_img = psm;
  uint8_t* b_img = psm + a_bytes;
  if (a_major & 2) {  // raw image mode: A is copied linearly (decoding which element the MMA reads)
    for (int e = tid; e < (int)(a_bytes / 4); e += 128) reinterpret_cast<float*>(a_img)[e] = A[e];
  }
  for (int e = tid; e < 128 * K && !(a_major & 2); e += 128) {
    const int m = e / K, k = e - m * K;
    uint32_t off = a_major == 0 ? (k / 4) * a_s1 + (m / 8) * a_s2 + (m % 8) * 16 + (k % 4) * 4
                                : (m / 4) * a_s1 + (k / 8) * a_s2 + (k % 8) * 16 + (m % 4) * 4;
    *reinterpret_cast<float*>(a_img + off) = A[e];
  }
  for (int e = tid; e < NB * K; e += 128) {
    const int n = e / K, k = e - n * K;
    const uint32_t off = (k / 4) * b_s1 + (n / 8) * b_s2 + (n % 8) * 16 + (k % 4) * 4;
    *reinterpret_cast<float*>(b_img + off) = B[e];
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  if (warp == 0) tmem_alloc(&tmem_base_s, NB < 32 ? 32 : NB);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  if (tid == 0) {
    const uint32_t idesc = umma_idesc_tf32(NB, a_major & 1);
    for (int s = 0; s < K / 8; ++s) {
      const uint64_t ad = umma_smem_desc(smem_u32(a_img) + s * a_kstep, a_lbo, a_sbo);
      const uint64_t bd = umma_smem_desc(smem_u32(b_img) + s * 2 * b_s1, b_lbo, b_sbo);
      umma_tf32(tmem_base, ad, bd, idesc, s > 0 ? 1u : 0u);
    }
    umma_commit(&bar);
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < NB; c0 += 16) {
    float v[16];
    tmem_ld16(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(size_t)tid * NB + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, NB < 32 ? 32 : NB);
}

}  // namespace gmpc

namespace gmpc {

// tcgen05.mma issue/execute-rate microbenchmark (timing only; operand contents are arbitrary).
// One elected thread issues `reps` rounds of `ksteps` MMAs (M=128, N, K=8 tf32), walking the A
// descriptor through a `ksteps`-deep image exactly like the planner does, then commits and waits.
// layout_type 0 = SWIZZLE_NONE (lbo/sbo as given), 2 = SWIZZLE_128B.  out[blockIdx] = cycles.
__global__ void __launch_bounds__(128) tc_mma_bench_kernel(long long* out, int N, int ksteps,
                                                           int reps, uint32_t a_lbo, uint32_t a_sbo,
                                                           uint32_t a_kstep, uint32_t b_lbo,
                                                           uint32_t b_sbo, uint32_t b_kstep,
                                                           uint32_t layout_type, int two_mma) {
  extern __shared__ __align__(1024) uint8_t bsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 160 * 1024 / 4; i += 128) reinterpret_cast<float*>(bsm)[i] = 1.0f;
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_s;
  long long cyc = 0;
  if (warp == 1) {
    const uint32_t idesc = umma_idesc_tf32(N, 0), idesc2 = umma_idesc_tf32(N / 2, 0);
    const uint32_t a_base = smem_u32(bsm), b_base = smem_u32(bsm) + 112 * 1024;
    const uint64_t lt = (uint64_t)layout_type << 61;
    const long long t0 = clock64();
    if (elect_one()) {
      // descriptors are advanced by adding (bytes >> 4) to the low word: one add per operand
      const uint64_t ad0 = umma_smem_desc(a_base, a_lbo, a_sbo) | lt;
      const uint64_t bd0 = umma_smem_desc(b_base, b_lbo, b_sbo) | lt;
      const uint64_t a_inc = a_kstep >> 4, b_inc = b_kstep >> 4;
      for (int r = 0; r < reps; ++r) {
        uint64_t ad = ad0, bd = bd0;
        uint32_t acc = 0;
#pragma unroll 4
        for (int s = 0; s < ksteps; ++s) {
          // two_mma <= 1: same accumulator every time (+ optional dependent half-width MMA);
          // two_mma >= 2: cycle through `two_mma` independent accumulators (TMEM column ranges)
          umma_tf32(tmem_base + acc, ad, bd, idesc, 1u);
          if (two_mma == 1) umma_tf32(tmem_base + N / 2, ad, bd, idesc2, 1u);
          if (two_mma >= 2) acc = (acc + N >= (uint32_t)(two_mma * N)) ? 0u : acc + N;
          ad += a_inc;
          bd += b_inc;
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    cyc = clock64() - t0;
    if ((tid & 31) == 0) out[blockIdx.x] = cyc;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace gmpc
