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

// tcgen05 kind::f16 dense peak of the pipe the planner kernels use: every SM issues back-to-back
// M=128, N=256, K=16 MMAs (shared-memory operands, contents irrelevant) into two accumulators.
__global__ void __launch_bounds__(128) f16_mma_peak_kernel(int mmas) {
  extern __shared__ __align__(128) uint8_t dsm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(dsm)[i] = 0x3C003C00u;
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
  const uint32_t tb = tmem_base_s;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t idesc = (1u << 4) | ((uint32_t)(256 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
      const uint64_t ad = umma_smem_desc(smem_u32(dsm), 128 * 16, 128);
      const uint64_t bd = umma_smem_desc(smem_u32(dsm) + 8192, 256 * 16, 128);
      for (int i = 0; i < mmas; ++i)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                     "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tb + (i & 1) * 256), "l"(ad),
                     "l"(bd), "r"(idesc), "r"(1u) : "memory");
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tb, 512);
  }
}

}  // namespace gmpc
