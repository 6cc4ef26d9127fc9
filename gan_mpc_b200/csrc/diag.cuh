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
