// mma_dual_rate.cu -- does the tensor pipe run two issue streams as fast as one?
// Issuer A (warp 1): Wh x [ah|al] (N = 64) per block-k-step; issuer B (warp 3): Wl x ah (N = 32) into
// separate accumulator columns.  mode 0: both products from issuer A alone; mode 1: split over A and B;
// mode 2: issuer A alone issues only the N = 64 product; mode 3: mode 1 with 8 extra warps spinning on an
// mbarrier (like the planner's epilogue warps waiting for an accumulator).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_dual_rate tools/mma_dual_rate.cu
#include <stdio.h>
#include "../gan_mpc_b200/csrc/h16_common.cuh"
using namespace gmpc;

__global__ void __launch_bounds__(384) bench(long long* out, int mode, int reps) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar[2], never;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 384) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c003c00u;
  if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); mbar_init(&never, 1); mbar_fence_init(); }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t ring = smem_u32(sm), bop = smem_u32(sm) + 160 * 1024;
  const int nbk = 26;
  if (warp == 1 || (warp == 3 && mode >= 1 && mode != 2)) {
    const int which = warp == 1 ? 0 : 1;
    if (elect_one()) {
      const uint32_t i64 = h16_idesc(64, 0, 1), i32 = h16_idesc(32, 0, 1);
      const uint64_t a0 = umma_smem_desc(ring, H_A_LBO, H_A_SBO), b0 = umma_smem_desc(bop, H_B_LBO, H_B_SBO);
      const long long t0 = clock64();
      for (int r = 0; r < reps; ++r) {
#pragma unroll
        for (int i = 0; i < nbk; ++i) {
          const int b = i / 13, j = i % 13, slot = i % 18;
          const uint64_t ah = a0 + (uint64_t)((slot * 8192) >> 4), bd = b0 + (uint64_t)((j * H_B_KSTEP) >> 4);
          if (which == 0) {
            umma_f16(tb + b * 96, ah, bd, i64, 1u);
            if (mode == 0) umma_f16(tb + b * 96 + 64, ah + 256, bd, i32, 1u);
          } else {
            umma_f16(tb + b * 96 + 64, ah + 256, bd, i32, 1u);
          }
        }
      }
      umma_commit(&bar[which]);
      mbar_wait(&bar[which], 0);
      out[blockIdx.x * 2 + which] = clock64() - t0;
    }
    __syncwarp();
  } else if (warp >= 4 && mode >= 4) {
    // mode 4: 8 warps read the accumulators with tcgen05.ld in a loop (epilogue-like TMEM traffic);
    // mode 5: the same warps do 16-byte shared-memory stores in a loop (epilogue-like STS traffic)
    const uint32_t t_lane = (uint32_t)((warp & 3) * 32) << 16;
    float acc = 0.f;
    uint4* dst = reinterpret_cast<uint4*>(sm + 100 * 1024) + (tid - 128);
    while (!mbar_try_wait(&bar[0], 0)) {
      if (mode == 4 || mode == 6) {
#pragma unroll
        for (int k = 0; k < 6; ++k) {
          uint32_t d[16];
          tmem_ld16_issue(tb + t_lane + (mode == 4 ? 256 : 0) + k * 16, d);
          tmem_ld_wait();
          acc += __uint_as_float(d[3]);
        }
      } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) dst[k * 256] = make_uint4(1, 2, 3, 4);
      }
    }
    if (acc == 123.f) out[1000] = 1;
  } else if (warp >= 4 && mode == 3) {
    // spinning waiters: leave when issuer A is done (bar[0] completes), polling `never` in between
    while (!mbar_try_wait(&bar[0], 0)) { (void)mbar_try_wait(&never, 0); }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tb, 512); }
}

int main() {
  long long* d; cudaMalloc(&d, 4096 * 8);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  const char* names[] = {"one issuer, both products", "two issuers (N=64 | N=32)", "one issuer, N=64 only", "two issuers + 8 warps polling an mbarrier", "two issuers + 8 warps in a tcgen05.ld loop", "two issuers + 8 warps in an STS.128 loop", "two issuers + 8 warps tcgen05.ld on the accumulator columns"};
  for (int mode = 0; mode < 7; ++mode) {
    cudaMemset(d, 0, 4096 * 8);
    bench<<<128, 384, 200 * 1024>>>(d, mode, 200);
    if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
    long long h[256]; cudaMemcpy(h, d, 256 * 8, cudaMemcpyDeviceToHost);
    long long mx = 0; for (int i = 0; i < 256; ++i) mx = h[i] > mx ? h[i] : mx;
    printf("%-45s: %6.1f cycles per block-k-step\n", names[mode], (double)mx / (200.0 * 26));
  }
  return 0;
}
