// bulk_copy_rate.cu -- L2 -> shared memory cp.async.bulk throughput per SM as a function of the
// copy size, issued by ONE thread into a ring with `depth` copies in flight (all 128 CTAs stream the
// same L2-resident 1.25 MB image, like the planner's weight stream).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/bulk_copy_rate tools/bulk_copy_rate.cu
#include <stdio.h>
#include "../gan_mpc_b200/csrc/tc_common.cuh"
using namespace gmpc;

__global__ void __launch_bounds__(64) bench(long long* out, const uint8_t* g, int bytes, int depth, int ncopies, int img, int lanes) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t full[32];
  if (threadIdx.x == 0) { for (int s = 0; s < 32; ++s) mbar_init(&full[s], 1); mbar_fence_init(); }
  __syncthreads();
  if (threadIdx.x < lanes) {
    const int ln = threadIdx.x;
    const uint32_t fa = smem_u32(full), ra = smem_u32(sm);
    const long long t0 = clock64();
    uint32_t off = ln * bytes;
    for (int i = 0; i < ncopies + depth; ++i) {
      const int s = (i % depth) * lanes + ln;
      if (i >= depth) mbar_wait_a(fa + s * 8, ((i / depth) - 1) & 1);  // copy i-depth landed: slot free
      if (i < ncopies) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fa + s * 8), "r"((uint32_t)bytes) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                     ::"r"(ra + s * bytes), "l"(g + off), "r"((uint32_t)bytes), "r"(fa + s * 8) : "memory");
        off += bytes * lanes; if (off + bytes > (uint32_t)img) off = ln * bytes;
      }
    }
    if (ln == 0) out[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  long long* d; uint8_t* g; const int img = 1280 * 1024;
  cudaMalloc(&d, 1024 * 8); cudaMalloc(&g, img); cudaMemset(g, 1, img);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int lanes : {1, 2, 4, 8})
    for (int bytes : {2048, 8192, 16384})
      for (int depth : {2, 4}) {
        const int grid = 128;
        if (bytes * depth * lanes > 192 * 1024 || depth * lanes > 32) continue;
        const int nc = 4000;
        bench<<<grid, 64, 200 * 1024>>>(d, g, bytes, depth, nc, img, lanes);
        if (cudaDeviceSynchronize() != cudaSuccess) { printf("error\n"); return 1; }
        long long h[128]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
        long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
        printf("lanes=%d copy=%5d B depth=%2d: %7.1f cycles/round  %6.1f B/cycle/SM\n", lanes, bytes, depth,
               (double)mx / nc, (double)bytes * nc * lanes / mx);
      }
  return 0;
}
