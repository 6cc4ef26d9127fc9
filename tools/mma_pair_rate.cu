// mma_pair_rate.cu -- cycles per block-k-step of the fp16-split planner's MMA pair
// (Wh x [ah|al] with N=64, then Wl x ah with N=32 into the upper half), A units walked through a
// ring exactly like plan_h16_kernel, B K-major or MN-major, with or without a concurrent bulk-copy
// stream refilling the ring (shared-memory write traffic next to the operand reads).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_pair_rate tools/mma_pair_rate.cu
#include <stdio.h>
#include "../gan_mpc_b200/csrc/h16_common.cuh"
using namespace gmpc;

struct Cfg { int b_mn, stream, reps, nslot, n_hi, n_lo, single, fence; };

__global__ void __launch_bounds__(128) bench(long long* out, const uint8_t* gsrc, Cfg c) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar, full[16], empty[16];
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0x2c003c00u ^ (i * 2654435761u & 0x03ff03ffu);
  if (tid == 0) {
    mbar_init(&bar, 1);
    for (int s = 0; s < 16; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 256);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const int NS = c.nslot;
  uint8_t* ring = sm;                      // NS x 16 KB
  uint8_t* bop = sm + 160 * 1024;          // B operand, 26 KB
  const int nbk = 26;                      // one 200x200 layer: 2 blocks x 13 k-steps
  if (warp == 2 && c.stream == 1) {             // producer: refill the ring from L2 like the planner
    uint32_t slot = 0, ph = 0;
    for (int r = 0; r < c.reps; ++r)
      for (int i = 0; i < nbk; i += 2) {
        mbar_wait(&empty[slot], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&full[slot], 16384);
          bulk_copy_g2s(ring + slot * 16384, gsrc + (size_t)(i / 2) * 16384, 16384, &full[slot]);
        }
        __syncwarp();
        if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
      }
  }
  if (warp == 1) {
    const uint32_t ihi = h16_idesc(c.n_hi, 0, c.b_mn), ilo = h16_idesc(c.n_lo, 0, c.b_mn);
    const uint64_t a0 = umma_smem_desc(smem_u32(ring), H_A_LBO, H_A_SBO);
    const uint64_t b0 = c.b_mn ? umma_smem_desc(smem_u32(bop), H_B_LBO, H_B_SBO)
                               : umma_smem_desc(smem_u32(bop), 64 * 16 + 16, 128);
    const uint32_t b_kstep = c.b_mn ? H_B_KSTEP : 2 * (64 * 16 + 16);
    const uint32_t full_a = smem_u32(&full[0]), empty_a = smem_u32(&empty[0]);
    long long t0 = clock64();
    if (c.single) {
      if (elect_one()) {
    uint32_t slot = 0, ph = 0;
    for (int r = 0; r < c.reps; ++r) {
      int b = 0, j = 0;
#pragma unroll
      for (int i = 0; i < nbk; i += 2) {
        if (c.stream == 1) { mbar_wait_a(full_a + slot * 8, ph); if (c.fence) tc_fence_after(); }
        const int b0i = b, j0 = j;
        int b1 = b0i, j1 = j0 + 1;
        if (j1 == 13) { j1 = 0; ++b1; }
        {
          const uint64_t ah = a0 + (uint64_t)((slot * 16384) >> 4);
          umma_f16(tb + b0i * 64, ah, b0 + (uint64_t)((j0 * b_kstep) >> 4), ihi, 1u);
          if (c.n_lo) umma_f16(tb + b0i * 64 + 32, ah + 256, b0 + (uint64_t)((j0 * b_kstep) >> 4), ilo, 1u);
          const uint64_t ah1 = ah + 512;
          umma_f16(tb + b1 * 64, ah1, b0 + (uint64_t)((j1 * b_kstep) >> 4), ihi, 1u);
          if (c.n_lo) umma_f16(tb + b1 * 64 + 32, ah1 + 256, b0 + (uint64_t)((j1 * b_kstep) >> 4), ilo, 1u);
          if (c.stream == 1 || c.stream == 2 || (c.stream == 3 && (i & 6) == 6)) umma_commit_a(empty_a + slot * 8);
        }
        b = b1; j = j1 + 1;
        if (j == 13) { j = 0; ++b; }
        if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
      }
    }
      }
      __syncwarp();
    } else {
    uint32_t slot = 0, ph = 0;
    for (int r = 0; r < c.reps; ++r) {
      int b = 0, j = 0;
      for (int i = 0; i < nbk; i += 2) {
        if (c.stream == 1) { mbar_wait(&full[slot], ph); if (c.fence) tc_fence_after(); }
        const int b0i = b, j0 = j;
        int b1 = b0i, j1 = j0 + 1;
        if (j1 == 13) { j1 = 0; ++b1; }
        if (elect_one()) {
          const uint64_t ah = a0 + (uint64_t)((slot * 16384) >> 4);
          umma_f16(tb + b0i * 64, ah, b0 + (uint64_t)((j0 * b_kstep) >> 4), ihi, 1u);
          if (c.n_lo) umma_f16(tb + b0i * 64 + 32, ah + 256, b0 + (uint64_t)((j0 * b_kstep) >> 4), ilo, 1u);
          const uint64_t ah1 = ah + 512;
          umma_f16(tb + b1 * 64, ah1, b0 + (uint64_t)((j1 * b_kstep) >> 4), ihi, 1u);
          if (c.n_lo) umma_f16(tb + b1 * 64 + 32, ah1 + 256, b0 + (uint64_t)((j1 * b_kstep) >> 4), ilo, 1u);
          if (c.stream) umma_commit(&empty[slot]);
        }
        __syncwarp();
        b = b1; j = j1 + 1;
        if (j == 13) { j = 0; ++b; }
        if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
      }
    }
    }
    if (elect_one()) umma_commit(&bar);
    __syncwarp();
    mbar_wait(&bar, 0);
    if ((tid & 31) == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { __syncwarp(); tmem_dealloc(tb, 256); }
}

int main() {
  long long* d; uint8_t* g;
  cudaMalloc(&d, 1024 * 8); cudaMalloc(&g, 1 << 20); cudaMemset(g, 0, 1 << 20);
  cudaFuncSetAttribute(bench, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  for (int fence = 1; fence < 2; ++fence)
    for (int stream = 0; stream < 4; ++stream)
      for (int b_mn = 1; b_mn < 2; ++b_mn)
        for (int pair = 0; pair < 3; ++pair) {
          Cfg c{b_mn, stream, 200, 9, pair == 2 ? 96 : 64, pair == 0 ? 0 : (pair == 2 ? 0 : 32), 1, fence};
          const int grid = 128, single = fence;
          bench<<<grid, 128, 200 * 1024>>>(d, g, c);
          cudaError_t e = cudaDeviceSynchronize();
          if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
          long long h[128]; cudaMemcpy(h, d, grid * 8, cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("f=%d mode=%d (0 none, 1 stream+commit, 2 commit only, 3 commit every 4th stage) B=%s %s: %6.1f cycles per block-k-step (%.0f per 200x200 layer)\n", single, stream,
                 b_mn ? "MN-major" : "K-major ", pair == 0 ? "N=64 only   " : pair == 1 ? "N=64 + N=32 " : "N=96 only   ",
                 (double)mx / (200.0 * 26), (double)mx / 200.0);
        }
  return 0;
}
