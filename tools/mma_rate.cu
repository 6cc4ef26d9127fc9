// mma_rate.cu -- tcgen05.mma execution-rate microbenchmark (timing only, operand contents arbitrary).
// Decides the operand roles of the planner's contraction: cycles per MMA as a function of kind
// (tf32 / f16), M, N, A source (shared memory descriptor or TMEM) and shared-memory layout
// (SWIZZLE_NONE core matrices vs SWIZZLE_128B).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/mma_rate tools/mma_rate.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../gan_mpc_b200/csrc/tc_common.cuh"

using namespace gmpc;

__device__ __forceinline__ uint32_t make_idesc(int kind_f16, int M, int N) {
  const uint32_t fmt = kind_f16 ? 0u : 2u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss(int f16, uint32_t d, uint64_t a, uint64_t b, uint32_t id) {
  if (f16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(id), "r"(1u)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d),
                 "l"(a), "l"(b), "r"(id), "r"(1u)
                 : "memory");
}
__device__ __forceinline__ void mma_ts(int f16, uint32_t d, uint32_t a, uint64_t b, uint32_t id) {
  if (f16)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(id), "r"(1u)
                 : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d),
                 "r"(a), "l"(b), "r"(id), "r"(1u)
                 : "memory");
}

struct Cfg {
  int f16, M, N, a_tmem, swz, ksteps, reps, nacc, walk;
};

template <int F16, int ATMEM, int WALK>
__global__ void __launch_bounds__(128) bench(long long* out, Cfg c) {
  extern __shared__ __align__(1024) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 200 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(sm)[i] = 0u;
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
    const uint32_t idesc = make_idesc(c.f16, c.M, c.N);
    const uint32_t a_base = smem_u32(sm), b_base = smem_u32(sm) + 64 * 1024;
    // K bytes per MMA are 32 in both kinds.  SWIZZLE_NONE K-major: two 16-byte k-chunks LBO
    // apart, 8-row core matrices SBO = 128 apart.  SWIZZLE_128B: rows of 128 B (4 MMAs of K),
    // 8-row atoms 1024 B apart; k advance inside the atom is +32 B.
    uint64_t ad0, bd0, a_inc, b_inc;
    if (c.swz) {
      ad0 = umma_smem_desc(a_base, 16, 1024) | ((uint64_t)2 << 61);
      bd0 = umma_smem_desc(b_base, 16, 1024) | ((uint64_t)2 << 61);
      a_inc = b_inc = 32 >> 4;  // (wraps into the next atom column every 4 steps in a real kernel)
    } else {
      const uint32_t a_lbo = c.M * 16, b_lbo = c.N * 16 + 16;
      ad0 = umma_smem_desc(a_base, a_lbo, 128);
      bd0 = umma_smem_desc(b_base, b_lbo, 128);
      a_inc = (2 * a_lbo) >> 4;
      b_inc = (2 * b_lbo) >> 4;
    }
    const uint32_t a_t0 = tb + 256;  // A in TMEM: columns 256.. (8 columns of 32 bit per MMA)
    const long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < c.reps; ++r) {
        uint64_t ad = ad0, bd = bd0;
        uint32_t at = a_t0, acc = 0;
#pragma unroll 8
        for (int s = 0; s < c.ksteps; ++s) {
          if (ATMEM)
            mma_ts(F16, tb + acc, at, bd, idesc);
          else
            mma_ss(F16, tb + acc, ad, bd, idesc);
          if (WALK) {  // walk the descriptors like a real k loop (else: same operands every time)
            if (c.nacc > 1) acc = (acc + c.N >= (uint32_t)(c.nacc * c.N)) ? 0u : acc + c.N;
            ad += a_inc;
            bd += b_inc;
            at += 8;
          }
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    if ((tid & 31) == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tb, 512);
  }
}

static double run(Cfg c, int grid) {
  static long long* d = nullptr;
  if (!d) cudaMalloc(&d, 1024 * sizeof(long long));
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    kern<<<grid, 128, 200 * 1024>>>(d, c);
  };
  const int sel = c.f16 * 4 + c.a_tmem * 2 + c.walk;
  switch (sel) {
    case 0: go(bench<0, 0, 0>); break;
    case 1: go(bench<0, 0, 1>); break;
    case 2: go(bench<0, 1, 0>); break;
    case 3: go(bench<0, 1, 1>); break;
    case 4: go(bench<1, 0, 0>); break;
    case 5: go(bench<1, 0, 1>); break;
    case 6: go(bench<1, 1, 0>); break;
    default: go(bench<1, 1, 1>); break;
  }
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("  [error %s]", cudaGetErrorString(e));
    return -1;
  }
  long long h[1024];
  cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
  return (double)mx / ((double)c.ksteps * c.reps);
}

int main() {
  const int Ns[] = {16, 32, 64, 96, 104, 128, 208, 256};
  for (int walk = 0; walk < 2; ++walk) {
    const int grid = 148;
    for (int f16 = 0; f16 < 2; ++f16)
      for (int M : {64, 128})
        for (int a_tmem = 0; a_tmem < 2; ++a_tmem)
          for (int swz = 0; swz < 2; ++swz) {
            if (a_tmem && swz) continue;  // B layout tested via the SS rows
            printf("walk=%d kind=%s M=%3d A=%s layout=%s  cyc/MMA:", walk, f16 ? "f16 " : "tf32", M,
                   a_tmem ? "tmem" : "smem", swz ? "SW128" : "NONE ");
            for (int N : Ns) {
              if (M == 128 && (N % 16)) { printf("  N%-3d   n/a", N); continue; }
              // operand images must fit: ksteps * bytes per kstep <= 64 KB (A) / 128 KB (B)
              int ks = 8;
              Cfg c{f16, M, N, a_tmem, swz, ks, 400, (2 * N <= 256) ? 2 : 1, walk};
              printf("  N%-3d %6.1f", N, run(c, grid));
            }
            printf("\n");
          }
  }
  return 0;
}
