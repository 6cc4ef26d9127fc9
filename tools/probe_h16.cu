// probe_h16.cu -- pins the kind::f16 operand images of h16_common.cuh on hardware:
// D[128][64] = A[128][K] * B[64][K]^T with A K-major units and B MN-major, K = 48 (3 k-steps),
// plus the N = 32 accumulate-into-the-upper-half pattern the planner uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/probe_h16 tools/probe_h16.cu
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../gan_mpc_b200/csrc/h16_common.cuh"

using namespace gmpc;

constexpr int K = 48, KS = K / 16;

__global__ void __launch_bounds__(128) probe(const float* A, const float* Al, const float* B, float* D) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* a_img = sm;                       // KS x (hi unit, lo unit)
  uint8_t* b_img = sm + KS * 2 * H_UNIT;     // 64 columns x K features
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int e = tid; e < 128 * K; e += 128) {
    const int r = e / K, kk = e - r * K;
    const int j = kk >> 4, k16 = kk & 15;
    const uint32_t off = (k16 >> 3) * H_A_LBO + (r >> 3) * H_A_SBO + (r & 7) * 16 + (k16 & 7) * 2;
    *reinterpret_cast<__half*>(a_img + j * 2 * H_UNIT + off) = __float2half_rn(A[e]);
    *reinterpret_cast<__half*>(a_img + j * 2 * H_UNIT + H_UNIT + off) = __float2half_rn(Al[e]);
  }
  for (int e = tid; e < 64 * K; e += 128) {
    const int n = e / K, f = e - n * K;
    *reinterpret_cast<__half*>(b_img + h16_b_off(n, f)) = __float2half_rn(B[e]);
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(&tmem_base_s, 64);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  if (warp == 1) {
    if (elect_one()) {
      const uint32_t i64 = h16_idesc(64, 0, 1), i32 = h16_idesc(32, 0, 1);
      for (int j = 0; j < KS; ++j) {
        const uint64_t ah = umma_smem_desc(smem_u32(a_img) + j * 2 * H_UNIT, H_A_LBO, H_A_SBO);
        const uint64_t al = umma_smem_desc(smem_u32(a_img) + j * 2 * H_UNIT + H_UNIT, H_A_LBO, H_A_SBO);
        const uint64_t bd = umma_smem_desc(smem_u32(b_img) + j * H_B_KSTEP, H_B_LBO, H_B_SBO);
        umma_f16(tb, ah, bd, i64, j > 0 ? 1u : 0u);  // D[:, 0:64] (+)= Ah * B[0:64]^T
        umma_f16(tb + 32, al, bd, i32, 1u);           // D[:, 32:64] += Al * B[0:32]^T
      }
      umma_commit(&bar);
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  for (int c0 = 0; c0 < 64; c0 += 16) {
    float v[16];
    tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + c0, v);
    for (int i = 0; i < 16; ++i) D[(size_t)tid * 64 + c0 + i] = v[i];
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tb, 64);
}

int main() {
  std::vector<float> A(128 * K), Al(128 * K), B(64 * K), D(128 * 64), R(128 * 64);
  srand(1);
  auto rnd = [] { return (float)((rand() % 2001) - 1000) / 256.f; };  // exact in fp16
  for (auto& v : A) v = rnd();
  for (auto& v : Al) v = rnd();
  for (auto& v : B) v = rnd();
  for (int r = 0; r < 128; ++r)
    for (int n = 0; n < 64; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[r * K + k] * B[n * K + k];
      if (n >= 32)
        for (int k = 0; k < K; ++k) s += (double)Al[r * K + k] * B[(n - 32) * K + k];
      R[r * 64 + n] = (float)s;
    }
  float *dA, *dAl, *dB, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dAl, Al.size() * 4);
  cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dAl, Al.data(), Al.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = KS * 2 * H_UNIT + 64 * K * 2 + 1024;
  cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  probe<<<1, 128, smem>>>(dA, dAl, dB, dD);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
  cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
  double mx = 0, ref = 0;
  for (size_t i = 0; i < D.size(); ++i) { mx = fmax(mx, fabs(D[i] - R[i])); ref = fmax(ref, fabs(R[i])); }
  printf("probe_h16: max |D - ref| = %g (max |ref| = %g) -> %s\n", mx, ref, mx <= 1e-3 * ref ? "OK" : "MISMATCH");
  if (mx > 1e-3 * ref)
    for (int r = 0; r < 2; ++r) { for (int n = 0; n < 8; ++n) printf(" %9.3f/%9.3f", D[r * 64 + n], R[r * 64 + n]); printf("\n"); }
  return mx <= 1e-3 * ref ? 0 : 2;
}
