// smallgemm.cuh -- the plain contractions left around the fused kernels, on the fp32 CUDA cores:
//   * dW_l = act_l cot_l^T of the dynamics fit (norm/dynamics_trainer.py:64-79: jax.grad of the batch-mean
//     predict_loss w.r.t. the Dense kernels), gmpc_gemm_nt;
//   * the cost-MLP part of cost_vjp (policy/optimizers.py:93-105: grad_theta of  w2 d/de |f(x_T + e dx_T)|^2),
//     gmpc_cost_mixed_vjp: a primal and a tangent forward pass through the ReLU masks, two cotangents back,
//     weight gradients reduced over the batch.
// These are a few hundred MFLOP per trainer step (plumbing next to the planner), so one deterministic tiled
// kernel with generic strides and a fused epilogue serves all of them: 64 x 64 output tile per CTA, 4 x 4 per
// thread; a long reduction with few output tiles (dW of the dynamics fit) is cut into slices over gridDim.z whose
// partial products are summed in slice order, so the result stays bitwise reproducible.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmpc {

enum { SG_EPI_NONE = 0, SG_EPI_BIAS_RELU_MASK = 1, SG_EPI_MASK = 2, SG_EPI_BIAS = 3 };

struct SGemm {
  const float* A;   // A(m, r) = A[m * sam + r * sar]
  const float* B;   // B(r, n) = B[r * sbr + n * sbn]
  float* C;         // C(m, n) = C[m * ldc + n]
  const float* bias;  // [N]   (epilogues 1, 3)
  float* mask;      // [M, ldc] 0 / 1: written by epilogue 1, read by epilogue 2
  long long sam, sar, sbr, sbn;
  int M, N, ldc;
  long long R;
  float alpha;
  int accumulate;   // C += alpha A B  instead of  C = alpha A B  (epilogue 0 only)
  long long rslice; // > 0: blockIdx.z takes the reduction range [z rslice, (z + 1) rslice) and writes its raw partial
  float* part;      //      product to part[z][M][ldc]; small_gemm_reduce_kernel sums the slices in a fixed order
};

template <int EPI>
__global__ void __launch_bounds__(256) small_gemm_kernel(const SGemm G) {
  constexpr int TM = 64, TN = 64, TK = 16;
  __shared__ float As[TK][TM + 4], Bs[TK][TN + 4];
  const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * TM, n0 = blockIdx.x * TN;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const long long rbeg = G.rslice > 0 ? (long long)blockIdx.z * G.rslice : 0;
  const long long rend = G.rslice > 0 ? (rbeg + G.rslice < G.R ? rbeg + G.rslice : G.R) : G.R;
  for (long long r0 = rbeg; r0 < rend; r0 += TK) {
    // 16 x 64 elements of each operand, 4 per thread; the faster-varying thread index follows the operand's
    // unit-stride dimension when it has one
    for (int e = tid; e < TK * TM; e += 256) {
      int kk, mm;
      if (G.sar == 1) { kk = e % TK; mm = e / TK; } else { mm = e % TM; kk = e / TM; }
      const long long r = r0 + kk;
      const int m = m0 + mm;
      As[kk][mm] = (r < rend && m < G.M) ? G.A[m * G.sam + r * G.sar] : 0.f;
    }
    for (int e = tid; e < TK * TN; e += 256) {
      int kk, nn;
      if (G.sbr == 1) { kk = e % TK; nn = e / TK; } else { nn = e % TN; kk = e / TN; }
      const long long r = r0 + kk;
      const int n = n0 + nn;
      Bs[kk][nn] = (r < rend && n < G.N) ? G.B[r * G.sbr + n * G.sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < TK; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= G.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= G.N) continue;
      const size_t ix = (size_t)m * G.ldc + n;
      if (EPI == SG_EPI_NONE && G.rslice > 0) {   // one slice of a split reduction: the raw partial product
        G.part[(size_t)blockIdx.z * G.M * G.ldc + ix] = acc[i][j];
        continue;
      }
      float v = G.alpha * acc[i][j];
      if (EPI == SG_EPI_NONE) {
        if (G.accumulate) v += G.C[ix];
      } else if (EPI == SG_EPI_BIAS_RELU_MASK) {
        v += G.bias[n];
        const float mk = v > 0.f ? 1.f : 0.f;   // jax relu: derivative 0 at exactly 0
        G.mask[ix] = mk;
        v *= mk;
      } else if (EPI == SG_EPI_MASK) {
        v *= G.mask[ix];
      } else {
        v += G.bias[n];
      }
      G.C[ix] = v;
    }
  }
}

// C = alpha * sum_z part[z] (+ C): the slices of a split reduction, summed in slice order (deterministic)
__global__ void small_gemm_reduce_kernel(const float* __restrict__ part, int slices, long long count, float alpha,
                                         int accumulate, float* __restrict__ C) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    for (int z = 0; z < slices; ++z) s += part[(size_t)z * count + i];
    C[i] = accumulate ? fmaf(alpha, s, C[i]) : alpha * s;
  }
}

// Slices a long reduction is cut into when the output has too few tiles to fill the machine (dW of the dynamics
// fit: 200 x 200 outputs, thousands of columns): 0 = no split.
inline int small_gemm_slices(int M, int N, long long R) {
  const long long tiles = (long long)((M + 63) / 64) * ((N + 63) / 64);
  if (R < 1024 || tiles >= 64) return 0;
  long long s = R / 256;
  if (s > 16) s = 16;
  return s < 2 ? 0 : (int)s;
}

// `part`: workspace of small_gemm_slices(M, N, R) * M * ldc floats (epilogue NONE with ldc == N only), or nullptr.
inline cudaError_t small_gemm(SGemm G, int epi, cudaStream_t st, float* part = nullptr) {
  if (G.M <= 0 || G.N <= 0) return cudaSuccess;
  const int slices = (epi == SG_EPI_NONE && part != nullptr && G.ldc == G.N) ? small_gemm_slices(G.M, G.N, G.R) : 0;
  if (slices > 0) {
    G.rslice = ((G.R + slices - 1) / slices + 15) / 16 * 16;
    G.part = part;
    const dim3 grid3((G.N + 63) / 64, (G.M + 63) / 64, slices);
    small_gemm_kernel<SG_EPI_NONE><<<grid3, 256, 0, st>>>(G);
    const long long count = (long long)G.M * G.N;
    small_gemm_reduce_kernel<<<(int)((count + 255) / 256), 256, 0, st>>>(part, slices, count, G.alpha, G.accumulate, G.C);
    return cudaGetLastError();
  }
  G.rslice = 0;
  const dim3 grid((G.N + 63) / 64, (G.M + 63) / 64);
  switch (epi) {
    case SG_EPI_NONE: small_gemm_kernel<SG_EPI_NONE><<<grid, 256, 0, st>>>(G); break;
    case SG_EPI_BIAS_RELU_MASK: small_gemm_kernel<SG_EPI_BIAS_RELU_MASK><<<grid, 256, 0, st>>>(G); break;
    case SG_EPI_MASK: small_gemm_kernel<SG_EPI_MASK><<<grid, 256, 0, st>>>(G); break;
    default: small_gemm_kernel<SG_EPI_BIAS><<<grid, 256, 0, st>>>(G); break;
  }
  return cudaGetLastError();
}

// out[n] = alpha * sum_r X(n, r),  X(n, r) = X[n * sn + r * sr]: one warp per n, fixed summation order.
__global__ void strided_sum_kernel(const float* __restrict__ X, long long sn, long long sr, int N, long long R, float alpha,
                                   float* __restrict__ out) {
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (n >= N) return;
  float s = 0.f;
  for (long long r = lane; r < R; r += 32) s += X[n * sn + r * sr];
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xFFFFFFFFu, s, o);
  if (lane == 0) out[n] = alpha * s;
}

// c = s y_tan, d = s y  with  s = 2 sigmoid(mpc_weights[2]) scale  (the two cotangents of cost_vjp's cost-MLP part)
__global__ void mixed_vjp_seed_kernel(const float* __restrict__ y, const float* __restrict__ dy, const float* __restrict__ mpcw,
                                      float scale, long long count, float* __restrict__ c, float* __restrict__ d) {
  const float s = 2.f * scale / (1.f + expf(-mpcw[2]));
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < count; i += (long long)gridDim.x * blockDim.x) {
    c[i] = s * dy[i];
    d[i] = s * y[i];
  }
}

}  // namespace gmpc
