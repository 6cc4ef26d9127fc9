// t128_probe.cu -- pins, on hardware, the operand images and the execution rates of the
// 128-trajectory tile (trajectories as the M dimension, activations as the A operand IN TMEM,
// weights streamed through shared memory as the B operand), for cta_group::1 and cta_group::2.
//
//   part 1 (functional): D[128 x N] = A[128 x K] * W[N x K]^T with the fp16 hi/lo split
//       D = Ah*Wh + Al*Wh + Ah*Wl, A written to TMEM with tcgen05.st (two halfs per 32-bit
//       column), W as K-major SWIZZLE_NONE core matrices, checked against a double reference.
//   part 2 (rates): cycles per k-step (three MMAs) with a static B, with the weight stream from L2
//       on all SMs at once, and with epilogue warps (tcgen05.ld -> split -> tcgen05.st) running
//       beside it; cycles of the epilogue alone.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o build/t128_probe tools/t128_probe.cu
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#include "../gan_mpc_b200/csrc/h16_common.cuh"

using namespace gmpc;

// ------------------------------------------------------------------------------------ helpers
__host__ __device__ constexpr uint32_t idesc_f16(int M, int N) {
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);  // fp16 x fp16 -> fp32, K-major A and B
}
template <int CG>
__device__ __forceinline__ void mma_ts(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  if (CG == 1)
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc),
                 "r"(idesc), "r"(acc) : "memory");
  else
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
                 "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc),
                 "r"(idesc), "r"(acc) : "memory");
}
template <int CG>
__device__ __forceinline__ void commit_to(uint32_t bar) {
  if (CG == 1)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  else
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"((uint16_t)3) : "memory");
}
template <int CG>
__device__ __forceinline__ void tmem_alloc_cg(uint32_t* holder, uint32_t ncols) {
  if (CG == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  } else {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(holder)), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tmem_dealloc_cg(uint32_t taddr, uint32_t ncols) {
  if (CG == 1)
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
  else
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool mbar_try_wait_cl(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_cl(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_cl(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_cl(bar, parity))
    if (clock64() - t0 > 2000000000LL) __trap();
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t local_bar, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_bar), "r"(cta));
  // relaxed: the data this arrival announces was written by the async proxy (bulk copy) and is read
  // by the async proxy (tcgen05.mma); a release at cluster scope costs MEMBAR.ALL.GPU (~1000 cycles)
  asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(r) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait_rx(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\t"
               "mbarrier.try_wait.parity.relaxed.cluster.shared::cta.b64 p, [%1], %2;\n\t"
               "selp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait_rx(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait_rx(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_rx(bar, parity))
    if (clock64() - t0 > 2000000000LL) __trap();
}

// B tile of one k-step (16 reduction elements) for `rows` output features: [2 k-chunks][rows][8 halfs]
__host__ __device__ inline uint32_t b_tile_off(int rows, int n, int k16) {
  return (uint32_t)((k16 >> 3) * rows * 16 + (n >> 3) * 128 + (n & 7) * 16 + (k16 & 7) * 2);
}

// ------------------------------------------------------------------------------------ part 1
// CG CTAs; CTA c owns rows [128 c, 128 c + 128) of A / D and holds output features [c NH, c NH + NH)
// of W in its shared memory (NH = N / CG).
template <int CG>
__global__ void __launch_bounds__(192) probe_kernel(const float* A, const float* W, float* D, int N, int K, int swap_halves) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int NH = N / CG, KS = K / 16;
  const uint32_t tile_b = (uint32_t)NH * 32;  // bytes of one (hi or lo) tile of one k-step
  // W hi/lo tiles: per k-step [hi tile | lo tile]
  for (int e = tid; e < NH * K; e += 192) {
    const int n = e / K, k = e - n * K;
    __half hi, lo;
    split_h1(W[(size_t)(rank * NH + n) * K + k], hi, lo);
    uint8_t* p = sm + (size_t)(k >> 4) * 2 * tile_b + b_tile_off(NH, n, k & 15);
    *reinterpret_cast<__half*>(p) = hi;
    *reinterpret_cast<__half*>(p + tile_b) = lo;
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 4) tmem_alloc_cg<CG>(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const uint32_t d_col = 0, ah_col = 256, al_col = 256 + 128;
  if (warp < 4) {  // A rows -> TMEM, lane = row
    const int r = rank * 128 + warp * 32 + lane;
    const uint32_t tl = tb + ((uint32_t)(warp * 32) << 16);
    for (int j = 0; j < KS; ++j) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float a = A[(size_t)r * K + j * 16 + 2 * c], b = A[(size_t)r * K + j * 16 + 2 * c + 1];
        if (swap_halves) split_h2(b, a, hi[c], lo[c]); else split_h2(a, b, hi[c], lo[c]);
      }
      tmem_st8(tl + ah_col + 8 * j, hi);
      tmem_st8(tl + al_col + 8 * j, lo);
    }
    tmem_st_wait();
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  if (warp == 5 && rank == 0) {
    if (elect_one()) {
      const uint32_t idesc = idesc_f16(128 * CG, N);
      for (int j = 0; j < KS; ++j) {
        const uint32_t s0 = smem_u32(sm) + j * 2 * tile_b;
        const uint64_t bh = umma_smem_desc(s0, (uint32_t)NH * 16, 128);
        const uint64_t bl = umma_smem_desc(s0 + tile_b, (uint32_t)NH * 16, 128);
        mma_ts<CG>(tb + d_col, tb + ah_col + 8 * j, bh, idesc, j > 0 ? 1u : 0u);
        mma_ts<CG>(tb + d_col, tb + al_col + 8 * j, bh, idesc, 1u);
        mma_ts<CG>(tb + d_col, tb + ah_col + 8 * j, bl, idesc, 1u);
      }
      commit_to<CG>(smem_u32(&bar));
    }
    __syncwarp();
  }
  mbar_wait(&bar, 0);
  tc_fence_after();
  if (warp < 4) {
    const int r = rank * 128 + warp * 32 + lane;
    for (int c0 = 0; c0 < N; c0 += 16) {
      float v[16];
      tmem_ld16(tb + ((uint32_t)(warp * 32) << 16) + d_col + c0, v);
      for (int i = 0; i < 16; ++i) D[(size_t)r * N + c0 + i] = v[i];
    }
  }
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 4) tmem_dealloc_cg<CG>(tb, 512);
}

template <int CG>
static int run_probe(int N, int K) {
  const int M = 128 * CG;
  std::vector<float> A((size_t)M * K), W((size_t)N * K), D((size_t)M * N);
  std::vector<double> R((size_t)M * N);
  srand(7);
  auto rnd = [] { return (float)((rand() % 200001) - 100000) / 7919.f; };
  for (auto& v : A) v = rnd();
  for (auto& v : W) v = rnd() * 0.1f;
  double ref = 0;
  for (int r = 0; r < M; ++r)
    for (int n = 0; n < N; ++n) {
      double s = 0;
      for (int k = 0; k < K; ++k) s += (double)A[(size_t)r * K + k] * W[(size_t)n * K + k];
      R[(size_t)r * N + n] = s;
      ref = fmax(ref, fabs(s));
    }
  float *dA, *dW, *dD;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dW, W.size() * 4); cudaMalloc(&dD, D.size() * 4);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dW, W.data(), W.size() * 4, cudaMemcpyHostToDevice);
  const int smem = (N / CG) * K * 4 + 1024;
  cudaFuncSetAttribute(probe_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int rc = 2;
  for (int swap = 0; swap < 2 && rc != 0; ++swap) {
    cudaMemset(dD, 0, D.size() * 4);
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(CG); cfg.blockDim = dim3(192); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, probe_kernel<CG>, (const float*)dA, (const float*)dW, dD, N, K, swap);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("probe cg=%d N=%d K=%d: CUDA error %s\n", CG, N, K, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double mx = 0;
    for (size_t i = 0; i < D.size(); ++i) mx = fmax(mx, fabs((double)D[i] - R[i]));
    const bool ok = mx <= 3e-6 * ref;
    printf("probe cg=%d N=%d K=%d halves %s: max |D - ref| = %.3g (max |ref| = %.3g, rel %.2e) -> %s\n", CG, N, K,
           swap ? "swapped" : "low=even k", mx, ref, mx / ref, ok ? "OK" : "MISMATCH");
    if (ok) rc = 0;
    else
      for (int r = 0; r < 2; ++r) { for (int n = 0; n < 6; ++n) printf(" %10.4f/%10.4f", D[(size_t)r * N + n], R[(size_t)r * N + n]); printf("\n"); }
  }
  cudaFree(dA); cudaFree(dW); cudaFree(dD);
  return rc;
}

// ------------------------------------------------------------------------------------ part 2
// A layer chain like the planner's: `layers` layers of KS k-steps; per k-step one ring slot =
// [Wh tile | Wl tile] of NH rows, three MMAs.  Warp 0: producers (4 lanes), warp 1: issuer (leader) /
// relay (peer), warps 2-9: epilogue emulation.
struct RateCfg {
  int N, KS, layers, nslot, mode;  // mode bit0: weight stream, bit1: epilogue beside it, bit2: epilogue only
  const uint8_t* wimg;             // weight image (>= layers * KS * slot bytes, cycled)
  long long img_slots;
  long long* out;                  // [grid][4]
};

template <int CG>
__global__ void __launch_bounds__(576, 1) rate_kernel(const RateCfg c) {
  extern __shared__ __align__(128) uint8_t sm[];
  __shared__ __align__(8) uint64_t full_bar[16], empty_bar[16], done_bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int rank = CG == 2 ? (int)cluster_ctarank() : 0;
  const int NH = c.N / CG;
  const uint32_t tile_b = (uint32_t)NH * 32, slot_b = 2 * tile_b;
  const int NS = c.nslot;
  const bool stream = c.mode & 1, epi = c.mode & 2, epi_only = c.mode & 4;
  for (uint32_t i = tid * 4; i < (uint32_t)NS * slot_b; i += 576 * 4) *reinterpret_cast<uint32_t*>(sm + i) = 0x3C003C00u;
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], (CG == 2 && rank == 0) ? 2 : 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(&done_bar, 1);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 2) tmem_alloc_cg<CG>(&tmem_base_s, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  tc_fence_after();
  const uint32_t tb = tmem_base_s;
  const long long total = (long long)c.layers * c.KS;
  const long long t0 = clock64();
  long long t_me = 0;
  if (warp == 0) {
    if (stream && !epi_only && lane < 4) {
      // lane k owns ring groups k, k+4, ...
      uint32_t slot = lane % NS, ph = (lane / NS) & 1;
      uint32_t src_slot = (uint32_t)((blockIdx.x / CG) * 7 + lane) % (uint32_t)c.img_slots;
      for (long long g = lane; g < total; g += 4) {
        mbar_wait_a(smem_u32(&empty_bar[slot]), ph ^ 1);
        const uint8_t* src = c.wimg + (size_t)(src_slot * CG + rank) * slot_b;
        src_slot += 4;
        if (src_slot >= (uint32_t)c.img_slots) src_slot -= (uint32_t)c.img_slots;
        mbar_arrive_expect_tx(&full_bar[slot], slot_b);
        bulk_copy_g2s(sm + (size_t)slot * slot_b, src, slot_b, &full_bar[slot]);
        slot += 4;
        while (slot >= (uint32_t)NS) { slot -= NS; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (epi_only) {
    } else if (rank == 0) {
      if (elect_one()) {
        const uint32_t idesc = idesc_f16(128 * CG, c.N);
        const uint32_t ah = tb + 256, al = tb + 360;  // 13 k-steps x 8 columns each
        const uint64_t d0 = umma_smem_desc(smem_u32(sm), (uint32_t)NH * 16, 128);
        const uint32_t d_hi = (uint32_t)(d0 >> 32), d_lo0 = (uint32_t)d0;
        const uint32_t full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]);
        uint32_t slot = 0, ph = 0;
        for (int l = 0; l < c.layers; ++l) {
#pragma unroll 1
          for (int j = 0; j < c.KS; ++j) {
            if (stream) {
              if (CG == 2) mbar_wait_rx(full0 + slot * 8, ph); else mbar_wait_a(full0 + slot * 8, ph);
              tc_fence_after();
            }
            const uint32_t lo = d_lo0 + ((slot * slot_b) >> 4);
            const uint64_t bh = ((uint64_t)d_hi << 32) | lo;
            const uint64_t bl = ((uint64_t)d_hi << 32) | (lo + (tile_b >> 4));
            mma_ts<CG>(tb, ah + 8 * j, bh, idesc, j > 0 ? 1u : 0u);
            mma_ts<CG>(tb, al + 8 * j, bh, idesc, 1u);
            mma_ts<CG>(tb, ah + 8 * j, bl, idesc, 1u);
            if (stream) commit_to<CG>(empty0 + slot * 8);
            if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
          }
        }
        commit_to<CG>(smem_u32(&done_bar));
      }
      __syncwarp();
    } else if (stream) {
      // peer CTA: relay "my half of slot s has landed" to the leader's full barrier
      if (lane == 0) {
        uint32_t slot = 0, ph = 0;
        for (long long g = 0; g < total; ++g) {
          mbar_wait_a(smem_u32(&full_bar[slot]), ph);
          mbar_arrive_remote(smem_u32(&full_bar[slot]), 0);
          if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
        }
      }
      __syncwarp();
    }
  } else if (epi || epi_only) {
    // epilogue emulation: every lane quadrant is served by four warps, each owning every fourth
    // 16-column chunk (k-step) of the layer output: all chunks are loaded to registers first (D is then
    // free for the next layer's MMAs), then chunk by chunk (fma, relu, mask bit, hi/lo split) -> 2 x tcgen05.st x8
    const int q = warp & 3, sub = (warp - 2) >> 2;  // sub 0..3
    const uint32_t tl = tb + ((uint32_t)(q * 32) << 16);
    const int nch = c.N / 16;
    float bias = 0.01f * lane, sc = 1.0009765625f;
    uint32_t mask_acc = 0;
    const long long e0 = clock64();
    for (int l = 0; l < c.layers; ++l) {
      uint32_t d[4][16];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = sub + 4 * i;
        if (ch < nch) tmem_ld16_issue(tl + (epi_only ? (uint32_t)(ch * 16) : 464u + 16u * (i & 1)), d[i]);
      }
      tmem_ld_wait();
      uint32_t mw = 0;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int ch = sub + 4 * i;
        if (ch < nch) {
          uint32_t hi[8], lo[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            float a = fmaf(__uint_as_float(d[i][2 * k]), sc, bias), b = fmaf(__uint_as_float(d[i][2 * k + 1]), sc, bias);
            mw = __funnelshift_l(__float_as_uint(0.f - a), mw, 1);
            mw = __funnelshift_l(__float_as_uint(0.f - b), mw, 1);
            a = fmaxf(a, 0.f);
            b = fmaxf(b, 0.f);
            split_h2(a, b, hi[k], lo[k]);
          }
          const uint32_t acol = epi_only ? 256u + (uint32_t)(ch * 8) : 496u;
          tmem_st8(tl + acol, hi);
          tmem_st8(tl + acol + (epi_only ? 104u : 8u), lo);
        }
      }
      tmem_st_wait();
      mask_acc ^= mw;
    }
    t_me = clock64() - e0;
    if (mask_acc == 0x12345678u) c.out[0] = 1;  // keep the work alive
  }
  if (!epi_only && rank == 0 && warp == 1) {
    mbar_wait(&done_bar, 0);
    if (lane == 0) c.out[blockIdx.x * 4 + 0] = clock64() - t0;
  }
  if ((epi || epi_only) && warp == 2 && lane == 0) c.out[blockIdx.x * 4 + 1] = t_me;
  tc_fence_before();
  __syncthreads();
  if (CG == 2) cluster_sync_all();
  if (warp == 2) tmem_dealloc_cg<CG>(tb, 512);
}

template <int CG>
static void run_rate(int N, int KS, int layers, int nslot, int mode, const uint8_t* wimg, long long img_slots, int grid) {
  static long long* d = nullptr;
  if (!d) cudaMalloc(&d, 4096 * sizeof(long long));
  cudaMemset(d, 0, 4096 * sizeof(long long));
  RateCfg c{N, KS, layers, nslot, mode, wimg, img_slots, d};
  const int smem = nslot * (N / CG) * 64 + 256;
  cudaFuncSetAttribute(rate_kernel<CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(576); cfg.dynamicSmemBytes = smem;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr; cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, rate_kernel<CG>, c);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("rate cg=%d N=%d mode=%d: CUDA error %s\n", CG, N, mode, cudaGetErrorString(e)); exit(3); }
  long long h[4096];
  cudaMemcpy(h, d, 4096 * sizeof(long long), cudaMemcpyDeviceToHost);
  long long mx = 0, mn = 1LL << 60, emx = 0;
  for (int i = 0; i < grid; ++i) {
    if (h[i * 4]) { mx = h[i * 4] > mx ? h[i * 4] : mx; mn = h[i * 4] < mn ? h[i * 4] : mn; }
    emx = h[i * 4 + 1] > emx ? h[i * 4 + 1] : emx;
  }
  const double ks = (double)layers * KS;
  printf("rate cg=%d grid=%3d N=%3d KS=%2d slots=%2d mode=%d (%s%s%s): cycles per k-step max %.1f min %.1f | epilogue cycles per layer %.0f\n",
         CG, grid, N, KS, nslot, mode, (mode & 1) ? "stream " : "static ", (mode & 2) ? "+epilogue " : "", (mode & 4) ? "epilogue-only" : "",
         mx / ks, (mn == (1LL << 60) ? 0 : mn) / ks, (double)emx / layers);
}

int main(int argc, char** argv) {
  int rc = 0;
  rc |= run_probe<1>(208, 208);
  rc |= run_probe<1>(32, 208);
  rc |= run_probe<1>(208, 32);
  rc |= run_probe<1>(256, 128);
  rc |= run_probe<2>(224, 208);
  rc |= run_probe<2>(32, 208);
  rc |= run_probe<2>(256, 256);
  if (argc > 1 && atoi(argv[1]) == 0) return rc;
  // weight image: 4 MB in L2 (bigger than any SM's shared memory, far smaller than L2)
  const size_t img_bytes = 4u << 20;
  uint8_t* wimg;
  cudaMalloc(&wimg, img_bytes);
  cudaMemset(wimg, 0x3C, img_bytes);
  int nsm = 148;
  cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  const int layers = 400;
  for (int N : {208, 224, 256}) {
    const int KS = 13;
    const long long slots1 = (long long)(img_bytes / (N * 64));
    for (int mode : {0, 1, 3}) {
      for (int nslot : {6, 10}) {
        if (mode == 0 && nslot != 6) continue;
        if (N % 16 == 0) run_rate<1>(N, KS, layers, nslot, mode, wimg, slots1, nsm);
        if (N % 32 == 0) run_rate<2>(N, KS, layers, nslot, mode, wimg, slots1, nsm / 2 * 2);
      }
    }
    if (N % 16 == 0) run_rate<1>(N, KS, 200, 6, 4, wimg, slots1, nsm);
  }
  // one SM alone (no L2 contention) for reference
  run_rate<1>(208, 13, layers, 10, 1, wimg, (long long)(img_bytes / (208 * 64)), 1);
  run_rate<2>(224, 13, layers, 10, 1, wimg, (long long)(img_bytes / (224 * 64)), 2);
  return rc;
}
