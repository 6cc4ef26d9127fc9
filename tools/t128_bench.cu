// t128_bench.cu -- standalone timing harness of plan_t128_kernel (no Python, no parity check: the parity
// tests are tests/test_gpu_planner.py).  Random lecun-normal weights, C2 dims by default.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo [-DGMPC_T128_TIMED] -o build/t128_bench tools/t128_bench.cu
//   ./build/t128_bench [B] [T] [iters] [hidden] [reps]
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../gan_mpc_b200/csrc/plan_t128.cuh"

using namespace gmpc;

static float* dev_rand(size_t n, float scale, unsigned seed) {
  std::vector<float> h(n);
  srand(seed);
  for (auto& v : h) v = scale * ((rand() % 20001) - 10000) / 10000.f;
  float* d;
  cudaMalloc(&d, n * 4);
  cudaMemcpy(d, h.data(), n * 4, cudaMemcpyHostToDevice);
  return d;
}

int main(int argc, char** argv) {
  const long long B = argc > 1 ? atoll(argv[1]) : 4096;
  const int T = argc > 2 ? atoi(argv[2]) : 32, iters = argc > 3 ? atoi(argv[3]) : 20;
  const int H = argc > 4 ? atoi(argv[4]) : 200, reps = argc > 5 ? atoi(argv[5]) : 3;
  gmpc_config c;
  memset(&c, 0, sizeof(c));
  c.n = 17; c.m = 6; c.T = T; c.dyn_layers = 4; c.dyn_hidden = H; c.cost_layers = 3; c.cost_hidden = 128; c.cost_fout = 10;
  int dyn_dims[MAXL + 1] = {c.n + c.m, H, H, H, c.n}, cost_dims[MAXL + 1] = {c.n, 128, 128, 10};
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  T128State S;
  int rc = t128_create(S, c, dyn_dims, cost_dims, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  if (rc || !S.supported) { printf("t128_create: rc %d supported %d (%s)\n", rc, (int)S.supported, S.why.c_str()); return 1; }
  const float *dW[4], *db[4], *cW[3], *cb[3];
  for (int l = 0; l < 4; ++l) {
    dW[l] = dev_rand((size_t)dyn_dims[l] * dyn_dims[l + 1], sqrtf(3.f / dyn_dims[l]), 11 + l);
    db[l] = dev_rand(dyn_dims[l + 1], 0.f, 1);
  }
  for (int l = 0; l < 3; ++l) {
    cW[l] = dev_rand((size_t)cost_dims[l] * cost_dims[l + 1], sqrtf(3.f / cost_dims[l]), 21 + l);
    cb[l] = dev_rand(cost_dims[l + 1], 0.f, 1);
  }
  int64_t launches = 0;
  rc = t128_set_weights(S, dW, db, cW, cb, 0, &launches);
  float mpcw_h[3] = {-2.f, 3.f, -3.f}, *mpcw;
  cudaMalloc(&mpcw, 12);
  cudaMemcpy(mpcw, mpcw_h, 12, cudaMemcpyHostToDevice);
  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.n = c.n; P.m = c.m; P.T = T; P.K = 1; P.fout = c.cost_fout;
  P.mode = MODE_PLAN; P.method = 1; P.iters = iters; P.use_cost = 1; P.final_fwd = 1;
  P.NQ = B;
  P.lr = 1e-2f; P.b1 = 0.9f; P.b2 = 0.999f; P.eps = 1e-8f;
  P.x0 = dev_rand(B * c.n, 1.f, 5);
  P.U_in = dev_rand(B * T * c.m, 1.f, 6);
  P.goal = dev_rand(B * (T + 1) * c.n, 1.f, 7);
  P.mpcw = mpcw;
  float *U, *X, *J;
  cudaMalloc(&U, B * T * c.m * 4); cudaMalloc(&X, B * (T + 1) * c.n * 4); cudaMalloc(&J, B * 4);
  P.U_out = U; P.X_out = X; P.J_out = J;
  long long* dbg;
  cudaMalloc(&dbg, 1024 * 8);
  cudaMemset(dbg, 0, 1024 * 8);
  S.d_dbg = dbg;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int r = 0; r < reps; ++r) {
    cudaEventRecord(e0);
    rc = t128_launch(S, P, 0, &launches);
    cudaEventRecord(e1);
    cudaError_t e = cudaEventSynchronize(e1);
    if (rc || e != cudaSuccess) { printf("launch rc %d: %s\n", rc, cudaGetErrorString(e)); return 2; }
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  const double mdyn = 23.0 * H + 2.0 * H * H + 17.0 * H, mcost = 17 * 128 + 128 * 128 + 1280;
  const double fl = (double)B * (iters * (4.0 * T * mdyn + 4 * mcost) + 2.0 * T * mdyn + 2 * mcost);
  printf("t128: B=%lld T=%d iters=%d H=%d: %.3f ms, %.1f k states/s, %.1f TFLOP/s algorithmic (slots %d, smem %zu)\n", B, T, iters,
         H, best, B / best, fl / best / 1e9, S.nslot, S.smem_bytes);
  long long h[1024];
  cudaMemcpy(h, dbg, sizeof(h), cudaMemcpyDeviceToHost);
  if (h[40]) {
    const double layers = (double)(iters * (2.0 * T * 4 + 6) + T * 4 + 3);
    for (int w = 0; w < 2; ++w)
      printf("  issuer %d: total %lld cycles (%.0f per layer) | wait act %lld  drained %lld  full %lld  hand-off %lld\n", w, h[40 + 8 * w],
             h[40 + 8 * w] / layers, h[41 + 8 * w], h[42 + 8 * w], h[43 + 8 * w], h[44 + 8 * w]);
    for (int s = 0; s < 2; ++s)
      printf("  epi sub%d: total %lld | wait acc %lld  hidden epilogue %lld (of which ld+drain %lld)  boundary (incl. its acc wait) %lld\n", s,
             h[8 + 8 * s], h[9 + 8 * s], h[10 + 8 * s], h[12 + 8 * s], h[11 + 8 * s]);
    const double hl = (double)(iters * (2.0 * T * 3 + 4) + T * 3 + 2);
    for (int s = 0; s < 2; ++s)
      printf("  epi sub%d per hidden layer (its ~3.3 k-steps): tcgen05.ld+wait %.0f | arithmetic %.0f | wait rel %.0f | st+wait+arrive %.0f\n", s,
             h[12 + 8 * s] / hl, h[24 + 8 * s] / hl, h[25 + 8 * s] / hl, h[26 + 8 * s] / hl);
  }
  if (h[40]) {
    const char* kn[4] = {"dyn fwd", "cost fwd", "cost bwd", "dyn bwd"};
    const double cnt[4] = {(double)T * (iters + 1), (double)iters + 1, (double)iters, (double)T * iters};
    for (int k = 0; k < 4; ++k) {
      printf("  %-8s cycles from 'part 0 of this layer complete' to the same event of the next layer:", kn[k]);
      for (int l = 0; l < 8; ++l)
        if (h[64 + 8 * k + l]) printf("  L%d %.0f", l, h[64 + 8 * k + l] / cnt[k]);
      printf("\n");
    }
  }
  if (h[40]) {
    const long long t0 = h[256];
    for (int L = 0; L < 3; ++L) {
      const long long* e = h + 256 + 128 * L;
      printf("  trace layer %d (cycles since issuer 0 reached layer 2000):\n", 2000 + L);
      for (int w = 0; w < 2; ++w) {
        printf("    issuer %d: start %lld  waits done %lld  k-steps issued:", w, e[40 * w] - t0, e[40 * w + 1] - t0);
        for (int j = 0; j < 13; ++j) if (e[40 * w + 2 + j]) printf(" %lld", e[40 * w + 2 + j] - t0);
        printf("\n");
      }
      for (int sb = 0; sb < 4; ++sb) {
        const long long* q = e + 80 + 12 * sb;
        printf("    sub %d: acc0 %lld | ld %lld pub %lld | ld %lld pub %lld || acc1 %lld | ld %lld pub %lld | ld %lld pub %lld\n", sb, q[0] - t0,
               q[1] ? q[1] - t0 : 0, q[2] ? q[2] - t0 : 0, q[3] ? q[3] - t0 : 0, q[4] ? q[4] - t0 : 0, q[6] ? q[6] - t0 : 0,
               q[7] ? q[7] - t0 : 0, q[8] ? q[8] - t0 : 0, q[9] ? q[9] - t0 : 0, q[10] ? q[10] - t0 : 0);
      }
    }
  }
  if (h[700]) {
    printf("  %lld waits timed out (tag: 1/2 drained0/1, 3 full, 4 act [+10: issuer 1]; 20-23 acc, 24 rel):\n", h[700]);
    for (int i = 0; i < 96 && i < h[700]; ++i)
      printf("    tag %lld warp %lld layer/id %lld\n", h[704 + i] >> 40, (h[704 + i] >> 32) & 255, h[704 + i] & 0xffffffffLL);
  }
  uint32_t ovf = 0;
  cudaMemcpy(&ovf, S.d_ovf, 4, cudaMemcpyDeviceToHost);
  printf("  clamped CTAs: %u\n", ovf);
  return 0;
}
