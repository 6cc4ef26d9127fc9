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
  if (getenv("T128_ODD")) {   // tests/util.py ODD: widths that are not multiples of 16, two-layer cost MLP
    c.n = 5; c.m = 3; c.dyn_layers = 3; c.dyn_hidden = 50; c.cost_layers = 2; c.cost_hidden = 30; c.cost_fout = 7;
    const int dd[4] = {8, 50, 50, 5}, cd[3] = {5, 30, 7};
    for (int i = 0; i < 4; ++i) dyn_dims[i] = dd[i];
    for (int i = 0; i < 3; ++i) cost_dims[i] = cd[i];
  }
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  T128State S;
  int rc = t128_create(S, c, dyn_dims, cost_dims, prop.multiProcessorCount, prop.sharedMemPerBlockOptin);
  if (rc || !S.supported) { printf("t128_create: rc %d supported %d (%s)\n", rc, (int)S.supported, S.why.c_str()); return 1; }
  const float *dW[4], *db[4], *cW[3], *cb[3];
  for (int l = 0; l < c.dyn_layers; ++l) {
    dW[l] = dev_rand((size_t)dyn_dims[l] * dyn_dims[l + 1], sqrtf(3.f / dyn_dims[l]), 11 + l);
    db[l] = dev_rand(dyn_dims[l + 1], 0.f, 1);
  }
  for (int l = 0; l < c.cost_layers; ++l) {
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
    for (int w = 0; w < 3; ++w)
      printf("  issuer %d: total %lld (%.0f per layer) | per layer: wait first block %.0f  wait other blocks %.0f  wait ring %.0f  issue+rest %.0f\n", w,
             h[40 + 8 * w], h[40 + 8 * w] / layers, h[41 + 8 * w] / layers, h[42 + 8 * w] / layers, h[43 + 8 * w] / layers,
             (h[40 + 8 * w] - h[41 + 8 * w] - h[42 + 8 * w] - h[43 + 8 * w]) / layers);
  }
  if (h[40]) {
    const double np = (double)T * (iters + 1);
    printf("  issuer 0, dyn fwd by layer (layer start -> first block available | -> last MMA issued):");
    for (int i = 0; i < 4; ++i) printf("  L%d %.0f | %.0f", i, h[100 + i] / np, h[104 + i] / np);
    printf("\n");
  }
  if (h[8]) {
    const double hl = (double)(iters * (2.0 * T * 3 + 4) + T * 3 + 2);
    for (int sb = 0; sb < 4; ++sb) {
      const long long* o = h + 8 + 8 * sb;
      printf("  epi sub%d: total %lld | wait acc %lld  hidden %lld  boundary %lld | per hidden layer: ld %.0f  alu %.0f  st %.0f  finish %.0f  (all %.0f)\n",
             sb, o[0], o[1], o[2], o[3], o[4] / hl, o[5] / hl, o[6] / hl, o[7] / hl, o[2] / hl);
    }
  }
  if (h[700]) {
    printf("  %lld waits timed out (tag: 1 operand, 3 full [+10: issuer 1]; 20 narrow acc, 22+p acc part p; 30 empty):\n", h[700]);
    for (int i = 0; i < 96 && i < h[700]; ++i)
      printf("    tag %lld warp %lld layer/id %lld\n", h[704 + i] >> 40, (h[704 + i] >> 32) & 255, h[704 + i] & 0xffffffffLL);
  }
  if (h[256 + 64]) {
    const long long t0 = h[520 + 3 * 5];   // sub 3 passes the acc wait of the first traced layer
    for (int L = 0; L < 4; ++L) {
      printf("  trace layer +%d (cycles since sub 3 saw layer +0 complete)\n", L);
      for (int w = 0; w < 3; ++w) {
        printf("    issuer %d block available:", w);
        for (int j = 0; j < 13; ++j) if (h[256 + L * 64 + w * 16 + j]) printf(" %lld", h[256 + L * 64 + w * 16 + j] - t0);
        printf("\n");
      }
      for (int sb = 0; sb < 4; ++sb) {
        const long long* e = h + 520 + L * 20 + sb * 5;
        printf("    sub %d: acc seen %lld | blocks stored:", sb, e[0] ? e[0] - t0 : 0);
        for (int i = 1; i < 5; ++i) if (e[i]) printf(" %lld", e[i] - t0);
        printf("\n");
      }
    }
  }
  if (h[700]) {
    printf("  last positions (warp: layer code; epilogue: 1 waiting acc, 2 acc passed, 16+nk boundary arrive, 32+j block j stored; issuer: 64+k waiting block k):\n   ");
    for (int w = 0; w < 20; ++w) printf(" w%d: %lld/%lld", w, h[840 + w] >> 8, h[840 + w] & 255);
    printf("\n");
  }
  uint32_t ovf = 0;
  cudaMemcpy(&ovf, S.d_ovf, 4, cudaMemcpyDeviceToHost);
  printf("  clamped CTAs: %u\n", ovf);
  return 0;
}
