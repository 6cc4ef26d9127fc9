// common.cuh -- shared definitions for the gmpc kernels (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace gmpc {

constexpr int RT = 32;              // trajectories per tile (one lane-row of 32 floats per feature)
constexpr int NTHREADS = 256;       // threads per CTA of the FFMA kernel
#ifndef GMPC_NSTAGE
#define GMPC_NSTAGE 3
#endif
constexpr int NSTAGE = GMPC_NSTAGE; // cp.async weight-ring depth
constexpr int STAGE_FLOATS = 4096;  // floats per ring stage (16 KB)
constexpr int MAXL = 8;             // max Dense layers per MLP
constexpr float ALPHA = 1e-2f;      // cost/cost_model.py:22

// One Dense layer as the kernel consumes it: out[o][r] = sum_k W[k][o] * in[k][r]  (+ bias[o]).
// W is row-major [Ki][ld] with ld = round_up(No, 4), zero padded, 16-byte aligned, and is streamed
// through shared memory in chunks of kc rows.
struct LayerDesc {
  const float* W;
  const float* bias;  // nullptr for the transposed (backward) direction
  int Ki, No, ld, kc, nchunks, pad_;
};

// An MLP pass in processing order (forward: layer 0..L-1; backward: transposed layer L-1..0).
struct DirDesc {
  LayerDesc layer[MAXL];
  int L;
  int pad_;
};

enum { DIR_DYN_F = 0, DIR_COST_F = 1, DIR_COST_B = 2, DIR_DYN_B = 3, DIR_END = 4 };
enum { MODE_PLAN = 0, MODE_OBJGRAD = 1, MODE_ROLLOUT = 2, MODE_L2GRAD = 3 };

struct PlanParams {
  DirDesc dir[4];
  int n, m, T, K;
  int hpad;        // rows of each activation buffer (max hidden width, rounded up to 4)
  int fout;        // cost MLP output width
  int mode;        // MODE_*
  int method;      // GMPC_METHOD_*
  int iters;       // planning iterations with a backward sweep
  int use_cost;    // cost MLP passes present in the schedule
  int final_fwd;   // a final forward-only evaluation follows the iterations
  int ntiles;
  long long NQ;    // trajectories = B*K
  float lr, b1, b2, eps;
  const float* x0;    // [B, n]
  const float* U_in;  // [NQ, T, m]
  const float* goal;  // [B, T+1, n]  (goal states, or desired states in MODE_L2GRAD)
  const float* mpcw;  // raw mpc weights [3]
  float* U_out;       // [NQ, T, m]   (nullable)
  float* X_out;       // [NQ, T+1, n] (nullable)
  float* J_out;       // [NQ]         (nullable)
  float* dU_out;      // [NQ, T, m]   (nullable)
  float* lam_out;     // [NQ, T+1, n] (nullable)
  // per-CTA scratch slabs (L2 resident)
  float* ws_X;        // [grid][(T+1)][n][RT]
  float* ws_G;        // [grid][(T+1)][n][RT]
  float* ws_U;        // [grid][T][m][RT]
  float* ws_M;        // [grid][T][m][RT]
  float* ws_V;        // [grid][T][m][RT]
  uint32_t* ws_mask;  // [grid][T*(Ld-1) + (Lc-1)][MAXT][NTHREADS]
};

__device__ __forceinline__ void cp_async16(void* smem_ptr, const void* gptr) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_ptr);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gptr) : "memory");
}
__device__ __forceinline__ void cp_async_commit() {
  asm volatile("cp.async.commit_group;\n" ::: "memory");
}
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory");
}

}  // namespace gmpc
