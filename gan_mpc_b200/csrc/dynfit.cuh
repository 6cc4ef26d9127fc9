// dynfit.cuh -- the dynamics trainer's loss and what its weight gradient needs
// (norm/dynamics_trainer.py:13-44 predict_loss, :47-84 train_per_update).
//
// predict_loss rolls the dynamics MLP over a window of S recorded steps -- from the recorded state
// at every step (teacher forcing) or from its own prediction -- and sums the discounted squared
// error against the recorded next states:  loss = sum_t gamma^t |x'_t - y_t|^2.
// Its gradient w.r.t. the MLP weights is  dW_l = sum_{samples, t} a_l (x) d_l  with a_l the input
// of layer l and d_l the cotangent of its output.  The sequential part -- rollout, loss, and the
// back-propagation through time of the state adjoint (only without teacher forcing does the
// adjoint of step t+1 reach step t) -- runs here, one tile of 32 windows per CTA on the layer
// machinery of plan_ffma.cuh, and writes a_l and d_l as column blocks of GEMM-ready matrices
// act_l [K_l x R], cot_l [N_l x R], R = 32 * tiles * S.  The contraction over R that remains,
// dW_l = act_l cot_l^T, is one plain GEMM per layer and is left to cuBLAS (host mirror:
// torch.matmul); db_l is the row sum of cot_l.
#pragma once
#include "plan_ffma.cuh"

namespace gmpc {

struct DynFitParams {
  PlanParams pp;       // dyn layer descriptors (DIR_DYN_F / DIR_DYN_B), n, m, hpad, NQ = windows, ntiles
  int S;               // window length
  int teacher_forcing;
  float gamma;
  const float *xseq, *useq, *yseq;  // [B,S,n], [B,S,m], [B,S,n]
  float* loss;         // [B]
  float* act[MAXL];    // layer inputs   [K_l][R]
  float* cot[MAXL];    // output cotangents [N_l][R]
  long long R;         // 32 * ntiles * S
  uint32_t* masks;     // [grid][S][(L-1)][MAXT][NTHREADS]
};

template <int MAXT>
__global__ void __launch_bounds__(NTHREADS, 1) dynfit_kernel(const __grid_constant__ DynFitParams Q) {
  const PlanParams& P = Q.pp;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, r = tid & 31;
  const int n = P.n, m = P.m, S = Q.S;
  const int n4 = (n + 3) & ~3, nm4 = (n + m + 3) & ~3;
  float* bufA = smem;
  float* bufB = bufA + P.hpad * RT;
  float* ring = bufB + P.hpad * RT;
  float* q_s = ring + NSTAGE * STAGE_FLOATS;  // [nm4] rows 0..n-1 = x, n.. = u
  float* g_s = q_s + nm4 * RT;                // [n4]  cotangent of the step's prediction
  float* dq_s = g_s + n4 * RT;                // [nm4] input cotangent of the MLP
  float* c_s = dq_s + nm4 * RT;               // [n4]  adjoint carried from step t+1 (free running)
  for (int i = tid; i < (2 * nm4 + 2 * n4) * RT; i += NTHREADS) q_s[i] = 0.f;
  const DirDesc& DF = P.dir[DIR_DYN_F];
  const DirDesc& DB = P.dir[DIR_DYN_B];
  const int L = DF.L;
  const size_t mask_layer = (size_t)MAXT * NTHREADS;
  uint32_t* maskb = Q.masks + (size_t)blockIdx.x * S * (L - 1) * mask_layer;
  const size_t R = (size_t)Q.R;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long q0 = (long long)tile * RT;
    const bool valid = (q0 + r) < P.NQ;
    const size_t col0 = (size_t)tile * S * RT;  // first column of this tile in the act/cot matrices
    __syncthreads();
    WeightPipe wp;
    wp.p = 0; wp.li = 0; wp.ci = 0; wp.issued = 0; wp.consumed = 0;
    wp.sched = SCHED_FIT;
    wp.na = wp.nb = 0;
    wp.kind = DIR_DYN_F;
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) pipe_issue(P, wp, ring, tid);
    float lossr = 0.f, disc = 1.f;
    // ------------------------------------------------------------------ forward over the window
#pragma unroll 1
    for (int t = 0; t < S; ++t) {
      // x = where(teacher_forcing, xseq[t], xprev)  (dynamics_trainer.py:28); xprev starts at xseq[0]
      for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
        const int row = e >> 5;
        const long long q = q0 + r;
        float v = 0.f;
        if (valid) {
          if (row >= n) v = Q.useq[(q * S + t) * m + (row - n)];
          else if (Q.teacher_forcing || t == 0) v = Q.xseq[(q * S + t) * n + row];
          else v = q_s[e];
        }
        q_s[e] = v;
        Q.act[0][(size_t)row * R + col0 + (size_t)t * RT + r] = v;
      }
      __syncthreads();
      {  // MLP forward, every hidden activation also stored as the next layer's input column block
        const float* in = q_s;
        bool swz_in = false;
        float acc[MAXT][8][4];
        Tiles<MAXT> tl;
#pragma unroll 1
        for (int l = 0; l < L; ++l) {
          const LayerDesc& Ly = DF.layer[l];
          tl.setup(tid, Ly.No);
          gemm_acc<MAXT>(P, Ly, in, swz_in, ring, wp, tid, tl, acc);
          if (l < L - 1) {
            float* out = (l & 1) ? bufB : bufA;
            epilogue<MAXT>(Ly, EPI_BIAS | EPI_RELU | EPI_MASK_OUT, out, true,
                           maskb + ((size_t)t * (L - 1) + l) * mask_layer + tid,
                           Q.act[l + 1] + col0 + (size_t)t * RT, tl, acc, R);
            in = out;
            swz_in = true;
          } else {
            epilogue<MAXT>(Ly, EPI_BIAS | EPI_RESID, q_s, false, nullptr, nullptr, tl, acc);
          }
        }
      }
      __syncthreads();
      // prediction x'_t is in q_s rows 0..n-1: loss and the direct part of its cotangent
      if (tid < RT) {
        float s2 = 0.f;
        for (int i = 0; i < n; ++i) {
          const float d = valid ? q_s[i * RT + r] - Q.yseq[((q0 + r) * S + t) * n + i] : 0.f;
          s2 = fmaf(d, d, s2);
          Q.cot[L - 1][(size_t)i * R + col0 + (size_t)t * RT + r] = 2.f * disc * d;
        }
        lossr = fmaf(disc, s2, lossr);
      }
      disc *= Q.gamma;
      __syncthreads();
    }
    if (tid < RT && valid) Q.loss[q0 + r] = lossr;
    // ------------------------------------------------------------------ backward over the window
    for (int e = tid; e < n4 * RT; e += NTHREADS) c_s[e] = 0.f;
    __syncthreads();
#pragma unroll 1
    for (int t = S - 1; t >= 0; --t) {
      // g_t = 2 gamma^t (x'_t - y_t) + [free running] adjoint carried from step t+1
      for (int e = tid; e < n * RT; e += NTHREADS) {
        const int i = e >> 5;
        float* cp = Q.cot[L - 1] + (size_t)i * R + col0 + (size_t)t * RT + r;
        const float g = *cp + (Q.teacher_forcing ? 0.f : c_s[e]);
        *cp = g;
        g_s[e] = g;
      }
      __syncthreads();
      {  // input-adjoint pass; the masked cotangent of every hidden layer is stored
        const float* in = g_s;
        bool swz_in = false;
        float acc[MAXT][8][4];
        Tiles<MAXT> tl;
#pragma unroll 1
        for (int lb = 0; lb < L; ++lb) {
          const LayerDesc& Ly = DB.layer[lb];
          tl.setup(tid, Ly.No);
          gemm_acc<MAXT>(P, Ly, in, swz_in, ring, wp, tid, tl, acc);
          if (lb < L - 1) {
            float* out = (lb & 1) ? bufB : bufA;
            const int l = L - 2 - lb;  // hidden layer whose relu mask gates this adjoint
            epilogue<MAXT>(Ly, EPI_MASK_IN, out, true, maskb + ((size_t)t * (L - 1) + l) * mask_layer + tid,
                           Q.cot[l] + col0 + (size_t)t * RT, tl, acc, R);
            in = out;
            swz_in = true;
          } else {
            epilogue<MAXT>(Ly, 0, dq_s, false, nullptr, nullptr, tl, acc);
          }
        }
      }
      __syncthreads();
      // adjoint of x_t for step t-1's prediction: residual + MLP input adjoint (x part)
      for (int e = tid; e < n * RT; e += NTHREADS) c_s[e] = g_s[e] + dq_s[e];
      __syncthreads();
    }
    cp_async_wait<0>();
    __syncthreads();
  }
}

}  // namespace gmpc
