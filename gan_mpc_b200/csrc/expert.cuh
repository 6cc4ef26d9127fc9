// expert.cuh -- the expert proposal network that feeds the planner: goal states and initial
// actions for every start state (policy/eval.py:87-107, policy/base.py:40-61).
//
// Restates expert/nn.py:22-60 (StackedMLPCell / LSTMCell: teacher-forcing select, trunk, two MLPCell
// heads: next_x = head_x(y) + x, u = tanh(head_u(y))), the scans of :63-131, and the call sequence of
// expert/expert_model.py:60-91: the LSTM carry is warmed up on the history rows with teacher
// forcing (get_history_carry), then the cell runs free for T steps from the last observed state
// (get_carry_next_state_and_action_seq with teacher_forcing=False); row 0 of the goal sequence is
// that state.  One CTA per start state (looping), thread j owns output column j of every layer
// (weights [in,out] row-major are read coalesced from L2); like the critic, the step is a chain of
// dependent matvecs and is latency-bound -- it runs once per plan, in front of the planner kernel.
#pragma once
#include "common.cuh"
#include "critic.cuh"  // sigmoidf_

namespace gmpc {

struct ExpertDims {
  int n, m, T, hist;   // state / action size, horizon, history rows in front of the current state
  int F;               // LSTM features (0: MLP trunk Dense(H)+relu)
  int H;               // hidden width of the heads (and of the MLP trunk)
  int Lh;              // Dense layers per head (LSTM: num_layers; MLP: num_layers - 1)
  int Y;               // trunk output width (F or H)
  long long P;
  long long oWi, oWh, obh;          // LSTM: Wi[n,4F] | Wh[F,4F] | bh[4F]   (gates i,f,g,o)
  long long oD0, ob0;               // MLP trunk: D0[n,H] | b0[H]
  long long oHk[2][MAXL], oHb[2][MAXL];  // head 0 (next_x) / head 1 (u): kernels [in,out], biases
};

__global__ void expert_kernel(const ExpertDims D, const float* __restrict__ history,
                              const float* __restrict__ prm, long long B, float* goal, float* useq) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x;
  const int n = D.n, m = D.m, F = D.F, G = 4 * D.F, H = D.H, Y = D.Y;
  float* x = sm;               // [n]
  float* h = x + n;            // [F]
  float* c = h + F;            // [F]
  float* gates = c + F;        // [4F]
  float* y = gates + G;        // [Y]   trunk output
  float* a0 = y + Y;           // [H]
  float* a1 = a0 + H;          // [H]
  float* nx = a1 + H;          // [n]
  const int rows = D.hist + 1;

  for (long long s = blockIdx.x; s < B; s += gridDim.x) {
    __syncthreads();
    for (int i = tid; i < F; i += blockDim.x) { h[i] = 0.f; c[i] = 0.f; }
    __syncthreads();
    // steps r < hist: teacher forcing on the history rows (only the LSTM carry survives,
    // expert_model.py:66-73); steps r >= hist: free running from the last observed state
    for (int r = 0; r < D.hist + D.T; ++r) {
      const bool warm = r < D.hist;
      if (warm || r == D.hist) {
        for (int i = tid; i < n; i += blockDim.x) {
          const float v = history[(s * rows + min(r, D.hist)) * n + i];
          x[i] = v;
          if (!warm) goal[(s * (D.T + 1)) * n + i] = v;  // next_xseq = vstack([xseq[0], ...])
        }
        __syncthreads();
      }
      if (F > 0) {  // OptimizedLSTMCell (expert/nn.py:53)
        if (tid < G) {
          float z = prm[D.obh + tid];
          for (int i = 0; i < n; ++i) z = fmaf(x[i], __ldg(prm + D.oWi + (size_t)i * G + tid), z);
          for (int i = 0; i < F; ++i) z = fmaf(h[i], __ldg(prm + D.oWh + (size_t)i * G + tid), z);
          const bool is_g = (tid >= 2 * F) && (tid < 3 * F);
          gates[tid] = is_g ? tanhf(z) : sigmoidf_(z);
        }
        __syncthreads();
        if (tid < F) {
          const float cc = gates[F + tid] * c[tid] + gates[tid] * gates[2 * F + tid];
          c[tid] = cc;
          const float hh = gates[3 * F + tid] * tanhf(cc);
          h[tid] = hh;
          y[tid] = hh;
        }
        __syncthreads();
        if (warm) continue;
      } else {
        if (warm) continue;  // the MLP cell has no state besides x
        if (tid < H) {      // y = relu(Dense(H)(x)) (expert/nn.py:31)
          float z = prm[D.ob0 + tid];
          for (int i = 0; i < n; ++i) z = fmaf(x[i], __ldg(prm + D.oD0 + (size_t)i * H + tid), z);
          y[tid] = fmaxf(z, 0.f);
        }
        __syncthreads();
      }
      // two MLPCell heads (expert/nn.py:10-19)
      for (int hd = 0; hd < 2; ++hd) {
        const float* in = y;
        int din = Y;
        for (int l = 0; l < D.Lh - 1; ++l) {
          float* out = (l & 1) ? a1 : a0;
          if (tid < H) {
            float z = prm[D.oHb[hd][l] + tid];
            for (int i = 0; i < din; ++i) z = fmaf(in[i], __ldg(prm + D.oHk[hd][l] + (size_t)i * H + tid), z);
            out[tid] = fmaxf(z, 0.f);
          }
          __syncthreads();
          in = out;
          din = H;
        }
        const int dout = hd == 0 ? n : m;
        if (tid < dout) {
          float z = prm[D.oHb[hd][D.Lh - 1] + tid];
          for (int i = 0; i < din; ++i)
            z = fmaf(in[i], __ldg(prm + D.oHk[hd][D.Lh - 1] + (size_t)i * dout + tid), z);
          const int t = r - D.hist;
          if (hd == 0) {
            z += x[tid];                                   // next_x = head_x(y) + x
            nx[tid] = z;
            goal[(s * (D.T + 1) + t + 1) * n + tid] = z;
          } else {
            useq[(s * D.T + t) * m + tid] = tanhf(z);      // u = tanh(head_u(y))
          }
        }
        __syncthreads();
      }
      for (int i = tid; i < n; i += blockDim.x) x[i] = nx[i];  // carry: xprev = next_x
      __syncthreads();
    }
  }
}

}  // namespace gmpc
