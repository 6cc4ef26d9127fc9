// ilqr.cuh -- the reference's own planner step, trajax iLQR (policy/optimizers.py:10-21 with the
// options of policy/eval.py:10-20), as ONE persistent kernel on the fp32 CUDA cores.
//
// A CTA owns a tile of RT = 32 trajectories and runs the whole iLQR loop for it in lock step with
// vmap semantics (a lane whose own loop condition is false is frozen while the tile continues):
//
//   rollout -> [ linearize + adjoint scan -> continuation test -> Riccati sweep (tvlqr)
//                -> backtracking line search of feedback rollouts ]*
//
// * linearize: the dynamics Jacobians A_t = I + dMLP/dx, B_t = dMLP/du are n adjoint passes with
//   unit seeds through the ReLU masks of one forward pass (what jax.jacobian's n VJPs do); with the
//   trajectories as the 32 columns of the tile every pass is a full [features x 32] FFMA tile and
//   reuses the layer machinery of plan_ffma.cuh (weights streamed from L2 through the cp.async ring).
//   The cost Hessians are closed forms: pseudo-Huber (cost/cost_model.py:20-28) and the Gauss-Newton
//   term 2 w2 Jf^T Jf of the terminal cost MLP (cost/nn.py:23-29; exact, the MLP is piecewise linear).
// * Riccati sweep and adjoint scan: thread (lane r = trajectory, w = tid / 32) owns every 8th entry
//   of each small matrix of trajectory r; all matrices live in shared memory as [entry][32], so
//   every access is conflict free and no shuffles are needed.  This region aliases the MLP buffers.
// * G_ = G + max(0, 1e-8 - lambda_min(G)) I (trajax lqr_step): a Cholesky factorisation that succeeds
//   proves lambda_min > 0 and the shift is 0 (to within 1e-8); only when it fails is lambda_min computed
//   (cyclic Jacobi) and the shifted matrix factorised.
//
// trajax is an un-vendored dependency (requirements.txt:51); the algorithm is restated from the
// published method, the checker is oracle/ilqr.py.
#pragma once
#include "plan_ffma.cuh"

namespace gmpc {

constexpr int IL_MAXM = 16;      // action size limit of the per-lane Cholesky / Jacobi code
constexpr float IL_DELTA = 1e-8f;

struct IlqrParams {
  PlanParams pp;        // layer descriptors, sizes, x0 / U_in / goal, X_out / U_out / J_out / dU_out / lam_out
  int maxiter;
  float gthr, alpha0, alpha_min;
  int grad_lag;         // gmpc_ilqr_options::gradient_lag: returned gradient / adjoints and the continuation test lag one iterate
  int* it_out;          // [B] iterations run per trajectory
  float* A_out;         // [B,T,n,n] nullable: dynamics Jacobians at the returned trajectory (lqr[5])
  float* B_out;         // [B,T,n,m] nullable (lqr[6])
  float* ws;            // per-CTA slabs, see IlqrWs
  long long ws_stride;  // floats per CTA
  long long* stats;     // nullable: [0] outer iterations (tile level), [1] line-search rollouts
  // bilevel tail (policy/optimizers.py:59-73 for loss = L2MPC.loss), run when `desired` is set
  const float* desired; // [B,T+1,n]
  float* bl_loss;       // [B]        high-level loss L2MPC.loss(X, desired)
  float* bl_B;          // [B,T,m]    nullable: loss_grad_wrt_control
  float* bl_hess;       // [B,Tm,Tm]  nullable: cost_hessian_wrt_control
  float* bl_H;          // [B,T,m]    solve(hessian, B)
  float* bl_dxT;        // [B,n]      d x_T / dU . H  (tangent of the terminal state along H)
  float* bl_gw;         // [B,3]      d (H . grad_U J) / d mpc_weights (raw, pre-sigmoid)
  int tile_traj;        // trajectories per tile (<= 32)
  int pack_small;       // small tiles pack (trajectory, Jacobian row) pairs into the lanes
  int bl_generic;       // `desired` holds d loss / d X [B,T+1,n] of an arbitrary loss (bl_loss unused)
  const float* bl_V;    // [B,T,m]    nullable: cost_vjp's direction V given by the caller -- used instead
                        //            of H (no Hessian, no solve; bl_H returns V)
};

// per-CTA global scratch, every array is [..][RT]
struct IlqrWs {
  size_t X, Xn, G, lam, U, Un, k, grad, A, B, K, Jf, QT, qT, S, QS, H, rhs, lamP, gradP, gnP, total;
};
__host__ __device__ inline IlqrWs ilqr_ws_layout(int n, int m, int T, int fout, int bilevel) {
  IlqrWs s;
  size_t o = 0;
  const size_t sx = (size_t)(T + 1) * n * RT, su = (size_t)T * m * RT;
  s.X = o; o += sx;
  s.Xn = o; o += sx;
  s.G = o; o += sx;
  s.lam = o; o += sx;
  s.U = o; o += su;
  s.Un = o; o += su;
  s.k = o; o += su;
  s.grad = o; o += su;
  s.A = o; o += (size_t)T * n * n * RT;
  s.B = o; o += (size_t)T * n * m * RT;
  s.K = o; o += (size_t)T * m * n * RT;
  s.Jf = o; o += (size_t)fout * n * RT;
  s.QT = o; o += (size_t)n * n * RT;
  s.qT = o; o += (size_t)n * RT;
  s.lamP = o; o += sx;      // gradient_lag: adjoints / gradient / gradient norm of the iterate BEFORE the last step
  s.gradP = o; o += su;
  s.gnP = o; o += RT;
  s.S = s.QS = s.H = s.rhs = o;
  if (bilevel) {  // sensitivities d x_t / dU [n][Tm], Q S, the (Tm)^2 Hessian, the right-hand side
    const size_t TM = (size_t)T * m;
    s.S = o; o += (size_t)n * TM * RT;
    s.QS = o; o += (size_t)n * TM * RT;
    s.H = o; o += TM * TM * RT;
    s.rhs = o; o += TM * RT;
  }
  s.total = o;
  return s;
}

struct IlqrSmem {
  int un_floats;     // union of the MLP buffers and the Riccati region
  int small_floats;  // persistent small arrays
  size_t bytes;
};
__host__ __device__ inline IlqrSmem ilqr_smem_layout(int n, int m, int fout, int hpad) {
  IlqrSmem s;
  const int n4 = (n + 3) & ~3, nm4 = (n + m + 3) & ~3, f4 = (fout + 3) & ~3;
  const int s4 = n4 > f4 ? n4 : f4;
  const int mlp = 2 * hpad * RT + NSTAGE * STAGE_FLOATS;
  const int ric = (2 * n * n + 3 * m * n + 2 * m * m + 3 * n + 3 * m + 2) * RT;
  s.un_floats = mlp > ric ? mlp : ric;
  s.small_floats = (2 * nm4 + s4 + f4 + 11 + 4) * RT + 2 * RT;
  s.bytes = sizeof(float) * ((size_t)s.un_floats + s.small_floats);
  return s;
}

// smallest eigenvalue of the symmetric m x m matrix of lane r (cyclic Jacobi, rare path)
__device__ __noinline__ float il_min_eig(const float* G_s, int m, int r) {
  float a[IL_MAXM * IL_MAXM];
  for (int i = 0; i < m * m; ++i) a[i] = G_s[i * RT + r];
  for (int sweep = 0; sweep < 12; ++sweep) {
    float off = 0.f;
    for (int p = 0; p < m - 1; ++p)
      for (int q = p + 1; q < m; ++q) {
        const float apq = a[p * m + q];
        off += apq * apq;
        if (fabsf(apq) < 1e-37f) continue;
        const float theta = (a[q * m + q] - a[p * m + p]) / (2.f * apq);
        const float t = copysignf(1.f, theta) / (fabsf(theta) + sqrtf(theta * theta + 1.f));
        const float c = rsqrtf(t * t + 1.f), s = t * c;
        for (int k = 0; k < m; ++k) {
          const float akp = a[k * m + p], akq = a[k * m + q];
          a[k * m + p] = c * akp - s * akq;
          a[k * m + q] = s * akp + c * akq;
        }
        for (int k = 0; k < m; ++k) {
          const float apk = a[p * m + k], aqk = a[q * m + k];
          a[p * m + k] = c * apk - s * aqk;
          a[q * m + k] = s * apk + c * aqk;
        }
      }
    if (off < 1e-30f) break;
  }
  float mn = a[0];
  for (int i = 1; i < m; ++i) mn = fminf(mn, a[i * m + i]);
  return mn;
}

// Cholesky factor of (G + shift I) of lane r into L_s (lower triangle); false if a pivot is not > 0
__device__ __forceinline__ bool il_cholesky(const float* G_s, float* L_s, int m, int r, float shift) {
  bool ok = true;
  for (int a = 0; a < m; ++a)
    for (int b = 0; b <= a; ++b) {
      float s = G_s[(a * m + b) * RT + r] + (a == b ? shift : 0.f);
      for (int c = 0; c < b; ++c) s = fmaf(-L_s[(a * m + c) * RT + r], L_s[(b * m + c) * RT + r], s);
      if (a == b) {
        if (!(s > 0.f)) ok = false;
        L_s[(a * m + a) * RT + r] = sqrtf(s);
      } else {
        L_s[(a * m + b) * RT + r] = s / L_s[(b * m + b) * RT + r];
      }
    }
  return ok;
}

template <int MAXT>
__global__ void __launch_bounds__(NTHREADS, 1) ilqr_kernel(const __grid_constant__ IlqrParams Q) {
  const PlanParams& P = Q.pp;
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x, r = tid & 31, w = tid >> 5;
  const int n = P.n, m = P.m, T = P.T, fout = P.fout;
  const int n4 = (n + 3) & ~3, nm4 = (n + m + 3) & ~3, f4 = (fout + 3) & ~3;
  const int s4 = n4 > f4 ? n4 : f4;
  const IlqrSmem SL = ilqr_smem_layout(n, m, fout, P.hpad);
  // MLP view of the union region
  float* bufA = smem;
  float* bufB = bufA + P.hpad * RT;
  float* ring = bufB + P.hpad * RT;
  // Riccati view of the same region
  float* P_s = smem;                  // [n*n]  value-function Hessian
  float* T_s = P_s + n * n * RT;      // [n*n]  A^T P
  float* BtP_s = T_s + n * n * RT;    // [m*n]  B^T P, later H + G K
  float* H_s = BtP_s + m * n * RT;    // [m*n]  B^T P A
  float* K_s = H_s + m * n * RT;      // [m*n]  feedback gain
  float* G_s = K_s + m * n * RT;      // [m*m]
  float* L_s = G_s + m * m * RT;      // [m*m]  Cholesky factor of G_
  float* p_s = L_s + m * m * RT;      // [n]
  float* pn_s = p_s + n * RT;         // [n]
  float* d_s = pn_s + n * RT;         // [n]    x_t - goal_t
  float* u_s = d_s + n * RT;          // [m]
  float* h_s = u_s + m * RT;          // [m]
  float* k_s = h_s + m * RT;          // [m]
  float* sd_s = k_s + m * RT;         // [1]    sqrt(|d|^2 + a^2)
  float* su_s = sd_s + RT;            // [1]    sqrt(|u|^2 + a^2)
  // persistent small arrays
  float* q_s = smem + SL.un_floats;   // [nm4]  rows 0..n-1 = x, n..n+m-1 = u
  float* seed_s = q_s + nm4 * RT;     // [s4]   unit seeds of the Jacobian passes
  float* dq_s = seed_s + s4 * RT;     // [nm4]
  float* y_s = dq_s + nm4 * RT;       // [f4]
  float* obj_s = y_s + f4 * RT;
  float* objn_s = obj_s + RT;
  float* alpha_s = objn_s + RT;
  float* gpart_s = alpha_s + RT;      // [8]
  int* act_s = reinterpret_cast<int*>(gpart_s + 8 * RT);
  int* srch_s = act_s + RT;
  int* acc_s = srch_s + RT;
  int* it_s = acc_s + RT;
  int* slist_s = it_s + RT;           // [32] compacted list of the searching lanes, [32] = their number
  for (int i = tid; i < SL.small_floats; i += NTHREADS) q_s[i] = 0.f;

  const float w0 = 1.f / (1.f + expf(-P.mpcw[0]));
  const float w1 = 1.f / (1.f + expf(-P.mpcw[1]));
  const float w2 = 1.f / (1.f + expf(-P.mpcw[2]));
  const float a2 = ALPHA * ALPHA;

  const IlqrWs WL = ilqr_ws_layout(n, m, T, fout, Q.desired != nullptr);
  float* wsb = Q.ws + (size_t)blockIdx.x * Q.ws_stride;
  float *wsX = wsb + WL.X, *wsXn = wsb + WL.Xn, *wsG = wsb + WL.G, *wsLam = wsb + WL.lam;
  float *wsU = wsb + WL.U, *wsUn = wsb + WL.Un, *wsk = wsb + WL.k, *wsGrad = wsb + WL.grad;
  float *wsLamP = wsb + WL.lamP, *wsGradP = wsb + WL.gradP, *wsGnP = wsb + WL.gnP;
  float *wsA = wsb + WL.A, *wsB = wsb + WL.B, *wsK = wsb + WL.K, *wsJf = wsb + WL.Jf;
  float *wsQT = wsb + WL.QT, *wsqT = wsb + WL.qT;
  const int Ld = P.dir[DIR_DYN_F].L, Lc = P.dir[DIR_COST_F].L;
  const size_t mask_layer = (size_t)MAXT * NTHREADS;
  uint32_t* dynMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)(Ld - 1) + (Lc - 1)) * mask_layer;
  uint32_t* costMask = dynMask + (size_t)(Ld - 1) * mask_layer;
  long long n_outer = 0, n_roll = 0;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long q0 = (long long)tile * Q.tile_traj;
    // trajectories of this tile (lanes r < nv); Q.tile_traj <= 32 is chosen by the host so that a batch
    // smaller than 32 x SMs still spreads over all SMs (fewer lanes in lock step per tile)
    const int nv = (int)((P.NQ - q0) < (long long)Q.tile_traj ? (P.NQ - q0) : (long long)Q.tile_traj);
    __syncthreads();
    // ------------------------------------------------------------------ stage the tile
    for (int e = tid; e < RT * n; e += NTHREADS) {
      const int rr = e / n, i = e - rr * n;
      const long long q = q0 + rr;
      wsX[i * RT + rr] = (rr < nv) ? P.x0[q * n + i] : 0.f;
    }
    {
      const int per = (T + 1) * n;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int rr = e / per, rest = e - rr * per;
        const long long q = q0 + rr;
        wsG[rest * RT + rr] = (rr < nv) ? P.goal[q * per + rest] : 0.f;
      }
    }
    {
      const int per = T * m;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int rr = e / per, rest = e - rr * per;
        const long long q = q0 + rr;
        wsU[rest * RT + rr] = (rr < nv) ? P.U_in[q * per + rest] : 0.f;
      }
    }
    if (tid < RT) {
      act_s[r] = (r < nv) ? 1 : 0;
      srch_s[r] = 0;
      acc_s[r] = 0;
      it_s[r] = 0;
      alpha_s[r] = Q.alpha0;
      obj_s[r] = 0.f;
    }
    bool init = true;
    int j0 = 0;   // line-search trials already done in the current outer iteration
    WeightPipe wp;
    wp.na = n; wp.nb = fout;
    const bool packed_d = Q.pack_small && nv * n <= RT;     // dynamics Jacobian rows fit the lanes
    const bool packed_c = Q.pack_small && nv * fout <= RT;  // cost-MLP Jacobian rows fit the lanes
    int NA = 0;  // number of line-search step sizes alpha_0 2^-j > alpha_min
    for (float aa = Q.alpha0; aa > Q.alpha_min && NA < 64; aa *= 0.5f) ++NA;
    if (packed_d) {  // lanes that carry no trajectory are never written by the packed linearisation
      for (int e = tid; e < T * n * n * RT; e += NTHREADS) if (r >= nv) wsA[e] = 0.f;
      for (int e = tid; e < T * n * m * RT; e += NTHREADS) if (r >= nv) wsB[e] = 0.f;
    }
    if (packed_c)
      for (int e = tid; e < fout * n * RT; e += NTHREADS) if (r >= nv) wsJf[e] = 0.f;

#pragma unroll 1
    while (true) {
      // ================================================================ rollout (plain or feedback)
      // trajax rollout (first pass) / ddp_rollout: u = U[t] + alpha k[t] + K[t] (x_new[t] - X[t]).
      // Line search: the ns lanes still searching are compacted, and each gets G = min(32 / ns, trials
      // left) lanes that roll out G consecutive step sizes alpha_0 2^-j at once; the decision takes the
      // first strict decrease in trial order -- what the sequential backtracking loop returns, trial by
      // trial (G = 1 while every lane searches).  A single state evaluates all 15 step sizes in ONE rollout.
      float* Xd = init ? wsX : wsXn;
      float* Ud = init ? wsU : wsUn;
      int ns = 0, G = 1;
      if (!init) {
        if (tid < RT) {
          const unsigned mask = __ballot_sync(0xffffffffu, srch_s[r] != 0);
          if (srch_s[r]) slist_s[__popc(mask & ((1u << r) - 1u))] = r;
          if (r == 0) slist_s[RT] = __popc(mask);
        }
        __syncthreads();
        ns = slist_s[RT];
        G = Q.pack_small ? RT / ns : 1;
        if (G > NA - j0) G = NA - j0;
      }
      const int lk = init ? r : r / G;                          // searching trajectory this lane works for
      const int lt = init ? r : (lk < ns ? slist_s[lk] : -1);    // its lane (-1: this lane idles)
      float al = Q.alpha0;                                      // this lane's step size
      if (!init)
        for (int j = j0 + (r % G); j > 0; --j) al *= 0.5f;
      wp.p = 0; wp.li = 0; wp.ci = 0; wp.issued = 0; wp.consumed = 0;
      wp.sched = SCHED_ROLL;
      wp.kind = DIR_DYN_F;
      __syncthreads();
#pragma unroll
      for (int s = 0; s < NSTAGE - 1; ++s) pipe_issue(P, wp, ring, tid);
      for (int i = tid; i < n * RT; i += NTHREADS) {
        const float v = lt < 0 ? 0.f : wsX[(i & ~31) + lt];
        q_s[i] = v;
        if (!init) Xd[i] = v;
      }
      float Jr = 0.f;
      ++n_roll;
      __syncthreads();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        for (int e = tid; e < m * RT; e += NTHREADS) {
          const int a = e >> 5;
          float u = lt < 0 ? 0.f : wsU[(t * m + a) * RT + lt];
          if (!init && lt >= 0) {
            float s = al * wsk[(t * m + a) * RT + lt];
            const float* Kr = wsK + (size_t)((t * m + a) * n) * RT + lt;
            const float* Xr = wsX + (size_t)(t * n) * RT + lt;
            for (int j = 0; j < n; ++j) s = fmaf(Kr[j * RT], q_s[j * RT + r] - Xr[j * RT], s);
            u += s;
          }
          if (!init) Ud[(t * m + a) * RT + r] = u;
          q_s[(n + a) * RT + r] = u;
        }
        __syncthreads();
        if (tid < RT && lt >= 0) {  // cost/cost_model.py:20-28 staging cost at (x_t, u_t, goal_t)
          float uu = 0.f, dd = 0.f;
          for (int j = 0; j < m; ++j) {
            const float u = q_s[(n + j) * RT + r];
            uu = fmaf(u, u, uu);
          }
          for (int i = 0; i < n; ++i) {
            const float d = q_s[i * RT + r] - wsG[(t * n + i) * RT + lt];
            dd = fmaf(d, d, dd);
          }
          Jr += w0 * (sqrtf(uu + a2) - ALPHA) + w1 * (sqrtf(dd + a2) - ALPHA);
        }
        mlp_forward<MAXT>(P, P.dir[DIR_DYN_F], q_s, q_s, EPI_RESID, Xd + (size_t)(t + 1) * n * RT, dynMask,
                          bufA, bufB, ring, wp, tid);
        __syncthreads();
      }
      mlp_forward<MAXT>(P, P.dir[DIR_COST_F], q_s, y_s, 0, nullptr, costMask, bufA, bufB, ring, wp, tid);
      cp_async_wait<0>();
      __syncthreads();
      if (tid < RT) {
        float yy = 0.f;
        for (int o = 0; o < fout; ++o) {
          const float y = y_s[o * RT + r];
          yy = fmaf(y, y, yy);
        }
        Jr += w2 * yy;  // cost/cost_model.py:30-31
        objn_s[r] = Jr;
      }
      __syncthreads();
      // ---------------------------------------------------------------- line_search_ddp decision
      if (tid < RT) {
        if (init) obj_s[r] = Jr;
        acc_s[r] = 0;
      }
      __syncthreads();
      if (!init && tid < ns) {  // thread k decides for the k-th searching trajectory
        const int tl = slist_s[tid];
        const float o = obj_s[tl];
        const float oc = (o != o) ? INFINITY : o;  // NaN objective -> inf
        float aj = Q.alpha0;
        for (int j = j0; j > 0; --j) aj *= 0.5f;
        int src = 0;
        float ar = 0.f;
        for (int jj = 0; jj < G; ++jj, aj *= 0.5f) {  // trials in backtracking order
          const float v = objn_s[tid * G + jj];
          const float on = (v != v) ? oc : v;       // NaN trial -> rejected
          ar = 0.5f * aj;
          if (on < oc) { src = tid * G + jj + 1; obj_s[tl] = on; break; }  // strict decrease only
        }
        acc_s[tl] = src;                            // source lane + 1 of the accepted trial
        alpha_s[tl] = ar;
        srch_s[tl] = (src == 0 && ar > Q.alpha_min) ? 1 : 0;
      }
      __syncthreads();
      if (!init) {
        for (int e = tid; e < (T + 1) * n * RT; e += NTHREADS)
          if (acc_s[r]) wsX[e] = wsXn[(e & ~31) + acc_s[r] - 1];
        for (int e = tid; e < T * m * RT; e += NTHREADS)
          if (acc_s[r]) wsU[e] = wsUn[(e & ~31) + acc_s[r] - 1];
        const int more = __syncthreads_or(tid < RT && srch_s[r]);
        if (more) {
          j0 += G;
          continue;
        }
      }

      // ================================================================ linearize (get_lqr_params)
      // Full tiles: lane = trajectory, one adjoint pass per Jacobian row (n per step, fout for the cost
      // MLP).  Small tiles (nv trajectories with nv n <= 32 and nv fout <= 32, e.g. the single state of
      // an acting call): lane = (trajectory, row) pair, so ONE adjoint pass yields the whole Jacobian.
      // Per lane the arithmetic is the same sequence either way (results are bitwise identical).
      const int nbd = packed_d ? 1 : n, nbc = packed_c ? 1 : fout;
      wp.p = 0; wp.li = 0; wp.ci = 0; wp.issued = 0; wp.consumed = 0;
      wp.sched = SCHED_LIN;
      wp.na = nbd; wp.nb = nbc;
      wp.kind = DIR_DYN_F;
      __syncthreads();
#pragma unroll
      for (int s = 0; s < NSTAGE - 1; ++s) pipe_issue(P, wp, ring, tid);
      const int lr_d = packed_d ? (r < nv * n ? r / n : -1) : r;      // trajectory whose data lane r carries
      const int li_d = packed_d ? (r < nv * n ? r % n : -1) : 0;      // Jacobian row lane r is seeded with
      const int lr_c = packed_c ? (r < nv * fout ? r / fout : -1) : r;
      const int li_c = packed_c ? (r < nv * fout ? r % fout : -1) : 0;
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
          const int row = e >> 5;
          q_s[e] = lr_d < 0 ? 0.f
                            : (row < n ? wsX[(t * n + row) * RT + lr_d] : wsU[(t * m + row - n) * RT + lr_d]);
        }
        __syncthreads();
        mlp_forward<MAXT>(P, P.dir[DIR_DYN_F], q_s, dq_s, 0, nullptr, dynMask, bufA, bufB, ring, wp, tid);
#pragma unroll 1
        for (int i = 0; i < nbd; ++i) {
          __syncthreads();
          const int row_seed = packed_d ? li_d : i;
          for (int e = tid; e < s4 * RT; e += NTHREADS) seed_s[e] = ((e >> 5) == row_seed) ? 1.f : 0.f;
          mlp_backward<MAXT>(P, P.dir[DIR_DYN_B], seed_s, dq_s, dynMask, bufA, bufB, ring, wp, tid);
          __syncthreads();
          if (lr_d >= 0) {
            const int ii = packed_d ? li_d : i;
            for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
              const int j = e >> 5;
              const float v = dq_s[e];
              if (j < n) wsA[(size_t)((t * n + ii) * n + j) * RT + lr_d] = v + (ii == j ? 1.f : 0.f);
              else wsB[(size_t)((t * n + ii) * m + (j - n)) * RT + lr_d] = v;
            }
          }
        }
        __syncthreads();
      }
      for (int e = tid; e < n * RT; e += NTHREADS) q_s[e] = lr_c < 0 ? 0.f : wsX[(size_t)T * n * RT + (e & ~31) + lr_c];
      __syncthreads();
      mlp_forward<MAXT>(P, P.dir[DIR_COST_F], q_s, y_s, 0, nullptr, costMask, bufA, bufB, ring, wp, tid);
#pragma unroll 1
      for (int o = 0; o < nbc; ++o) {
        __syncthreads();
        const int row_seed = packed_c ? li_c : o;
        for (int e = tid; e < s4 * RT; e += NTHREADS) seed_s[e] = ((e >> 5) == row_seed) ? 1.f : 0.f;
        mlp_backward<MAXT>(P, P.dir[DIR_COST_B], seed_s, dq_s, costMask, bufA, bufB, ring, wp, tid);
        __syncthreads();
        if (lr_c >= 0) {
          const int oo = packed_c ? li_c : o;
          for (int e = tid; e < n * RT; e += NTHREADS) wsJf[(size_t)(oo * n + (e >> 5)) * RT + lr_c] = dq_s[e];
        }
      }
      const int yl = packed_c ? r * fout : r;  // a lane that holds y of trajectory r (r < nv when packed)
      cp_async_wait<0>();
      __syncthreads();
      // terminal quadratisation: Q_T = 2 w2 Jf^T Jf, q_T = 2 w2 Jf^T y
      for (int e = tid; e < n * n * RT; e += NTHREADS) {
        const int ij = e >> 5, i = ij / n, j = ij - i * n;
        float s = 0.f;
        for (int o = 0; o < fout; ++o) s = fmaf(wsJf[(o * n + i) * RT + r], wsJf[(o * n + j) * RT + r], s);
        wsQT[e] = 2.f * w2 * s;
      }
      for (int e = tid; e < n * RT; e += NTHREADS) {
        const int i = e >> 5;
        float s = 0.f;
        for (int o = 0; o < fout; ++o) s = fmaf(wsJf[(o * n + i) * RT + r], y_s[o * RT + (yl < RT ? yl : r)], s);
        s *= 2.f * w2;
        wsqT[e] = s;
        p_s[e] = s;
        wsLam[(size_t)T * n * RT + e] = s;
      }
      // ================================================================ adjoint scan (gradient, adjoints)
      float g2 = 0.f;
      __syncthreads();
#pragma unroll 1
      for (int t = T - 1; t >= 0; --t) {
        if (tid < RT) {
          float uu = 0.f, dd = 0.f;
          for (int j = 0; j < m; ++j) {
            const float u = wsU[(t * m + j) * RT + r];
            uu = fmaf(u, u, uu);
          }
          for (int i = 0; i < n; ++i) {
            const float d = wsX[(t * n + i) * RT + r] - wsG[(t * n + i) * RT + r];
            dd = fmaf(d, d, dd);
          }
          sd_s[r] = sqrtf(dd + a2);
          su_s[r] = sqrtf(uu + a2);
        }
        __syncthreads();
        for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
          const int row = e >> 5;
          if (row < n) {
            const int i = row;
            float s = w1 * (wsX[(t * n + i) * RT + r] - wsG[(t * n + i) * RT + r]) / sd_s[r];
            const float* Ar = wsA + (size_t)(t * n * n + i) * RT + r;
            for (int k = 0; k < n; ++k) s = fmaf(Ar[(size_t)k * n * RT], p_s[k * RT + r], s);
            pn_s[i * RT + r] = s;
            wsLam[(size_t)(t * n + i) * RT + r] = s;
          } else {
            const int a = row - n;
            float s = w0 * wsU[(t * m + a) * RT + r] / su_s[r];
            const float* Br = wsB + (size_t)(t * n * m + a) * RT + r;
            for (int k = 0; k < n; ++k) s = fmaf(Br[(size_t)k * m * RT], p_s[k * RT + r], s);
            wsGrad[(t * m + a) * RT + r] = s;
            g2 = fmaf(s, s, g2);
          }
        }
        __syncthreads();
        for (int e = tid; e < n * RT; e += NTHREADS) p_s[e] = pn_s[e];
        __syncthreads();
      }
      gpart_s[w * RT + r] = g2;
      __syncthreads();
      if (tid < RT) {  // trajax ilqr continuation_criterion, per lane
        float gn2 = 0.f;
        for (int k = 0; k < NTHREADS / 32; ++k) gn2 += gpart_s[k * RT + r];
        const float gn = (gn2 != gn2) ? __int_as_float(0x7f800000) : sqrtf(gn2);   // trajax: NaN grad_norm -> inf (keeps iterating)
        if (!init && act_s[r]) it_s[r] += 1;
        // gradient_lag: the test sees the gradient of the iterate BEFORE the step just taken (the initial one twice)
        float gt = gn;
        if (Q.grad_lag) {
          if (init) wsGnP[r] = gn;
          gt = wsGnP[r];
        }
        act_s[r] = (act_s[r] && it_s[r] < Q.maxiter && gt > Q.gthr && alpha_s[r] > Q.alpha_min) ? 1 : 0;
        if (Q.grad_lag && act_s[r]) wsGnP[r] = gn;
      }
      const int any = __syncthreads_or(tid < RT && act_s[r]);
      if (Q.grad_lag) {
        // what a lane returns: the gradient / adjoints of its previous iterate -- they move on only for the lanes
        // that go on iterating (all lanes after the initial linearisation)
        for (int e = tid; e < T * m * RT; e += NTHREADS)
          if (init || act_s[e & 31]) wsGradP[e] = wsGrad[e];
        for (int e = tid; e < (T + 1) * n * RT; e += NTHREADS)
          if (init || act_s[e & 31]) wsLamP[e] = wsLam[e];
      }
      if (!any) break;
      ++n_outer;

      // ================================================================ tvlqr backward sweep
      for (int e = tid; e < n * n * RT; e += NTHREADS) P_s[e] = wsQT[e];
      for (int e = tid; e < n * RT; e += NTHREADS) p_s[e] = wsqT[e];
      __syncthreads();
#pragma unroll 1
      for (int t = T - 1; t >= 0; --t) {
        const float* Ag = wsA + (size_t)t * n * n * RT + r;  // A[k][j] at Ag[(k*n+j)*RT]
        const float* Bg = wsB + (size_t)t * n * m * RT + r;  // B[k][a] at Bg[(k*m+a)*RT]
        // 0: per-lane norms, d, u
        for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
          const int row = e >> 5;
          if (row < n) d_s[e] = wsX[(t * n + row) * RT + r] - wsG[(t * n + row) * RT + r];
          else u_s[(row - n) * RT + r] = wsU[(t * m + row - n) * RT + r];
        }
        __syncthreads();
        if (tid < RT) {
          float uu = 0.f, dd = 0.f;
          for (int j = 0; j < m; ++j) uu = fmaf(u_s[j * RT + r], u_s[j * RT + r], uu);
          for (int i = 0; i < n; ++i) dd = fmaf(d_s[i * RT + r], d_s[i * RT + r], dd);
          sd_s[r] = sqrtf(dd + a2);
          su_s[r] = sqrtf(uu + a2);
        }
        // 1: AtP = A^T P, BtP = B^T P
        for (int e = tid; e < (n * n + m * n) * RT; e += NTHREADS) {
          const int ee = e >> 5;
          float s = 0.f;
          if (ee < n * n) {
            const int i = ee / n, j = ee - i * n;
            for (int k = 0; k < n; ++k) s = fmaf(Ag[(size_t)(k * n + i) * RT], P_s[(k * n + j) * RT + r], s);
            T_s[ee * RT + r] = s;
          } else {
            const int e2 = ee - n * n, a = e2 / n, j = e2 - a * n;
            for (int k = 0; k < n; ++k) s = fmaf(Bg[(size_t)(k * m + a) * RT], P_s[(k * n + j) * RT + r], s);
            BtP_s[e2 * RT + r] = s;
          }
        }
        __syncthreads();
        // 2: AtPA -> P_s, H = BtPA, G = R + BtP B, h = r + B^T p
        for (int e = tid; e < (n * n + m * n + m * m + m) * RT; e += NTHREADS) {
          const int ee = e >> 5;
          float s = 0.f;
          if (ee < n * n) {
            const int i = ee / n, j = ee - i * n;
            for (int k = 0; k < n; ++k) s = fmaf(T_s[(i * n + k) * RT + r], Ag[(size_t)(k * n + j) * RT], s);
            P_s[ee * RT + r] = s;
          } else if (ee < n * n + m * n) {
            const int e2 = ee - n * n, a = e2 / n, j = e2 - a * n;
            for (int k = 0; k < n; ++k) s = fmaf(BtP_s[(a * n + k) * RT + r], Ag[(size_t)(k * n + j) * RT], s);
            H_s[e2 * RT + r] = s;
          } else if (ee < n * n + m * n + m * m) {
            const int e2 = ee - n * n - m * n, a = e2 / m, b = e2 - a * m;
            const float su = su_s[r];
            s = w0 * ((a == b ? 1.f : 0.f) / su - u_s[a * RT + r] * u_s[b * RT + r] / (su * su * su));
            for (int k = 0; k < n; ++k) s = fmaf(BtP_s[(a * n + k) * RT + r], Bg[(size_t)(k * m + b) * RT], s);
            G_s[e2 * RT + r] = s;
          } else {
            const int a = ee - n * n - m * n - m * m;
            s = w0 * u_s[a * RT + r] / su_s[r];
            for (int k = 0; k < n; ++k) s = fmaf(Bg[(size_t)(k * m + a) * RT], p_s[k * RT + r], s);
            h_s[a * RT + r] = s;
          }
        }
        __syncthreads();
        // 3: symmetrise AtPA and G
        for (int e = tid; e < (n * n + m * m) * RT; e += NTHREADS) {
          const int ee = e >> 5;
          if (ee < n * n) {
            const int i = ee / n, j = ee - i * n;
            if (i < j) {
              const float v = 0.5f * (P_s[(i * n + j) * RT + r] + P_s[(j * n + i) * RT + r]);
              P_s[(i * n + j) * RT + r] = v;
              P_s[(j * n + i) * RT + r] = v;
            }
          } else {
            const int e2 = ee - n * n, a = e2 / m, b = e2 - a * m;
            if (a < b) {
              const float v = 0.5f * (G_s[(a * m + b) * RT + r] + G_s[(b * m + a) * RT + r]);
              G_s[(a * m + b) * RT + r] = v;
              G_s[(b * m + a) * RT + r] = v;
            }
          }
        }
        __syncthreads();
        // 4: G_ = G + max(0, delta - lambda_min) I, Cholesky factor
        if (tid < RT) {
          if (!il_cholesky(G_s, L_s, m, r, 0.f)) {
            const float s0 = il_min_eig(G_s, m, r);
            il_cholesky(G_s, L_s, m, r, fmaxf(0.f, IL_DELTA - s0));
          }
        }
        __syncthreads();
        // 5: K = -G_^-1 H (n columns), k = -G_^-1 h
        for (int e = tid; e < (n + 1) * RT; e += NTHREADS) {
          const int c = e >> 5;
          float* x = c < n ? K_s + c * RT + r : k_s + r;           // x[a] at x[a * xs]
          const int xs = c < n ? n * RT : RT;
          const float* b = c < n ? H_s + c * RT + r : h_s + r;
          for (int a = 0; a < m; ++a) {
            float s = b[a * xs];
            for (int bb = 0; bb < a; ++bb) s = fmaf(-L_s[(a * m + bb) * RT + r], x[bb * xs], s);
            x[a * xs] = s / L_s[(a * m + a) * RT + r];
          }
          for (int a = m - 1; a >= 0; --a) {
            float s = x[a * xs];
            for (int bb = a + 1; bb < m; ++bb) s = fmaf(-L_s[(bb * m + a) * RT + r], x[bb * xs], s);
            x[a * xs] = s / L_s[(a * m + a) * RT + r];
          }
          for (int a = 0; a < m; ++a) x[a * xs] = -x[a * xs];
        }
        __syncthreads();
        // 6: H_GK = H + G K  (unshifted G, as trajax)  -> BtP_s
        for (int e = tid; e < m * n * RT; e += NTHREADS) {
          const int ee = e >> 5, a = ee / n, j = ee - a * n;
          float s = H_s[ee * RT + r];
          for (int b = 0; b < m; ++b) s = fmaf(G_s[(a * m + b) * RT + r], K_s[(b * n + j) * RT + r], s);
          BtP_s[ee * RT + r] = s;
        }
        __syncthreads();
        // 7: P = Q + AtPA + H_GK^T K + K^T H ; p = q + A^T p + H_GK^T k + K^T h ; gains out
        for (int e = tid; e < (n * n + n) * RT; e += NTHREADS) {
          const int ee = e >> 5;
          const float sd = sd_s[r];
          if (ee < n * n) {
            const int i = ee / n, j = ee - i * n;
            float s = P_s[ee * RT + r] +
                      w1 * ((i == j ? 1.f : 0.f) / sd - d_s[i * RT + r] * d_s[j * RT + r] / (sd * sd * sd));
            for (int a = 0; a < m; ++a) {
              s = fmaf(BtP_s[(a * n + i) * RT + r], K_s[(a * n + j) * RT + r], s);
              s = fmaf(K_s[(a * n + i) * RT + r], H_s[(a * n + j) * RT + r], s);
            }
            T_s[ee * RT + r] = s;
          } else {
            const int i = ee - n * n;
            float s = w1 * d_s[i * RT + r] / sd;
            for (int k = 0; k < n; ++k) s = fmaf(Ag[(size_t)(k * n + i) * RT], p_s[k * RT + r], s);
            for (int a = 0; a < m; ++a) {
              s = fmaf(BtP_s[(a * n + i) * RT + r], k_s[a * RT + r], s);
              s = fmaf(K_s[(a * n + i) * RT + r], h_s[a * RT + r], s);
            }
            pn_s[i * RT + r] = s;
          }
        }
        for (int e = tid; e < m * n * RT; e += NTHREADS) wsK[(size_t)t * m * n * RT + e] = K_s[e];
        for (int e = tid; e < m * RT; e += NTHREADS) wsk[(size_t)t * m * RT + e] = k_s[e];
        __syncthreads();
        // 8: P <- sym(.), p <- pn
        for (int e = tid; e < n * n * RT; e += NTHREADS) {
          const int ee = e >> 5, i = ee / n, j = ee - i * n;
          P_s[ee * RT + r] = 0.5f * (T_s[(i * n + j) * RT + r] + T_s[(j * n + i) * RT + r]);
        }
        for (int e = tid; e < n * RT; e += NTHREADS) p_s[e] = pn_s[e];
        __syncthreads();
      }
      // line search setup: every lane still iterating searches from alpha_0
      if (tid < RT) srch_s[r] = (act_s[r] && Q.alpha0 > Q.alpha_min) ? 1 : 0;
      j0 = 0;
      init = false;
      __syncthreads();
      if (!__syncthreads_or(tid < RT && srch_s[r])) {
        // alpha_0 <= alpha_min: no trial can run, every lane stops at the next test
        if (tid < RT && act_s[r]) { it_s[r] += 1; act_s[r] = 0; }
        break;
      }
    }


    // ================================================================== bilevel tail
    // policy/optimizers.py:59-73 at the planned U, loss = L2MPC.loss (norm/l2_policy.py:12-18).
    // The dynamics MLP is piecewise linear, so d^2 x_t / dU^2 = 0 almost everywhere and jax.hessian of
    // the objective is  A = sum_t S_t^T Q_t S_t + blockdiag(R_t),  S_t = d x_t / dU  (S_{t+1} = A_t S_t,
    // block t <- B_t), with the Jacobians A_t, B_t and Q_T = 2 w2 Jf^T Jf of the last linearisation.
    if (Q.desired != nullptr) {
      const int TM = T * m;
      float* wsD = wsXn;                       // desired states (the trial buffers are free now)
      float* wsBv = wsUn;                      // loss_grad_wrt_control, kept for the output
      float* wsS = wsb + WL.S;
      float* wsQS = wsb + WL.QS;
      float* wsH = wsb + WL.H;
      float* wsR = wsb + WL.rhs;
      float* Qm_s = P_s;                       // [n*n]
      float* cand_v = T_s;                     // [8]
      int* cand_i = reinterpret_cast<int*>(T_s + 8 * RT);  // [8]
      float* part_s = T_s + 16 * RT;           // [8]
      float* dx_s = p_s;
      float* dxn_s = pn_s;
      const float cl2 = 2.f / (float)(T + 1);
      __syncthreads();
      {
        const int per = (T + 1) * n;
        for (int e = tid; e < RT * per; e += NTHREADS) {
          const int rr = e / per, rest = e - rr * per;
          const long long q = q0 + rr;
          wsD[rest * RT + rr] = (rr < nv) ? Q.desired[q * per + rest] : 0.f;
        }
      }
      for (int e = tid; e < TM * TM * RT; e += NTHREADS) wsH[e] = 0.f;
      for (int e = tid; e < n * TM * RT; e += NTHREADS) wsS[e] = 0.f;
      __syncthreads();
      // ---- loss and B = d loss / dU (adjoint scan with q_t = 2/(T+1) (x_t - desired_t), r_t = 0)
      const bool generic = Q.bl_generic != 0;
      if (tid < RT && !generic) {
        float s = 0.f;
        for (int e = 0; e < (T + 1) * n; ++e) {
          const float d = wsX[e * RT + r] - wsD[e * RT + r];
          s = fmaf(d, d, s);
        }
        if (r < nv) Q.bl_loss[q0 + r] = s / (float)(T + 1);
      }
      if (!generic) {  // q_t = d loss / d x_t of the L2 loss, in place of the desired states
        __syncthreads();
        for (int e = tid; e < (T + 1) * n * RT; e += NTHREADS) wsD[e] = cl2 * (wsX[e] - wsD[e]);
        __syncthreads();
      }
      for (int e = tid; e < n * RT; e += NTHREADS) p_s[e] = wsD[(size_t)T * n * RT + e];
      __syncthreads();
#pragma unroll 1
      for (int t = T - 1; t >= 0; --t) {
        for (int e = tid; e < (n + m) * RT; e += NTHREADS) {
          const int row = e >> 5;
          if (row < n) {
            const int i = row;
            float s = wsD[(t * n + i) * RT + r];
            const float* Ar = wsA + (size_t)(t * n * n + i) * RT + r;
            for (int k = 0; k < n; ++k) s = fmaf(Ar[(size_t)k * n * RT], p_s[k * RT + r], s);
            pn_s[i * RT + r] = s;
          } else {
            const int a = row - n;
            float s = 0.f;
            const float* Br = wsB + (size_t)(t * n * m + a) * RT + r;
            for (int k = 0; k < n; ++k) s = fmaf(Br[(size_t)k * m * RT], p_s[k * RT + r], s);
            wsR[(t * m + a) * RT + r] = s;
            wsBv[(t * m + a) * RT + r] = s;
          }
        }
        __syncthreads();
        for (int e = tid; e < n * RT; e += NTHREADS) p_s[e] = pn_s[e];
        __syncthreads();
      }
      if (Q.bl_V != nullptr) {
        const int per = T * m;
        for (int e = tid; e < RT * per; e += NTHREADS) {
          const int rr = e / per, rest = e - rr * per;
          const long long q = q0 + rr;
          wsR[rest * RT + rr] = (rr < nv) ? Q.bl_V[q * per + rest] : 0.f;
        }
        __syncthreads();
      } else {
      // ---- Hessian: blockdiag(R_t) + sum_t S_t^T Q_t S_t
      for (int e = tid; e < T * m * m * RT; e += NTHREADS) {
        const int ee = e >> 5, t = ee / (m * m), ab = ee - t * m * m, a = ab / m, b = ab - a * m;
        float uu = 0.f;
        for (int j = 0; j < m; ++j) uu = fmaf(wsU[(t * m + j) * RT + r], wsU[(t * m + j) * RT + r], uu);
        const float su = sqrtf(uu + a2);
        wsH[(size_t)((t * m + a) * TM + t * m + b) * RT + r] =
            w0 * ((a == b ? 1.f : 0.f) / su - wsU[(t * m + a) * RT + r] * wsU[(t * m + b) * RT + r] / (su * su * su));
      }
      __syncthreads();
#pragma unroll 1
      for (int t = 0; t <= T; ++t) {
        const int ncol = t * m;
        if (ncol > 0) {
          if (t < T) {
            for (int e = tid; e < n * RT; e += NTHREADS) d_s[e] = wsX[(size_t)t * n * RT + e] - wsG[(size_t)t * n * RT + e];
            __syncthreads();
            if (tid < RT) {
              float dd = 0.f;
              for (int i = 0; i < n; ++i) dd = fmaf(d_s[i * RT + r], d_s[i * RT + r], dd);
              sd_s[r] = sqrtf(dd + a2);
            }
            __syncthreads();
            for (int e = tid; e < n * n * RT; e += NTHREADS) {
              const int ee = e >> 5, i = ee / n, j = ee - i * n;
              const float sd = sd_s[r];
              Qm_s[e] = w1 * ((i == j ? 1.f : 0.f) / sd - d_s[i * RT + r] * d_s[j * RT + r] / (sd * sd * sd));
            }
          } else {
            for (int e = tid; e < n * n * RT; e += NTHREADS) Qm_s[e] = wsQT[e];
          }
          __syncthreads();
          for (int e = tid; e < n * ncol * RT; e += NTHREADS) {  // QS = Q_t S_t
            const int ee = e >> 5, i = ee / ncol, c = ee - i * ncol;
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(Qm_s[(i * n + j) * RT + r], wsS[(size_t)(j * TM + c) * RT + r], s);
            wsQS[(size_t)(i * TM + c) * RT + r] = s;
          }
          __syncthreads();
          for (int e = tid; e < ncol * ncol * RT; e += NTHREADS) {  // lower triangle += S^T (Q S)
            const int ee = e >> 5, c1 = ee / ncol, c2 = ee - c1 * ncol;
            if (c2 > c1) continue;
            float s = 0.f;
            for (int i = 0; i < n; ++i)
              s = fmaf(wsS[(size_t)(i * TM + c1) * RT + r], wsQS[(size_t)(i * TM + c2) * RT + r], s);
            wsH[(size_t)(c1 * TM + c2) * RT + r] += s;
          }
          __syncthreads();
        }
        if (t < T) {  // S_{t+1} = A_t S_t, block t <- B_t
          const float* Ag = wsA + (size_t)t * n * n * RT + r;
          for (int e = tid; e < n * ncol * RT; e += NTHREADS) {
            const int ee = e >> 5, i = ee / ncol, c = ee - i * ncol;
            float s = 0.f;
            for (int j = 0; j < n; ++j) s = fmaf(Ag[(size_t)(i * n + j) * RT], wsS[(size_t)(j * TM + c) * RT + r], s);
            wsQS[(size_t)(i * TM + c) * RT + r] = s;
          }
          __syncthreads();
          for (int e = tid; e < n * (ncol + m) * RT; e += NTHREADS) {
            const int ee = e >> 5, i = ee / (ncol + m), c = ee - i * (ncol + m);
            wsS[(size_t)(i * TM + c) * RT + r] =
                c < ncol ? wsQS[(size_t)(i * TM + c) * RT + r] : wsB[(size_t)((t * n + i) * m + (c - ncol)) * RT + r];
          }
          __syncthreads();
        }
      }
      for (int e = tid; e < TM * TM * RT; e += NTHREADS) {  // mirror the lower triangle
        const int ee = e >> 5, c1 = ee / TM, c2 = ee - c1 * TM;
        if (c2 > c1) wsH[e] = wsH[(size_t)(c2 * TM + c1) * RT + r];
      }
      __syncthreads();
      if (Q.bl_hess != nullptr) {
        const size_t per = (size_t)TM * TM;
        for (size_t e = tid; e < RT * per; e += NTHREADS) {
          const size_t rr = e / per, rest = e - rr * per;
          const long long q = q0 + (long long)rr;
          if ((int)rr < nv) Q.bl_hess[q * per + rest] = wsH[rest * RT + rr];
        }
        __syncthreads();
      }
      // ---- H = solve(A, B): Gaussian elimination with partial pivoting, one system per lane
      // (jax.scipy.linalg.solve is LU with partial pivoting; no regularisation, policy/optimizers.py:67)
#pragma unroll 1
      for (int k = 0; k < TM; ++k) {
        float best = -1.f;
        int bi = k;
        for (int i = k + w; i < TM; i += NTHREADS / 32) {
          const float v = fabsf(wsH[(size_t)(i * TM + k) * RT + r]);
          if (v > best) { best = v; bi = i; }
        }
        cand_v[w * RT + r] = best;
        cand_i[w * RT + r] = bi;
        __syncthreads();
        int piv = k;
        best = -1.f;
        for (int ww = 0; ww < NTHREADS / 32; ++ww) {
          const float v = cand_v[ww * RT + r];
          const int ci = cand_i[ww * RT + r];
          if (v > best || (v == best && ci < piv)) { best = v; piv = ci; }
        }
        if (piv != k) {
          for (int j = k + w; j < TM; j += NTHREADS / 32) {
            const float a = wsH[(size_t)(k * TM + j) * RT + r], b = wsH[(size_t)(piv * TM + j) * RT + r];
            wsH[(size_t)(k * TM + j) * RT + r] = b;
            wsH[(size_t)(piv * TM + j) * RT + r] = a;
          }
          if (w == 0) {
            const float a = wsR[k * RT + r], b = wsR[piv * RT + r];
            wsR[k * RT + r] = b;
            wsR[piv * RT + r] = a;
          }
        }
        __syncthreads();
        const float pk = wsH[(size_t)(k * TM + k) * RT + r];
        const float rk = wsR[k * RT + r];
        for (int i = k + 1 + w; i < TM; i += NTHREADS / 32) {
          const float f = wsH[(size_t)(i * TM + k) * RT + r] / pk;
          float* row = wsH + (size_t)i * TM * RT + r;
          const float* prow = wsH + (size_t)k * TM * RT + r;
          for (int j = k + 1; j < TM; ++j) row[(size_t)j * RT] = fmaf(-f, prow[(size_t)j * RT], row[(size_t)j * RT]);
          wsR[i * RT + r] = fmaf(-f, rk, wsR[i * RT + r]);
        }
        __syncthreads();
      }
#pragma unroll 1
      for (int k = TM - 1; k >= 0; --k) {  // back substitution, the dot product split over the 8 warps
        float s = 0.f;
        for (int j = k + 1 + w; j < TM; j += NTHREADS / 32) s = fmaf(wsH[(size_t)(k * TM + j) * RT + r], wsR[j * RT + r], s);
        part_s[w * RT + r] = s;
        __syncthreads();
        if (w == 0) {
          float tot = 0.f;
          for (int ww = 0; ww < NTHREADS / 32; ++ww) tot += part_s[ww * RT + r];
          wsR[k * RT + r] = (wsR[k * RT + r] - tot) / wsH[(size_t)(k * TM + k) * RT + r];
        }
        __syncthreads();
      }
      }
      // ---- tangent rollout along H and the mpc_weights part of grad_theta (H . grad_U J)
      float g0 = 0.f, g1 = 0.f;
      for (int e = tid; e < n * RT; e += NTHREADS) dx_s[e] = 0.f;
      __syncthreads();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        if (tid < RT) {
          float uu = 0.f, dd = 0.f, uh = 0.f, ddx = 0.f;
          for (int j = 0; j < m; ++j) {
            const float u = wsU[(t * m + j) * RT + r];
            uu = fmaf(u, u, uu);
            uh = fmaf(u, wsR[(t * m + j) * RT + r], uh);
          }
          for (int i = 0; i < n; ++i) {
            const float d = wsX[(t * n + i) * RT + r] - wsG[(t * n + i) * RT + r];
            dd = fmaf(d, d, dd);
            ddx = fmaf(d, dx_s[i * RT + r], ddx);
          }
          g0 += uh / sqrtf(uu + a2);
          g1 += ddx / sqrtf(dd + a2);
        }
        for (int e = tid; e < n * RT; e += NTHREADS) {
          const int i = e >> 5;
          float s = 0.f;
          const float* Ar = wsA + (size_t)((t * n + i) * n) * RT + r;
          const float* Br = wsB + (size_t)((t * n + i) * m) * RT + r;
          for (int j = 0; j < n; ++j) s = fmaf(Ar[(size_t)j * RT], dx_s[j * RT + r], s);
          for (int a = 0; a < m; ++a) s = fmaf(Br[(size_t)a * RT], wsR[(t * m + a) * RT + r], s);
          dxn_s[e] = s;
        }
        __syncthreads();
        for (int e = tid; e < n * RT; e += NTHREADS) dx_s[e] = dxn_s[e];
        __syncthreads();
      }
      if (tid < RT && r < nv) {
        float g2 = 0.f;  // 2 y^T Jf dx_T = q_T . dx_T / w2
        for (int i = 0; i < n; ++i) {
          g2 = fmaf(wsqT[i * RT + r], dx_s[i * RT + r], g2);
          Q.bl_dxT[(q0 + r) * n + i] = dx_s[i * RT + r];
        }
        g2 /= w2;
        Q.bl_gw[(q0 + r) * 3 + 0] = g0 * w0 * (1.f - w0);
        Q.bl_gw[(q0 + r) * 3 + 1] = g1 * w1 * (1.f - w1);
        Q.bl_gw[(q0 + r) * 3 + 2] = g2 * w2 * (1.f - w2);
      }
      __syncthreads();
    }
    __syncthreads();
    // ------------------------------------------------------------------ write the tile out
    if (tid < RT && r < nv) {
      if (P.J_out != nullptr) P.J_out[q0 + r] = obj_s[r];
      if (Q.it_out != nullptr) Q.it_out[q0 + r] = it_s[r];
    }
    const bool bl = Q.desired != nullptr;
    const struct { float* dst; const float* src; int per; } outs[8] = {
        {P.U_out, wsU, T * m},          {P.X_out, wsX, (T + 1) * n}, {P.dU_out, Q.grad_lag ? wsGradP : wsGrad, T * m},
        {P.lam_out, Q.grad_lag ? wsLamP : wsLam, (T + 1) * n}, {Q.A_out, wsA, T * n * n},   {Q.B_out, wsB, T * n * m},
        {bl ? Q.bl_B : nullptr, wsUn, T * m}, {bl ? Q.bl_H : nullptr, wsb + WL.rhs, T * m}};
#pragma unroll 1
    for (int k = 0; k < 8; ++k) {
      if (outs[k].dst == nullptr) continue;
      const int per = outs[k].per;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int rr = e / per, rest = e - rr * per;
        const long long q = q0 + rr;
        if (rr < nv) outs[k].dst[q * per + rest] = outs[k].src[(size_t)rest * RT + rr];
      }
    }
  }
  if (Q.stats != nullptr && tid == 0) {
    atomicAdd(reinterpret_cast<unsigned long long*>(Q.stats), (unsigned long long)n_outer);
    atomicAdd(reinterpret_cast<unsigned long long*>(Q.stats + 1), (unsigned long long)n_roll);
  }
}

}  // namespace gmpc
