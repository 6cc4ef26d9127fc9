// plan_ffma.cuh -- fused rollout + cost + adjoint + update kernel, fp32 CUDA-core (FFMA) path.
//
// One persistent CTA per SM loops over tiles of RT=32 trajectories.  For a tile the whole planning
// loop (N iterations of forward rollout, terminal-cost MLP, adjoint sweep, in-place gradient/Adam
// update, then the final evaluation) runs inside the kernel: states, actions, adjoints and
// activations stay in shared memory / per-CTA L2-resident scratch; the MLP weights are streamed
// from L2 through a 3-stage cp.async ring (they never fit in one SM's shared memory: the default
// dynamics MLP is 346 KB).  HBM sees only the per-state input/output stream.
//
// Restates: dynamics/nn.py:27-34, cost/nn.py:23-29, cost/cost_model.py:20-42,
// policy/optimizers.py:24-31 (objective) and :78-83 (BPTT action gradient); the update rule is the
// north-star first-order planner (optax adam semantics of norm/runner.py:53).
//
// Activation layout: [feature][RT] (feature-major, 32 trajectories contiguous).  A thread owns an
// 8-trajectory x 4-feature output tile; all lanes of a warp share the trajectory group, so
// activation reads are warp broadcasts and weight reads are contiguous float4s.  Hidden buffers
// are XOR-swizzled at float4 granularity so the epilogue stores are bank-conflict free.
#pragma once
#include "common.cuh"

namespace gmpc {

struct WeightPipe {
  int p, li, ci;  // prefetch cursor: pass, layer in pass, chunk in layer
  int kind;       // DIR_* of pass p
  int issued, consumed;
  int sched;      // SCHED_*: which pass schedule the prefetch cursor follows
  int na, nb;     // SCHED_LIN: adjoint passes per dynamics step / after the cost MLP forward
};

// Pass schedules: the planner's (pass_kind), and the two phases of the iLQR kernel (ilqr.cuh):
// SCHED_LIN  = per step one forward pass + n adjoint passes (the Jacobian rows), then the cost MLP
//              forward + fout adjoint passes (n / fout become 1 when a small tile packs the
//              (trajectory, seed) pairs into the 32 lanes);  SCHED_ROLL = T forward passes + the cost MLP;
// SCHED_FIT  = P.iters forward passes then P.iters adjoint passes of the dynamics MLP (dynfit.cuh).
enum { SCHED_PLAN = 0, SCHED_LIN = 1, SCHED_ROLL = 2, SCHED_FIT = 3 };

__device__ __forceinline__ int pass_kind(const PlanParams& P, int p) {
  const int period = 2 * P.T + (P.use_cost ? 2 : 0);
  const int nb = P.iters * period;
  if (p < nb) {
    const int pp = p % period;
    if (pp < P.T) return DIR_DYN_F;
    if (P.use_cost) {
      if (pp == P.T) return DIR_COST_F;
      if (pp == P.T + 1) return DIR_COST_B;
    }
    return DIR_DYN_B;
  }
  if (!P.final_fwd) return DIR_END;
  const int pp = p - nb;
  if (pp < P.T) return DIR_DYN_F;
  if (P.use_cost && pp == P.T) return DIR_COST_F;
  return DIR_END;
}

__device__ __forceinline__ int sched_kind(const PlanParams& P, const WeightPipe& w, int p) {
  const int sched = w.sched;
  if (sched == SCHED_PLAN) return pass_kind(P, p);
  if (sched == SCHED_LIN) {
    const int per = w.na + 1, nd = P.T * per;
    if (p < nd) return (p % per == 0) ? DIR_DYN_F : DIR_DYN_B;
    if (p == nd) return DIR_COST_F;
    return (p - nd <= w.nb) ? DIR_COST_B : DIR_END;
  }
  if (sched == SCHED_FIT)  // dynfit.cuh: P.iters forward passes then P.iters adjoint passes
    return p < P.iters ? DIR_DYN_F : (p < 2 * P.iters ? DIR_DYN_B : DIR_END);
  return p < P.T ? DIR_DYN_F : (p == P.T ? DIR_COST_F : DIR_END);
}

// Issue the cp.async copies of the next weight chunk in schedule order (all threads, uniform).
__device__ __forceinline__ void pipe_issue(const PlanParams& P, WeightPipe& w, float* ring,
                                           int tid) {
  if (w.kind != DIR_END) {
    const DirDesc& D = P.dir[w.kind];
    const LayerDesc& L = D.layer[w.li];
    const int k0 = w.ci * L.kc;
    const int rows = min(L.kc, L.Ki - k0);
    const int nfl = rows * L.ld;
    const float* src = L.W + (size_t)k0 * L.ld;
    float* dst = ring + (w.issued % NSTAGE) * STAGE_FLOATS;
    for (int i = tid * 4; i < nfl; i += NTHREADS * 4) cp_async16(dst + i, src + i);
    if (++w.ci == L.nchunks) {
      w.ci = 0;
      if (++w.li == D.L) {
        w.li = 0;
        ++w.p;
        w.kind = sched_kind(P, w, w.p);
      }
    }
  }
  cp_async_commit();
  ++w.issued;
}

template <int MAXT>
struct Tiles {
  int rg[MAXT], cg[MAXT];
  bool act[MAXT];
  __device__ __forceinline__ void setup(int tid, int No) {
    const int CG = (No + 3) >> 2;
#ifndef GMPC_SPARSE_TILES
    const int CGr = CG;               // tiles packed densely over the threads (a warp may straddle two row groups)
#else
    const int CGr = (CG + 31) & ~31;  // a warp never straddles two row groups
#endif
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      const int tau = tid + j * NTHREADS;
      rg[j] = tau / CGr;
      cg[j] = tau - rg[j] * CGr;
      act[j] = (rg[j] < RT / 8) && (cg[j] < CG);
    }
  }
};

// One register-buffered group of GR reduction rows: 8 activations x 4 weights per row.
constexpr int GR = 2;
struct KGroup {
  float4 a0[GR], a1[GR], w[GR];
  __device__ __forceinline__ void load(const float* in_row, const float* w_row, int ld, int o0,
                                       int o1) {
#pragma unroll
    for (int u = 0; u < GR; ++u) {
      a0[u] = *reinterpret_cast<const float4*>(in_row + u * RT + o0);
      a1[u] = *reinterpret_cast<const float4*>(in_row + u * RT + o1);
      w[u] = *reinterpret_cast<const float4*>(w_row + u * ld);
    }
  }
  __device__ __forceinline__ void fma(float (&acc)[8][4]) const {
#pragma unroll
    for (int u = 0; u < GR; ++u) {
      const float av[8] = {a0[u].x, a0[u].y, a0[u].z, a0[u].w, a1[u].x, a1[u].y, a1[u].z, a1[u].w};
      const float wv[4] = {w[u].x, w[u].y, w[u].z, w[u].w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) acc[a][b] = fmaf(av[a], wv[b], acc[a][b]);
    }
  }
};

// One staged weight chunk for one output tile: software-pipelined (the operands of row group
// g+1 are loaded from shared memory while the FFMAs of group g issue).  rows is a multiple of 4,
// so the swizzle term is constant within an aligned pair of 2-row groups.
__device__ __forceinline__ void chunk_fma(const float* in_s, bool swz_in, const float* st, int ld,
                                          int k0, int rows, int rg, int cg, float (&acc)[8][4]) {
  const float* wcol = st + cg * 4;
  KGroup g0, g1;
  int sw = swz_in ? ((k0 >> 2) & 7) : 0;
  int o0 = ((rg * 2) ^ sw) << 2, o1 = ((rg * 2 + 1) ^ sw) << 2;
  g0.load(in_s + k0 * RT, wcol, ld, o0, o1);
#pragma unroll 1
  for (int kk = 0; kk < rows; kk += 4) {
    g1.load(in_s + (k0 + kk + 2) * RT, wcol + (kk + 2) * ld, ld, o0, o1);
    g0.fma(acc);
    if (kk + 4 < rows) {
      sw = swz_in ? (((k0 + kk + 4) >> 2) & 7) : 0;
      o0 = ((rg * 2) ^ sw) << 2;
      o1 = ((rg * 2 + 1) ^ sw) << 2;
      g0.load(in_s + (k0 + kk + 4) * RT, wcol + (kk + 4) * ld, ld, o0, o1);
    }
    g1.fma(acc);
  }
}

// acc[j][rr][cc] += sum_k W[k][cg*4+cc] * in[k][rg*8+rr], consuming the layer's chunks in order.
template <int MAXT>
__device__ __forceinline__ void gemm_acc(const PlanParams& P, const LayerDesc& L,
                                         const float* in_s, bool swz_in, float* ring,
                                         WeightPipe& wp, int tid, const Tiles<MAXT>& tl,
                                         float (&acc)[MAXT][8][4]) {
#pragma unroll
  for (int j = 0; j < MAXT; ++j)
#pragma unroll
    for (int a = 0; a < 8; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[j][a][b] = 0.f;

#pragma unroll 1
  for (int c = 0; c < L.nchunks; ++c) {
    cp_async_wait<NSTAGE - 2>();
    __syncthreads();
    pipe_issue(P, wp, ring, tid);
    const float* st = ring + (wp.consumed % NSTAGE) * STAGE_FLOATS;
    ++wp.consumed;
    const int k0 = c * L.kc;
    const int rows = min(L.kc, L.Ki - k0);  // multiple of 4 (Ki and kc are)
#pragma unroll
    for (int j = 0; j < MAXT; ++j) {
      if (tl.act[j]) chunk_fma(in_s, swz_in, st, L.ld, k0, rows, tl.rg[j], tl.cg[j], acc[j]);
    }
  }
}

enum { EPI_BIAS = 1, EPI_RELU = 2, EPI_MASK_OUT = 4, EPI_MASK_IN = 8, EPI_RESID = 16 };

// Write the accumulators out.  swz_out: hidden buffer (swizzled, all padded columns written);
// otherwise a small linear array where only columns o < No are touched.
// gout (nullable): also store rows o < No linearly to a global [No][gstride] slab (trajectory log;
// gstride > RT when the slab is a column block of a wider [No][R] matrix, dynfit.cuh).
template <int MAXT>
__device__ __forceinline__ void epilogue(const LayerDesc& L, int flags, float* out_s, bool swz_out,
                                         uint32_t* maskp, float* gout, const Tiles<MAXT>& tl,
                                         float (&acc)[MAXT][8][4], size_t gstride = RT) {
#pragma unroll
  for (int j = 0; j < MAXT; ++j) {
    if (!tl.act[j]) continue;
    uint32_t mw = 0;
    if (flags & EPI_MASK_IN) mw = maskp[j * NTHREADS];
    const int sw = swz_out ? (tl.cg[j] & 7) : 0;
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int o = tl.cg[j] * 4 + b;
      const bool in_range = o < L.No;
      if (!swz_out && !in_range) continue;
      const float bias = ((flags & EPI_BIAS) && in_range) ? L.bias[o] : 0.f;
      float v[8];
#pragma unroll
      for (int a = 0; a < 8; ++a) {
        float z = acc[j][a][b] + bias;
        if (flags & EPI_RELU) {
          if (z > 0.f) mw |= 1u << (a * 4 + b);
          z = fmaxf(z, 0.f);
        }
        if (flags & EPI_MASK_IN) z = ((mw >> (a * 4 + b)) & 1u) ? z : 0.f;
        v[a] = z;
      }
      float* p0 = out_s + o * RT + (((tl.rg[j] * 2) ^ sw) << 2);
      float* p1 = out_s + o * RT + (((tl.rg[j] * 2 + 1) ^ sw) << 2);
      if (flags & EPI_RESID) {
        const float4 r0 = *reinterpret_cast<const float4*>(p0);
        const float4 r1 = *reinterpret_cast<const float4*>(p1);
        v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
        v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
      }
      *reinterpret_cast<float4*>(p0) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(p1) = make_float4(v[4], v[5], v[6], v[7]);
      if (gout != nullptr && in_range) {
        float* g = gout + (size_t)o * gstride + tl.rg[j] * 8;
        *reinterpret_cast<float4*>(g) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(g + 4) = make_float4(v[4], v[5], v[6], v[7]);
      }
    }
    if (flags & EPI_MASK_OUT) maskp[j * NTHREADS] = mw;
  }
}

// One MLP pass.  fwd: relu hidden layers with mask capture, last layer linear (+ residual).
template <int MAXT>
__device__ __forceinline__ void mlp_forward(const PlanParams& P, const DirDesc& D,
                                            const float* in0, float* out_last, int last_flags,
                                            float* gout_last, uint32_t* mask_base, float* bufA,
                                            float* bufB, float* ring, WeightPipe& wp, int tid) {
  const float* in = in0;
  bool swz_in = false;
  float acc[MAXT][8][4];
  Tiles<MAXT> tl;
#pragma unroll 1
  for (int l = 0; l < D.L; ++l) {
    const LayerDesc& L = D.layer[l];
    tl.setup(tid, L.No);
    gemm_acc<MAXT>(P, L, in, swz_in, ring, wp, tid, tl, acc);
    if (l < D.L - 1) {
      float* out = (l & 1) ? bufB : bufA;
      epilogue<MAXT>(L, EPI_BIAS | EPI_RELU | EPI_MASK_OUT, out, true,
                     mask_base + (size_t)l * MAXT * NTHREADS + tid, nullptr, tl, acc);
      in = out;
      swz_in = true;
    } else {
      epilogue<MAXT>(L, EPI_BIAS | last_flags, out_last, false, nullptr, gout_last, tl, acc);
    }
  }
}

// Input-adjoint pass through the transposed layers (D.layer[0] is W_{L-1}^T ... D.layer[L-1] is W_0^T).
template <int MAXT>
__device__ __forceinline__ void mlp_backward(const PlanParams& P, const DirDesc& D,
                                             const float* in0, float* out_last,
                                             uint32_t* mask_base, float* bufA, float* bufB,
                                             float* ring, WeightPipe& wp, int tid) {
  const float* in = in0;
  bool swz_in = false;
  float acc[MAXT][8][4];
  Tiles<MAXT> tl;
#pragma unroll 1
  for (int lb = 0; lb < D.L; ++lb) {
    const LayerDesc& L = D.layer[lb];
    tl.setup(tid, L.No);
    gemm_acc<MAXT>(P, L, in, swz_in, ring, wp, tid, tl, acc);
    if (lb < D.L - 1) {
      float* out = (lb & 1) ? bufB : bufA;
      const int l = D.L - 2 - lb;  // hidden layer whose relu mask gates this adjoint
      epilogue<MAXT>(L, EPI_MASK_IN, out, true, mask_base + (size_t)l * MAXT * NTHREADS + tid,
                     nullptr, tl, acc);
      in = out;
      swz_in = true;
    } else {
      epilogue<MAXT>(L, 0, out_last, false, nullptr, nullptr, tl, acc);
    }
  }
}

template <int MAXT>
__global__ void __launch_bounds__(NTHREADS, 1)
plan_ffma_kernel(const __grid_constant__ PlanParams P) {
  extern __shared__ __align__(16) float smem[];
  const int tid = threadIdx.x;
  const int n = P.n, m = P.m, T = P.T;
  const int n4 = (n + 3) & ~3, nm4 = (n + m + 3) & ~3, f4 = (P.fout + 3) & ~3;
  float* bufA = smem;
  float* bufB = bufA + P.hpad * RT;
  float* ring = bufB + P.hpad * RT;
  float* q_s = ring + NSTAGE * STAGE_FLOATS;  // [nm4][RT]  rows 0..n-1 = x, n..n+m-1 = u
  float* lam_s = q_s + nm4 * RT;              // [n4][RT]
  float* dq_s = lam_s + n4 * RT;              // [nm4][RT]
  float* y_s = dq_s + nm4 * RT;               // [f4][RT]
  float* x0_s = y_s + f4 * RT;                // [n][RT]
  const int small_floats = (nm4 + n4 + nm4 + f4 + n) * RT;
  for (int i = tid; i < small_floats; i += NTHREADS) q_s[i] = 0.f;

  const bool cost_mode = (P.mode == MODE_PLAN || P.mode == MODE_OBJGRAD);
  float w0 = 0.f, w1 = 0.f, w2 = 0.f;
  if (cost_mode) {
    w0 = 1.f / (1.f + expf(-P.mpcw[0]));
    w1 = 1.f / (1.f + expf(-P.mpcw[1]));
    w2 = 1.f / (1.f + expf(-P.mpcw[2]));
  }
  const float a2 = ALPHA * ALPHA;
  const float l2scale = 2.f / (float)(T + 1);

  const int Ld = P.dir[DIR_DYN_F].L;
  const int Lc = P.use_cost ? P.dir[DIR_COST_F].L : 1;
  float* wsX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * RT;
  float* wsG = P.ws_G + (size_t)blockIdx.x * (T + 1) * n * RT;
  float* wsU = P.ws_U + (size_t)blockIdx.x * T * m * RT;
  float* wsM = P.ws_M + (size_t)blockIdx.x * T * m * RT;
  float* wsV = P.ws_V + (size_t)blockIdx.x * T * m * RT;
  const size_t mask_layer = (size_t)MAXT * NTHREADS;
  uint32_t* wsMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)T * (Ld - 1) + (Lc - 1)) * mask_layer;
  uint32_t* costMask = wsMask + (size_t)T * (Ld - 1) * mask_layer;

#pragma unroll 1
  for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
    const long long q0 = (long long)tile * RT;
    __syncthreads();
    // ------------------------------------------------------------------ stage the tile
    for (int e = tid; e < RT * n; e += NTHREADS) {
      const int r = e / n, i = e - r * n;
      const long long q = q0 + r;
      x0_s[i * RT + r] = (q < P.NQ) ? P.x0[(q / P.K) * n + i] : 0.f;
    }
    if (P.goal != nullptr) {
      const int per = (T + 1) * n;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int r = e / per, rest = e - r * per;
        const long long q = q0 + r;
        wsG[rest * RT + r] = (q < P.NQ) ? P.goal[(q / P.K) * per + rest] : 0.f;
      }
    }
    {
      const int per = T * m;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int r = e / per, rest = e - r * per;
        const long long q = q0 + r;
        wsU[rest * RT + r] = (q < P.NQ) ? P.U_in[q * per + rest] : 0.f;
        if (P.mode == MODE_PLAN && P.method == 1) {
          wsM[rest * RT + r] = 0.f;
          wsV[rest * RT + r] = 0.f;
        }
      }
    }
    WeightPipe wp;
    wp.p = 0; wp.li = 0; wp.ci = 0; wp.issued = 0; wp.consumed = 0;
    wp.sched = SCHED_PLAN;
    wp.kind = pass_kind(P, 0);
    __syncthreads();
#pragma unroll
    for (int s = 0; s < NSTAGE - 1; ++s) pipe_issue(P, wp, ring, tid);

    const long long qr = q0 + tid;  // this thread's trajectory when tid < RT
    const bool rvalid = (tid < RT) && (qr < P.NQ);
    float Jr = 0.f;

#pragma unroll 1
    for (int it = 0;; ++it) {
      const bool last = (it == P.iters);
      if (last && !P.final_fwd) break;
      // -------------------------------------------------------------- forward rollout
      for (int i = tid; i < n * RT; i += NTHREADS) {
        q_s[i] = x0_s[i];
        wsX[i] = x0_s[i];
      }
      Jr = 0.f;
      __syncthreads();
#pragma unroll 1
      for (int t = 0; t < T; ++t) {
        if (tid < RT) {
          const int r = tid;
          float uu = 0.f;
#pragma unroll 4
          for (int j = 0; j < m; ++j) {
            const float u = wsU[(t * m + j) * RT + r];
            q_s[(n + j) * RT + r] = u;
            uu = fmaf(u, u, uu);
          }
          if (cost_mode) {  // cost/cost_model.py:20-28 staging cost at (X[t], U[t], goal[t])
            float dd = 0.f;
            for (int i = 0; i < n; ++i) {
              const float d = q_s[i * RT + r] - wsG[(t * n + i) * RT + r];
              dd = fmaf(d, d, dd);
            }
            Jr += w0 * (sqrtf(uu + a2) - ALPHA) + w1 * (sqrtf(dd + a2) - ALPHA);
          } else if (P.mode == MODE_L2GRAD) {  // norm/l2_policy.py:12-18
            float dd = 0.f;
            for (int i = 0; i < n; ++i) {
              const float d = q_s[i * RT + r] - wsG[(t * n + i) * RT + r];
              dd = fmaf(d, d, dd);
            }
            Jr += dd;
          }
        }
        __syncthreads();
        mlp_forward<MAXT>(P, P.dir[DIR_DYN_F], q_s, q_s, EPI_RESID, wsX + (size_t)(t + 1) * n * RT,
                          wsMask + (size_t)t * (Ld - 1) * mask_layer, bufA, bufB, ring, wp, tid);
        __syncthreads();
      }
      // -------------------------------------------------------------- terminal cost
      if (P.use_cost) {
        mlp_forward<MAXT>(P, P.dir[DIR_COST_F], q_s, y_s, 0, nullptr, costMask, bufA, bufB, ring,
                          wp, tid);
        __syncthreads();
        if (tid < RT) {
          float yy = 0.f;
          const float s = 2.f * w2;
          for (int o = 0; o < P.fout; ++o) {
            const float y = y_s[o * RT + tid];
            yy = fmaf(y, y, yy);
            y_s[o * RT + tid] = s * y;  // d(w2*y.y)/dy, the seed of the cost-MLP adjoint
          }
          Jr += w2 * yy;  // cost/cost_model.py:30-31, cost/nn.py:29
        }
      } else if (P.mode == MODE_L2GRAD) {
        if (tid < RT) {
          float dd = 0.f;
          for (int i = 0; i < n; ++i) {
            const float d = q_s[i * RT + tid] - wsG[(T * n + i) * RT + tid];
            dd = fmaf(d, d, dd);
          }
          Jr = (Jr + dd) / (float)(T + 1);
        }
      }
      if (last) break;
      // -------------------------------------------------------------- adjoint seed lambda_T
      if (P.use_cost) {
        __syncthreads();
        mlp_backward<MAXT>(P, P.dir[DIR_COST_B], y_s, lam_s, costMask, bufA, bufB, ring, wp, tid);
      } else {
        for (int e = tid; e < n * RT; e += NTHREADS)
          lam_s[e] = l2scale * (q_s[e] - wsG[(size_t)T * n * RT + e]);
      }
      __syncthreads();
      if (P.lam_out != nullptr && rvalid) {
        for (int i = 0; i < n; ++i) P.lam_out[(qr * (T + 1) + T) * n + i] = lam_s[i * RT + tid];
      }
      // Adam bias corrections for update count k = it+1 (optax scale_by_adam)
      float bc1 = 1.f, bc2 = 1.f;
      if (P.mode == MODE_PLAN && P.method == 1) {
        bc1 = (float)(1.0 - pow((double)P.b1, (double)(it + 1)));
        bc2 = (float)(1.0 - pow((double)P.b2, (double)(it + 1)));
      }
      // -------------------------------------------------------------- adjoint sweep + update
#pragma unroll 1
      for (int t = T - 1; t >= 0; --t) {
        mlp_backward<MAXT>(P, P.dir[DIR_DYN_B], lam_s, dq_s,
                           wsMask + (size_t)t * (Ld - 1) * mask_layer, bufA, bufB, ring, wp, tid);
        __syncthreads();
        if (tid < RT) {
          const int r = tid;
          float su = 0.f, sd = 0.f;
          if (cost_mode) {
            float uu = 0.f, dd = 0.f;
            for (int j = 0; j < m; ++j) {
              const float u = wsU[(t * m + j) * RT + r];
              uu = fmaf(u, u, uu);
            }
            for (int i = 0; i < n; ++i) {
              const float d = wsX[(t * n + i) * RT + r] - wsG[(t * n + i) * RT + r];
              dd = fmaf(d, d, dd);
            }
            su = sqrtf(uu + a2);
            sd = sqrtf(dd + a2);
          }
          for (int j = 0; j < m; ++j) {
            const int ix = (t * m + j) * RT + r;
            float u = wsU[ix];
            float g = dq_s[(n + j) * RT + r];
            if (cost_mode) g = (w0 * u) / su + g;
            if (P.mode == MODE_PLAN) {
              if (P.method == 0) {
                u = u - P.lr * g;
              } else {
                const float mo = P.b1 * wsM[ix] + (1.f - P.b1) * g;
                const float ve = P.b2 * wsV[ix] + (1.f - P.b2) * g * g;
                wsM[ix] = mo;
                wsV[ix] = ve;
                u = u - P.lr * (mo / bc1) / (sqrtf(ve / bc2) + P.eps);
              }
              wsU[ix] = u;
            } else if (P.dU_out != nullptr && rvalid) {
              P.dU_out[(qr * T + t) * m + j] = g;
            }
          }
          for (int i = 0; i < n; ++i) {
            const float d = wsX[(t * n + i) * RT + r] - wsG[(t * n + i) * RT + r];
            const float c = cost_mode ? (w1 * d) / sd : l2scale * d;
            const float lam = (c + lam_s[i * RT + r]) + dq_s[i * RT + r];
            lam_s[i * RT + r] = lam;
            if (P.lam_out != nullptr && rvalid) P.lam_out[(qr * (T + 1) + t) * n + i] = lam;
          }
        }
        __syncthreads();
      }
      if (P.mode != MODE_PLAN) break;
    }
    cp_async_wait<0>();
    __syncthreads();
    // ------------------------------------------------------------------ write the tile out
    if (P.J_out != nullptr && rvalid) P.J_out[qr] = Jr;
    if (P.U_out != nullptr) {
      const int per = T * m;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int r = e / per, rest = e - r * per;
        const long long q = q0 + r;
        if (q < P.NQ) P.U_out[q * per + rest] = wsU[rest * RT + r];
      }
    }
    if (P.X_out != nullptr) {
      const int per = (T + 1) * n;
      for (int e = tid; e < RT * per; e += NTHREADS) {
        const int r = e / per, rest = e - r * per;
        const long long q = q0 + r;
        if (q < P.NQ) P.X_out[q * per + rest] = wsX[rest * RT + r];
      }
    }
  }
}

}  // namespace gmpc
