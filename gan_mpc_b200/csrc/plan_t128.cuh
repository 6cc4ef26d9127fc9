// plan_t128.cuh -- fused rollout + cost + adjoint + update kernel, 128-trajectory tiles.
//
// The contraction of a layer is issued as
//     out[128 traj x N features] = act[128 traj x K] * W[K x N]
// with the TRAJECTORIES as the M dimension of the MMA (TMEM lane = trajectory), the activations as
// the A operand IN TENSOR MEMORY and the weights streamed from L2 through a shared-memory ring as
// the B operand.  Measured on B200 (tools/t128_probe.cu, profiles/r2_t128_probe.txt): with A in TMEM
// an M=128 x N=208 x K=16 kind::f16 MMA costs N/2 = 104 cycles, the full tensor rate, and the three
// products of the fp16 hi/lo split (ah Wh + al Wh + ah Wl, one fp32 accumulator) of one k-step run
// in 312 cycles with all 148 SMs streaming their weights from L2 at the same time (10 ring slots);
// cta_group::2 brings nothing on top (336 cycles at N = 224).  plan_h16.cuh (weights as the A
// operand from shared memory, 32 trajectories as N) pays 88 cycles per block-k-step for a quarter of
// the trajectories: this kernel does 4 x the trajectories per SM in 1.5 x the tensor time.
//
// Nothing but the weights ever touches shared memory: an epilogue thread owns ONE trajectory (its
// TMEM lane) and 4 of every 16 features: tcgen05.ld -> scale / bias / ReLU / mask bit -> fp16 hi/lo
// split -> tcgen05.st straight into the A operand of the next layer.  Per-trajectory quantities
// (forward and adjoint power-of-two scales, staging costs, Adam moments) are per-thread scalars.
// The next layer's MMAs start as soon as the first 16 features of its operand are written
// (one mbarrier per k-step), so the epilogue of layer l runs under the MMAs of layer l+1; the
// accumulator is read out completely into registers first, which frees it for those MMAs.
//
// Restates the same reference lines as plan_ffma.cuh / plan_h16.cuh (dynamics/nn.py:27-34,
// cost/nn.py:23-29, cost/cost_model.py:20-42, policy/optimizers.py:24-31 and :78-83; optax adam of
// norm/runner.py:53).
#pragma once
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/gmpc.h"
#include "common.cuh"
#include "h16_common.cuh"

namespace gmpc {

constexpr int T_NB = 128;                          // trajectories per tile
constexpr int T_EPI_WARPS = 16;                    // 4 per TMEM lane quarter
constexpr int T_EPI = T_EPI_WARPS * 32;            // 512 epilogue threads
constexpr int T_THREADS = T_EPI + 64;              // + producer warp + issuer warp
constexpr int T_MAXKS = 16;                        // k-steps of the widest operand (hidden <= 256)
constexpr int T_MAX_SLOTS = 16;
constexpr uint32_t T_D_COL = 0, T_AH_COL = 256, T_AL_COL = 384;  // TMEM columns: D | A hi | A lo

struct TLayer {
  uint32_t goff;           // byte offset of this layer's tiles in the image of its pass
  int M_true;              // output features of this (possibly transposed) layer
  int npad;                // MMA N = round_up(M_true, 16)
  int nks;                 // reduction k-steps (round_up(K, 16) / 16)
  int kpg;                 // k-steps per ring group (one bulk copy, one slot)
  int bias_off;            // offset of the layer's bias in the bias table (forward layers)
  int scale_idx;           // index into the inverse weight scale table
  int pad_;
};
struct TDir {
  TLayer layer[MAXL];
  const uint8_t* gsrc;     // image of the pass: per layer, per k-step: [hi tile | lo tile]
  int L;
  int pad_;
};

struct TParams {
  TDir dir[4];
  int n, m, T, K;
  int fout, mode, method, iters, use_cost, final_fwd, ntiles, nslot;
  uint32_t slot_bytes;
  int nbias, nscale, pad_;
  long long NQ;
  float lr, b1, b2, eps;
  const float *x0, *U_in, *goal, *mpcw;
  const float* bias;       // [nbias] forward biases, padded per layer to npad
  const float* inv_scale;  // [nscale] 1 / (power-of-two weight scale) per Dense layer
  float *U_out, *X_out, *J_out, *dU_out, *lam_out;
  float *ws_X, *ws_G, *ws_U, *ws_M, *ws_V;   // per-CTA scratch, [..][128] trajectory-minor
  uint32_t* ws_mask;       // [grid][T (Ld-1) + (Lc-1)][2][512]
  uint32_t* ovf;           // incremented by a CTA that clamped an fp16 operand
};

// The pass schedule: iters x {T dyn fwd, [cost fwd, cost bwd], T dyn bwd}, then the final evaluation.
struct TPassWalk {
  int pp = 0, itc = 0;
  __device__ __forceinline__ int next(const TParams& P) {
    int kind;
    if (itc < P.iters) {
      if (pp < P.T) kind = DIR_DYN_F;
      else if (P.use_cost && pp == P.T) kind = DIR_COST_F;
      else if (P.use_cost && pp == P.T + 1) kind = DIR_COST_B;
      else kind = DIR_DYN_B;
      if (++pp == 2 * P.T + (P.use_cost ? 2 : 0)) { pp = 0; ++itc; }
    } else if (!P.final_fwd) {
      kind = DIR_END;
    } else {
      if (pp < P.T) kind = DIR_DYN_F;
      else if (P.use_cost && pp == P.T) kind = DIR_COST_F;
      else kind = DIR_END;
      ++pp;
    }
    return kind;
  }
};

struct TSmem {
  uint32_t ring, bias, inv, xs, lam, us, ssc, sig, bars, total;
};
__host__ __device__ inline TSmem t_smem_layout(int nslot, uint32_t slot_bytes, int nbias, int nscale, int n, int m) {
  TSmem s;
  s.ring = 0;
  s.bias = (uint32_t)nslot * slot_bytes;
  s.inv = s.bias + (uint32_t)((nbias + 3) & ~3) * 4;
  s.xs = s.inv + (uint32_t)((nscale + 3) & ~3) * 4;
  s.lam = s.xs + (uint32_t)n * T_NB * 4;
  s.us = s.lam + (uint32_t)n * T_NB * 4;
  s.ssc = s.us + (uint32_t)m * T_NB * 4;
  s.sig = s.ssc + T_NB * 4;
  s.bars = s.sig + 2 * T_NB * 4;
  s.total = s.bars + 8 * (2 * T_MAX_SLOTS + 2 + T_MAXKS) + 16;
  return s;
}

__host__ __device__ constexpr uint32_t t_idesc(int N) {  // kind::f16, fp16 x fp16 -> fp32, M = 128, K-major A and B
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}
__device__ __forceinline__ void t_mma(uint32_t d, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
               "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d), "r"(a_tmem), "l"(bdesc),
               "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void t_ld4(uint32_t taddr, uint32_t (&r)[4]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void t_st2(uint32_t taddr, uint32_t a, uint32_t b) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void t_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void t_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

template <int MAXKS>
__global__ void __launch_bounds__(T_THREADS, 1) plan_t128_kernel(const __grid_constant__ TParams P) {
  extern __shared__ __align__(128) uint8_t tsm[];
  const TSmem L = t_smem_layout(P.nslot, P.slot_bytes, P.nbias, P.nscale, P.n, P.m);
  float* bias_s = reinterpret_cast<float*>(tsm + L.bias);
  float* inv_s = reinterpret_cast<float*>(tsm + L.inv);
  float* xs_s = reinterpret_cast<float*>(tsm + L.xs);     // [n][128] state of the step in flight (sub 0's)
  float* lam_s = reinterpret_cast<float*>(tsm + L.lam);   // [n][128] adjoint of the step in flight
  float* us_s = reinterpret_cast<float*>(tsm + L.us);     // [m][128] actions of the operand being assembled
  float* ssc_s = reinterpret_cast<float*>(tsm + L.ssc);   // [128] forward operand scale in flight
  float* sig_s = reinterpret_cast<float*>(tsm + L.sig);   // [2][128] adjoint operand scale, by step parity
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tsm + L.bars);
  uint64_t* empty_bar = full_bar + T_MAX_SLOTS;
  uint64_t* acc_bar = empty_bar + T_MAX_SLOTS;   // all MMAs of a layer are complete
  uint64_t* dr_bar = acc_bar + 1;                // every epilogue warp has read the accumulator out
  uint64_t* act_bar = dr_bar + 1;                // [MAXKS] k-step j of the next operand is in TMEM
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_bar + T_MAXKS);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = P.n, m = P.m, T = P.T, NS = P.nslot;
  const int my_tiles = ((int)blockIdx.x < P.ntiles) ? (P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  for (int i = tid; i < P.nbias; i += T_THREADS) bias_s[i] = P.bias[i];
  for (int i = tid; i < P.nscale; i += T_THREADS) inv_s[i] = P.inv_scale[i];
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init(dr_bar, T_EPI_WARPS);
    for (int j = 0; j < T_MAXKS; ++j) mbar_init(&act_bar[j], T_EPI_WARPS);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_holder, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================== weight-stream producer
    // One ring slot = one group = kpg k-steps of one layer ([hi tile | lo tile] each), one bulk copy.
    // tools/bulk_copy_rate.cu: a cp.async.bulk holds its issuing lane ~460 cycles, so four lanes take
    // every fourth group; all lanes walk the same schedule.
    if (lane < 4) {
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), ring_a = smem_u32(tsm + L.ring);
      uint32_t slot = 0, ph = 0, gc = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        TPassWalk walk;
        for (;;) {
          const int kind = walk.next(P);
          if (kind == DIR_END) break;
          const TDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TLayer& Y = D.layer[l];
            const uint32_t kb = (uint32_t)Y.npad * 64u;
            uint32_t off = Y.goff;
            for (int j = 0; j < Y.nks; j += Y.kpg) {
              const uint32_t bytes = (uint32_t)min(Y.kpg, Y.nks - j) * kb;
              if ((gc & 3u) == (uint32_t)lane) {
                mbar_wait_a(empty_a + slot * 8, ph ^ 1);
                const uint32_t bar = full_a + slot * 8;
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
                asm volatile(
                    "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                    ::"r"(ring_a + slot * P.slot_bytes), "l"(D.gsrc + off), "r"(bytes), "r"(bar) : "memory");
              }
              off += bytes;
              ++gc;
              if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ================================================================== MMA issuer (one thread)
    if (elect_one()) {
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), ring_a = smem_u32(tsm + L.ring);
      const uint32_t acc_a = smem_u32(acc_bar), dr_a = smem_u32(dr_bar), act_a = smem_u32(act_bar);
      const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 0, 128) >> 32);
      const uint32_t d_t = tmem_base + T_D_COL, ah_t = tmem_base + T_AH_COL, al_t = tmem_base + T_AL_COL;
      uint32_t slot = 0, ph = 0, act_par = 0, dr_par = 0;
      for (int ti = 0; ti < my_tiles; ++ti) {
        TPassWalk walk;
        for (;;) {
          const int kind = walk.next(P);
          if (kind == DIR_END) break;
          const TDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TLayer& Y = D.layer[l];
            const uint32_t idesc = t_idesc(Y.npad);
            const uint32_t tile16 = ((uint32_t)Y.npad * 32u) >> 4;           // one (hi or lo) tile, in 16-byte units
            const uint32_t lbo_f = (((uint32_t)Y.npad * 16u) >> 4) << 16;    // LBO field of the descriptor
            int j = 0;
            while (j < Y.nks) {
              const int nk = min(Y.kpg, Y.nks - j);
              uint32_t b_lo = ((ring_a + slot * P.slot_bytes) >> 4) | lbo_f;
              for (int jj = 0; jj < nk; ++jj, ++j, b_lo += 2 * tile16) {
                mbar_wait_a(act_a + 8 * j, (act_par >> j) & 1u);
                act_par ^= 1u << j;
                if (j == 0) { mbar_wait_a(dr_a, dr_par); dr_par ^= 1; }
                if (jj == 0) mbar_wait_a(full_a + slot * 8, ph);
                tc_fence_after();
                const uint64_t bh = ((uint64_t)desc_hi << 32) | b_lo;
                const uint64_t bl = ((uint64_t)desc_hi << 32) | (b_lo + tile16);
                t_mma(d_t, ah_t + 8 * j, bh, idesc, j > 0 ? 1u : 0u);
                t_mma(d_t, al_t + 8 * j, bh, idesc, 1u);
                t_mma(d_t, ah_t + 8 * j, bl, idesc, 1u);
              }
              umma_commit_a(empty_a + slot * 8);
              if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
            }
            umma_commit_a(acc_a);
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue / per-trajectory work
    const int ew = warp - 2;               // 0..15
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int sub = ew >> 2;               // which 4 of every 16 features; sub 0 also owns the trajectory's state,
                                           // sub 1 its action update
    const int r = q * 32 + lane;           // trajectory (TMEM lane) in the tile
    const int et = ew * 32 + lane;         // 0..511
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc_a = smem_u32(acc_bar), dr_a = smem_u32(dr_bar), act_a = smem_u32(act_bar);
    uint32_t acc_ph = 0;
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
    float* wsX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * T_NB + r;
    float* wsG = P.ws_G + (size_t)blockIdx.x * (T + 1) * n * T_NB + r;
    float* wsU = P.ws_U + (size_t)blockIdx.x * T * m * T_NB + r;
    float* wsM = P.ws_M + (size_t)blockIdx.x * T * m * T_NB + r;
    float* wsV = P.ws_V + (size_t)blockIdx.x * T * m * T_NB + r;
    float* xs = xs_s + r;     // this trajectory's column of the shared per-trajectory arrays
    float* lam = lam_s + r;
    float* us = us_s + r;
    constexpr int MSTR = 2 * T_EPI;  // mask words per (step, layer)
    uint32_t* wsMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)T * (Ld - 1) + (Lc - 1)) * MSTR + et;
    uint32_t* costMask = wsMask + (size_t)T * (Ld - 1) * MSTR;
    const bool adam = (P.mode == MODE_PLAN && P.method == 1);
    const bool need_goal = cost_mode || P.mode == MODE_L2GRAD;
    const int nk_dynf = P.dir[DIR_DYN_F].layer[0].nks, nk_dynb = P.dir[DIR_DYN_B].layer[0].nks;
    const int nk_costf = P.use_cost ? P.dir[DIR_COST_F].layer[0].nks : 0;
    const int nk_costb = P.use_cost ? P.dir[DIR_COST_B].layer[0].nks : 0;
    float opmax = 0.f;  // largest operand magnitude this thread has written (fp16 range check)

    auto wait_acc = [&]() {
      mbar_wait_a(acc_a, acc_ph);
      acc_ph ^= 1;
      tc_fence_after();
    };
    // "my reads of the accumulator are complete" (after tcgen05.wait::ld)
    auto arrive_drained = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(dr_a);
    };
    // "k-steps [0, nk) of the next operand: my part is in TMEM" (after tcgen05.wait::st)
    auto arrive_act_range = [&](int nk) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0)
        for (int j = 0; j < nk; ++j) mbar_arrive_a(act_a + 8 * j);
    };
    // sub 0: features [16 ks, 16 ks + 16) of a small operand, already scaled -> k-step ks of the A operand
    auto store_kstep = [&](int ks, const float (&v)[16]) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        opmax = fmaxf(opmax, fmaxf(fabsf(v[2 * k]), fabsf(v[2 * k + 1])));
        split_h2(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
      }
      t_st8(tl + T_AH_COL + 8 * ks, hi);
      t_st8(tl + T_AL_COL + 8 * ks, lo);
    };
    auto load16 = [&](int half, float (&o)[16]) {
      uint32_t d0[16];
      tmem_ld16_issue(tl + T_D_COL + 16 * half, d0);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = __uint_as_float(d0[c]);
    };

    for (int ti = 0; ti < my_tiles; ++ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      const long long q0 = (long long)tile * T_NB;
      named_bar_sync(1, T_EPI);
      // ---------------------------------------------------------------- stage the tile (trajectory-minor scratch)
      {
        float* bX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * T_NB;
        float* bG = P.ws_G + (size_t)blockIdx.x * (T + 1) * n * T_NB;
        float* bU = P.ws_U + (size_t)blockIdx.x * T * m * T_NB;
        float* bM = P.ws_M + (size_t)blockIdx.x * T * m * T_NB;
        float* bV = P.ws_V + (size_t)blockIdx.x * T * m * T_NB;
        for (int e = et; e < T_NB * n; e += T_EPI) {
          const int rr = e / n, i = e - rr * n;
          const long long qq = q0 + rr;
          bX[i * T_NB + rr] = (qq < P.NQ) ? P.x0[(qq / P.K) * n + i] : 0.f;
        }
        if (P.goal != nullptr) {
          const int per = (T + 1) * n;
          for (int e = et; e < T_NB * per; e += T_EPI) {
            const int rr = e / per, rest = e - rr * per;
            const long long qq = q0 + rr;
            bG[rest * T_NB + rr] = (qq < P.NQ) ? P.goal[(qq / P.K) * per + rest] : 0.f;
          }
        }
        const int per = T * m;
        for (int e = et; e < T_NB * per; e += T_EPI) {
          const int rr = e / per, rest = e - rr * per;
          const long long qq = q0 + rr;
          bU[rest * T_NB + rr] = (qq < P.NQ) ? P.U_in[qq * per + rest] : 0.f;
          if (adam) {
            bM[rest * T_NB + rr] = 0.f;
            bV[rest * T_NB + rr] = 0.f;
          }
        }
      }
      __threadfence_block();
      named_bar_sync(1, T_EPI);
      const long long qr = q0 + r;
      const bool rvalid = qr < P.NQ;
      float Jr = 0.f;
      float sg = 1.f;           // sub 0: scale of the adjoint operand in flight
      float bc1 = 1.f, bc2 = 1.f;
      // staging cost of step t for this thread's trajectory (sub 0): x from shared memory, u / goal from scratch
      auto stage_cost = [&](int t) {
        float dd = 0.f, uu = 0.f;
        if (need_goal)
          for (int i = 0; i < n; ++i) {
            const float d = xs[i * T_NB] - wsG[((size_t)t * n + i) * T_NB];
            dd = fmaf(d, d, dd);
          }
        if (cost_mode) {
          for (int j = 0; j < m; ++j) {
            const float u = wsU[((size_t)t * m + j) * T_NB];
            uu = fmaf(u, u, uu);
          }
          Jr += w0 * (sqrtf(uu + a2) - ALPHA) + w1 * (sqrtf(dd + a2) - ALPHA);
        } else {
          Jr += dd;
        }
      };
      // sub 0: the <= 32-feature operand [xs ; us] * sc (us: nu features after the n state features) -> k-steps [0, nk)
      auto publish_state_operand = [&](const float* xcol, int nu, float sc, int nk) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          if (h < nk) {
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              const int i = 16 * h + c;
              float x = 0.f;
              if (i < n) x = xcol[i * T_NB];
              else if (i < n + nu) x = us[(i - n) * T_NB];
              v[c] = x * sc;
            }
            store_kstep(h, v);
          }
        }
        t_st_wait();
      };
      // every layer's first MMA waits for "accumulator drained" by the epilogue of the layer before it (also
      // across tiles); only the very first layer of the kernel has no predecessor
      if (ti == 0) arrive_drained();

      TPassWalk walk;
      int it = 0, tf = 0, tb = T - 1;
      for (;;) {
        const int kind = walk.next(P);
        if (kind == DIR_END) break;
        const TDir& D = P.dir[kind];
        const bool last = (it == P.iters);
        const bool fwd = (kind == DIR_DYN_F || kind == DIR_COST_F);
        if (kind == DIR_DYN_F && tf == 0) {
          // ------------------------------------------------------------ start of a forward sweep
          named_bar_sync(1, T_EPI);  // the previous sweep's updates of U are visible
          if (sub == 0) {
            float mx = 1.f;
            for (int i = 0; i < n; ++i) {
              const float x = wsX[i * T_NB];
              xs[i * T_NB] = x;
              mx = fmaxf(mx, fabsf(x));
            }
            for (int j = 0; j < m; ++j) us[j * T_NB] = wsU[j * T_NB];
            const float s0 = pow2_scale_to_8(mx);
            ssc_s[r] = s0;
            publish_state_operand(xs, m, s0, nk_dynf);
          }
          arrive_act_range(nk_dynf);
          Jr = 0.f;
          if (sub == 0) stage_cost(0);
          if (adam && sub == 1) {
            bc1 = (float)(1.0 - pow((double)P.b1, (double)(it + 1)));
            bc2 = (float)(1.0 - pow((double)P.b2, (double)(it + 1)));
          }
        }
        // -------------------------------------------------------------- hidden layers of the pass
        for (int l = 0; l < D.L - 1; ++l) {
          const TLayer& Y = D.layer[l];
          uint32_t* maskp;
          if (kind == DIR_DYN_F) maskp = wsMask + ((size_t)tf * (Ld - 1) + l) * MSTR;
          else if (kind == DIR_COST_F) maskp = costMask + (size_t)l * MSTR;
          else if (kind == DIR_COST_B) maskp = costMask + (size_t)(D.L - 2 - l) * MSTR;
          else maskp = wsMask + ((size_t)tb * (Ld - 1) + (D.L - 2 - l)) * MSTR;
          // accumulator -> (+bias, ReLU, mask bit) or (mask gate) -> hi/lo -> next A operand.  The forward pass
          // runs in scaled units a' = a s (ReLU is positively homogeneous): the bias enters as b s.
          const int nko = Y.npad >> 4;
          uint32_t m0 = 0, m1 = 0;
          if (!fwd) { m0 = maskp[0]; m1 = maskp[T_EPI]; }
          const float inv = inv_s[Y.scale_idx];
          const float* bp = bias_s + Y.bias_off + 4 * sub;
          wait_acc();
          const float s = fwd ? ssc_s[r] : 1.f;
          uint32_t d[MAXKS][4];
#pragma unroll
          for (int j = 0; j < MAXKS; ++j)
            if (j < nko) t_ld4(tl + T_D_COL + 16 * j + 4 * sub, d[j]);
          tmem_ld_wait();
          arrive_drained();
#pragma unroll
          for (int j = 0; j < MAXKS; ++j) {
            if (j < nko) {
              float z[4];
              if (fwd) {
                const float4 b4 = *reinterpret_cast<const float4*>(bp + 16 * j);
                z[0] = fmaf(__uint_as_float(d[j][0]), inv, b4.x * s);
                z[1] = fmaf(__uint_as_float(d[j][1]), inv, b4.y * s);
                z[2] = fmaf(__uint_as_float(d[j][2]), inv, b4.z * s);
                z[3] = fmaf(__uint_as_float(d[j][3]), inv, b4.w * s);
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  // bit = (z > 0): 0 - z is negative exactly then (and +0 for z == 0)
                  if (j < 8) m0 = __funnelshift_l(__float_as_uint(0.f - z[c]), m0, 1);
                  else       m1 = __funnelshift_l(__float_as_uint(0.f - z[c]), m1, 1);
                  z[c] = fmaxf(z[c], 0.f);
                }
              } else {
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                  const float v = __uint_as_float(d[j][c]) * inv;
                  if (j < 8) { z[c] = ((int)m0 < 0) ? v : 0.f; m0 <<= 1; }
                  else       { z[c] = ((int)m1 < 0) ? v : 0.f; m1 <<= 1; }
                }
              }
              opmax = fmaxf(opmax, fmaxf(fmaxf(fabsf(z[0]), fabsf(z[1])), fmaxf(fabsf(z[2]), fabsf(z[3]))));
              uint32_t h0, l0, h1, l1;
              split_h2(z[0], z[1], h0, l0);
              split_h2(z[2], z[3], h1, l1);
              t_st2(tl + T_AH_COL + 8 * j + 2 * sub, h0, h1);
              t_st2(tl + T_AL_COL + 8 * j + 2 * sub, l0, l1);
              t_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_a(act_a + 8 * j);
            }
          }
          if (fwd) {  // first element's bit at bit 31 (the adjoint sweep shifts them out in the same order)
            const int nb0 = 4 * min(nko, 8), nb1 = 4 * max(nko - 8, 0);
            maskp[0] = nb0 < 32 ? m0 << (32 - nb0) : m0;
            maskp[T_EPI] = nb1 == 0 ? 0u : (nb1 < 32 ? m1 << (32 - nb1) : m1);
          }
        }
        // -------------------------------------------------------------- last layer of the pass (<= 32 outputs)
        const TLayer& Yf = D.layer[D.L - 1];
        const float invf = inv_s[Yf.scale_idx];
        if (kind == DIR_DYN_F) {
          // step boundary: x_{t+1} = x_t + Dense(h); sub 0 writes the next operand [x_{t+1} ; u_{t+1}] s_{t+1}
          const int t = tf;
          const bool with_u = (t + 1 < T);
          const bool more = with_u || P.use_cost;
          const int nk_next = with_u ? nk_dynf : nk_costf;
          if (sub == 0 && with_u)
            for (int j = 0; j < m; ++j) us[j * T_NB] = wsU[((size_t)(t + 1) * m + j) * T_NB];
          wait_acc();
          if (sub == 0) {
            const float rs = pow2_recip(ssc_s[r]);
            const float* bl = bias_s + Yf.bias_off;
            float mx = 1.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (16 * h < n) {
                float o[16];
                load16(h, o);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int i = 16 * h + c;
                  if (i < n) {
                    const float xn = fmaf(o[c] * invf, rs, bl[i]) + xs[i * T_NB];
                    xs[i * T_NB] = xn;
                    mx = fmaxf(mx, fabsf(xn));
                  }
                }
              }
            }
            arrive_drained();
            if (more) {
              const float sn = pow2_scale_to_8(mx);
              ssc_s[r] = sn;
              publish_state_operand(xs, with_u ? m : 0, sn, nk_next);
            }
          } else {
            arrive_drained();
          }
          if (more) arrive_act_range(nk_next);
          if (sub == 0) {  // off the critical path: keep x_{t+1} for the adjoint sweep, its staging cost
            for (int i = 0; i < n; ++i) wsX[((size_t)(t + 1) * n + i) * T_NB] = xs[i * T_NB];
            if (t + 1 < T) stage_cost(t + 1);
          }
          if (++tf == T) {
            tf = 0;
            if (!P.use_cost) {
              if (P.mode == MODE_L2GRAD && sub == 0) {
                float dd = 0.f;
                for (int i = 0; i < n; ++i) {
                  const float d = xs[i * T_NB] - wsG[((size_t)T * n + i) * T_NB];
                  dd = fmaf(d, d, dd);
                }
                Jr = (Jr + dd) / (float)(T + 1);
              }
              if (!last) {  // adjoint seed lambda_T of the L2 loss
                if (sub == 0) {
                  float mx = 0.f;
                  for (int i = 0; i < n; ++i) {
                    const float lv = l2scale * (xs[i * T_NB] - wsG[((size_t)T * n + i) * T_NB]);
                    lam[i * T_NB] = lv;
                    mx = fmaxf(mx, fabsf(lv));
                    if (P.lam_out != nullptr && rvalid) P.lam_out[(qr * (T + 1) + T) * n + i] = lv;
                  }
                  sg = pow2_scale_to_8(mx);
                  sig_s[((T - 1) & 1) * T_NB + r] = sg;
                  publish_state_operand(lam, 0, sg, nk_dynb);
                }
                arrive_act_range(nk_dynb);
              }
            }
          }
        } else if (kind == DIR_COST_F) {
          // ------------------------------------------------------------ terminal cost w2 |y|^2 and the adjoint seed
          wait_acc();
          if (sub == 0) {
            const float rs = pow2_recip(ssc_s[r]);
            const float* bl = bias_s + Yf.bias_off;
            const float s2 = 2.f * w2;
            float yy = 0.f, mx = 0.f;
            float y0[16], y1[16];
            load16(0, y0);
            if (P.fout > 16) load16(1, y1);
            arrive_drained();
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float ya = 0.f, yb = 0.f;
              if (c < P.fout) ya = fmaf(y0[c] * invf, rs, bl[c]);
              if (P.fout > 16 && 16 + c < P.fout) yb = fmaf(y1[c] * invf, rs, bl[16 + c]);
              yy = fmaf(ya, ya, fmaf(yb, yb, yy));
              mx = fmaxf(mx, fmaxf(fabsf(s2 * ya), fabsf(s2 * yb)));
              y0[c] = s2 * ya;
              y1[c] = s2 * yb;
            }
            Jr += w2 * yy;
            if (!last) {  // dJ/dy = 2 w2 y, scaled per trajectory into fp16 range
              sg = pow2_scale_to_8(mx);
#pragma unroll
              for (int c = 0; c < 16; ++c) { y0[c] *= sg; y1[c] *= sg; }
              store_kstep(0, y0);
              if (nk_costb > 1) store_kstep(1, y1);
              t_st_wait();
            }
          } else {
            arrive_drained();
          }
          if (!last) arrive_act_range(nk_costb);
        } else if (kind == DIR_COST_B) {
          // ------------------------------------------------------------ lambda_T = d(terminal cost)/dx_T
          wait_acc();
          if (sub == 0) {
            const float c0 = invf * pow2_recip(sg);
            float mx = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (16 * h < n) {
                float o[16];
                load16(h, o);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int i = 16 * h + c;
                  if (i < n) {
                    const float lv = o[c] * c0;
                    lam[i * T_NB] = lv;
                    mx = fmaxf(mx, fabsf(lv));
                  }
                }
              }
            }
            arrive_drained();
            sg = pow2_scale_to_8(mx);
            sig_s[((T - 1) & 1) * T_NB + r] = sg;
            publish_state_operand(lam, 0, sg, nk_dynb);
          } else {
            arrive_drained();
          }
          arrive_act_range(nk_dynb);
          if (sub == 0 && P.lam_out != nullptr && rvalid)
            for (int i = 0; i < n; ++i) P.lam_out[(qr * (T + 1) + T) * n + i] = lam[i * T_NB];
        } else {
          // ------------------------------------------------------------ adjoint step boundary
          // lambda_t = l_x(x_t) + lambda_{t+1} + dq_x (sub 0);  g_u = l_u(u_t) + dq_u -> update of u_t (sub 1)
          const int t = tb;
          const int nk_next = t > 0 ? nk_dynb : 0;
          float su = 1.f;
          if (sub == 0) {  // l_x + lambda_{t+1} while the last layer's MMAs run
            float dd = 0.f;
            if (need_goal) {
              for (int i = 0; i < n; ++i) {
                const float d = wsX[((size_t)t * n + i) * T_NB] - wsG[((size_t)t * n + i) * T_NB];
                dd = fmaf(d, d, dd);
              }
              const float f = cost_mode ? w1 / sqrtf(dd + a2) : l2scale;
              for (int i = 0; i < n; ++i) {
                const float d = wsX[((size_t)t * n + i) * T_NB] - wsG[((size_t)t * n + i) * T_NB];
                lam[i * T_NB] = fmaf(d, f, lam[i * T_NB]);
              }
            }
          } else if (sub == 1 && cost_mode) {
            float uu = 0.f;
            for (int j = 0; j < m; ++j) {
              const float u = wsU[((size_t)t * m + j) * T_NB];
              uu = fmaf(u, u, uu);
            }
            su = sqrtf(uu + a2);
          }
          wait_acc();
          if (sub == 0) {
            const float c0 = invf * pow2_recip(sg);
            float mx = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (16 * h < n) {
                float o[16];
                load16(h, o);
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int i = 16 * h + c;
                  if (i < n) {
                    const float lv = fmaf(o[c], c0, lam[i * T_NB]);
                    lam[i * T_NB] = lv;
                    mx = fmaxf(mx, fabsf(lv));
                  }
                }
              }
            }
            arrive_drained();
            if (t > 0) {
              sg = pow2_scale_to_8(mx);
              sig_s[((t - 1) & 1) * T_NB + r] = sg;
              publish_state_operand(lam, 0, sg, nk_next);
              arrive_act_range(nk_next);
            }
            if (P.lam_out != nullptr && rvalid)
              for (int i = 0; i < n; ++i) P.lam_out[(qr * (T + 1) + t) * n + i] = lam[i * T_NB];
          } else if (sub == 1) {
            const float c0 = invf * pow2_recip(sig_s[(t & 1) * T_NB + r]);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (16 * h < n + m && 16 * h + 16 > n) {
                float o[16];
                load16(h, o);
                if (h == 1 || n + m <= 16) {  // the accumulator is read out: free it, the update is off the critical path
                  arrive_drained();
                  if (t > 0) arrive_act_range(nk_next);
                }
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int i = 16 * h + c;
                  if (i >= n && i < n + m) {
                    const int j = i - n;
                    const size_t ix = ((size_t)t * m + j) * T_NB;
                    float u = wsU[ix];
                    float g = o[c] * c0;
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
                }
              } else if (h == 1 && n + m > 16) {   // (unreachable for n + m > 16: the second half always holds u features)
                arrive_drained();
                if (t > 0) arrive_act_range(nk_next);
              }
            }
          } else {
            arrive_drained();
            if (t > 0) arrive_act_range(nk_next);
          }
          if (--tb < 0) {
            tb = T - 1;
            ++it;
            __threadfence_block();  // this sweep's updates of U are read by sub 0 in the next sweep
          }
        }
      }
      __threadfence_block();
      named_bar_sync(1, T_EPI);
      // ---------------------------------------------------------------- write the tile out
      if (P.J_out != nullptr && sub == 0 && rvalid) P.J_out[qr] = Jr;
      {
        const float* bX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * T_NB;
        const float* bU = P.ws_U + (size_t)blockIdx.x * T * m * T_NB;
        if (P.U_out != nullptr) {
          const int per = T * m;
          for (int e = et; e < T_NB * per; e += T_EPI) {
            const int rr = e / per, rest = e - rr * per;
            const long long qq = q0 + rr;
            if (qq < P.NQ) P.U_out[qq * per + rest] = bU[rest * T_NB + rr];
          }
        }
        if (P.X_out != nullptr) {
          const int per = (T + 1) * n;
          for (int e = et; e < T_NB * per; e += T_EPI) {
            const int rr = e / per, rest = e - rr * per;
            const long long qq = q0 + rr;
            if (qq < P.NQ) P.X_out[qq * per + rest] = bX[rest * T_NB + rr];
          }
        }
      }
    }
    if (!(opmax <= 65000.f) && P.ovf != nullptr) atomicAdd(P.ovf, 1u);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// Pack one Dense kernel W[K][N] (flax layout) into the B-operand tile stream of one direction, scaled
// by the power of two that puts max |W| in [2^10, 2^11) (same rule as h16_pack_kernel).
//   transposed == 0 (forward):  operand row = output feature n, reduction index = input feature k.
//   transposed == 1 (adjoint):  operand row = input feature k,  reduction index = output feature n.
// Image: per k-step [hi tile | lo tile], tile = [2 k-chunks][npad rows][8 halfs]
// (K-major SWIZZLE_NONE core matrices: LBO = npad * 16, SBO = 128; pinned by tools/t128_probe.cu).
__global__ void t128_pack_kernel(const float* __restrict__ W, int K, int N, int transposed, uint8_t* dst, int npad,
                                 const uint32_t* absmax, float* inv_scale) {
  const float mx = __uint_as_float(*absmax);
  float sc = 1.f;
  if (mx > 0.f) {
    const int e = (int)((__float_as_uint(mx) >> 23) & 0xFF) - 127;
    int k = 10 - e;
    k = k > 60 ? 60 : (k < -60 ? -60 : k);
    sc = __uint_as_float((uint32_t)(k + 127) << 23);
  }
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0 && inv_scale != nullptr) *inv_scale = 1.f / sc;
  if (idx >= K * N) return;
  const int k = idx / N, o = idx - k * N;
  const int row = transposed ? k : o, kk = transposed ? o : k;
  __half hi, lo;
  split_h1(W[idx] * sc, hi, lo);
  const size_t tile_b = (size_t)npad * 32;
  const int j = kk >> 4, k16 = kk & 15;
  uint8_t* p = dst + (size_t)j * 2 * tile_b + (size_t)(k16 >> 3) * npad * 16 + (row >> 3) * 128 + (row & 7) * 16 + (k16 & 7) * 2;
  *reinterpret_cast<__half*>(p) = hi;
  *reinterpret_cast<__half*>(p + tile_b) = lo;
}

// ------------------------------------------------------------------------------------- host side
struct T128State {
  bool supported = false;
  std::string why = "not initialised";
  int dyn_dims[MAXL + 1], cost_dims[MAXL + 1], Ld = 0, Lc = 0;
  int n = 0, m = 0, T = 0, fout = 0, num_sms = 0;
  TDir dir[4];
  uint8_t* d_stream = nullptr;
  size_t stream_bytes = 0;
  float* d_bias = nullptr;
  float* d_scale = nullptr;
  uint32_t* d_absmax = nullptr;
  uint32_t* d_ovf = nullptr;
  float *ws_X = nullptr, *ws_G = nullptr, *ws_U = nullptr, *ws_M = nullptr, *ws_V = nullptr;
  uint32_t* ws_mask = nullptr;
  int nbias = 0, nscale = 0, nslot = 0, maxks = 13;
  uint32_t slot_bytes = 0;
  size_t smem_bytes = 0;
};

using T128Kernel = void (*)(const TParams);
inline T128Kernel t128_kernel_ptr(int maxks) { return maxks <= 13 ? plan_t128_kernel<13> : plan_t128_kernel<16>; }

inline int t_rup(int v, int a) { return (v + a - 1) / a * a; }

inline void t128_destroy(T128State& S) {
  cudaFree(S.d_stream); cudaFree(S.d_bias); cudaFree(S.d_scale); cudaFree(S.d_absmax); cudaFree(S.d_ovf);
  cudaFree(S.ws_X); cudaFree(S.ws_G); cudaFree(S.ws_U); cudaFree(S.ws_M); cudaFree(S.ws_V); cudaFree(S.ws_mask);
  S.d_stream = nullptr; S.d_bias = S.d_scale = nullptr; S.d_absmax = S.d_ovf = nullptr;
  S.ws_X = S.ws_G = S.ws_U = S.ws_M = S.ws_V = nullptr; S.ws_mask = nullptr;
  S.supported = false;
}

inline int t128_create(T128State& S, const gmpc_config& c, const int* dyn_dims, const int* cost_dims,
                       int num_sms, size_t smem_optin) {
  S.Ld = c.dyn_layers; S.Lc = c.cost_layers;
  for (int i = 0; i <= S.Ld; ++i) S.dyn_dims[i] = dyn_dims[i];
  for (int i = 0; i <= S.Lc; ++i) S.cost_dims[i] = cost_dims[i];
  S.n = c.n; S.m = c.m; S.T = c.T; S.fout = c.cost_fout; S.num_sms = num_sms;
  S.supported = false;
  int hmax = 16;
  for (int i = 1; i < S.Ld; ++i) hmax = std::max(hmax, dyn_dims[i]);
  for (int i = 1; i < S.Lc; ++i) hmax = std::max(hmax, cost_dims[i]);
  if (S.Ld < 2) { S.why = "dynamics MLP has no hidden layer"; return GMPC_OK; }
  if (hmax > 256) { S.why = "hidden width > 256 (accumulator + split operand exceed the 512 TMEM columns)"; return GMPC_OK; }
  if (c.n + c.m > 32 || c.cost_fout > 32) { S.why = "n+m or fout > 32"; return GMPC_OK; }
  S.maxks = t_rup(hmax, 16) / 16 <= 13 ? 13 : 16;
  S.slot_bytes = (uint32_t)t_rup(hmax, 16) * 64u;  // one k-step of the widest layer
  // geometry: biases, scales, images
  int nbias = 0, nscale = 0;
  auto geom = [&](const int* dims, int Ln, TDir& F, TDir& Bw, size_t& off_f, size_t& off_b) {
    F.L = Bw.L = Ln; F.pad_ = Bw.pad_ = 0;
    for (int l = 0; l < Ln; ++l) {
      TLayer& Y = F.layer[l];
      Y.M_true = dims[l + 1]; Y.npad = t_rup(dims[l + 1], 16); Y.nks = t_rup(dims[l], 16) / 16;
      Y.kpg = std::max(1, (int)(S.slot_bytes / (Y.npad * 64)));
      Y.bias_off = nbias; nbias += Y.npad;
      Y.scale_idx = nscale + l; Y.pad_ = 0;
      Y.goff = (uint32_t)off_f; off_f += (size_t)Y.nks * Y.npad * 64;
    }
    for (int i = 0; i < Ln; ++i) {
      const int lt = Ln - 1 - i;  // the adjoint pass visits the transposed layers L-1 .. 0
      TLayer& Y = Bw.layer[i];
      Y.M_true = dims[lt]; Y.npad = t_rup(dims[lt], 16); Y.nks = t_rup(dims[lt + 1], 16) / 16;
      Y.kpg = std::max(1, (int)(S.slot_bytes / (Y.npad * 64)));
      Y.bias_off = 0; Y.scale_idx = nscale + lt; Y.pad_ = 0;
      Y.goff = (uint32_t)off_b; off_b += (size_t)Y.nks * Y.npad * 64;
    }
    nscale += Ln;
  };
  size_t sz[4] = {0, 0, 0, 0};
  geom(S.dyn_dims, S.Ld, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B], sz[DIR_DYN_F], sz[DIR_DYN_B]);
  geom(S.cost_dims, S.Lc, S.dir[DIR_COST_F], S.dir[DIR_COST_B], sz[DIR_COST_F], sz[DIR_COST_B]);
  S.nbias = nbias; S.nscale = nscale;
  const TSmem L0 = t_smem_layout(0, S.slot_bytes, nbias, nscale, c.n, c.m);
  int nslot = (int)((smem_optin - std::min(smem_optin, (size_t)L0.total)) / S.slot_bytes);
  nslot = std::min(nslot, T_MAX_SLOTS);
  if (nslot < 6) { S.why = "shared memory"; return GMPC_OK; }
  S.nslot = nslot;
  S.smem_bytes = t_smem_layout(nslot, S.slot_bytes, nbias, nscale, c.n, c.m).total;
  S.stream_bytes = sz[0] + sz[1] + sz[2] + sz[3];
  if (cudaMalloc(&S.d_stream, S.stream_bytes + 256) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_stream, 0, S.stream_bytes + 256) != cudaSuccess) return GMPC_E_CUDA;
  {
    size_t off = 0;
    for (int d = 0; d < 4; ++d) { S.dir[d].gsrc = S.d_stream + off; off += sz[d]; }
  }
  if (cudaMalloc(&S.d_bias, (size_t)std::max(nbias, 1) * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_bias, 0, (size_t)std::max(nbias, 1) * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_scale, nscale * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_absmax, nscale * sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_ovf, sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_ovf, 0, sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  const size_t G = num_sms;
  const size_t sx = (size_t)(c.T + 1) * c.n * T_NB, su = (size_t)c.T * c.m * T_NB;
  const size_t smk = ((size_t)c.T * (c.dyn_layers - 1) + (c.cost_layers - 1)) * 2 * T_EPI + 1;
  if (cudaMalloc(&S.ws_X, G * sx * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.ws_G, G * sx * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.ws_U, G * su * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.ws_M, G * su * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.ws_V, G * su * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.ws_mask, G * smk * sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaFuncSetAttribute(t128_kernel_ptr(S.maxks), cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S.smem_bytes) !=
      cudaSuccess)
    return GMPC_E_CUDA;
  S.supported = true;
  S.why = "";
  return GMPC_OK;
}

inline int t128_set_weights(T128State& S, const float* const* dyn_W, const float* const* dyn_b,
                            const float* const* cost_W, const float* const* cost_b, cudaStream_t st, int64_t* launches) {
  if (!S.supported) return GMPC_OK;
  cudaMemsetAsync(S.d_absmax, 0, S.nscale * sizeof(uint32_t), st);
  cudaMemsetAsync(S.d_stream, 0, S.stream_bytes, st);
  auto one = [&](const int* dims, int Ln, const float* const* W, const float* const* b, TDir& F, TDir& Bw, int sbase) {
    for (int l = 0; l < Ln; ++l) {
      const int K = dims[l], N = dims[l + 1], blocks = (K * N + 255) / 256;
      const TLayer& f = F.layer[l];
      const TLayer& rv = Bw.layer[Ln - 1 - l];
      h16_absmax_kernel<<<std::min(blocks, 64), 256, 0, st>>>(W[l], K * N, S.d_absmax + sbase + l);
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 0, const_cast<uint8_t*>(F.gsrc) + f.goff, f.npad,
                                               S.d_absmax + sbase + l, S.d_scale + sbase + l);
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 1, const_cast<uint8_t*>(Bw.gsrc) + rv.goff, rv.npad,
                                               S.d_absmax + sbase + l, nullptr);
      *launches += 3;
      cudaMemcpyAsync(S.d_bias + f.bias_off, b[l], sizeof(float) * N, cudaMemcpyDeviceToDevice, st);
    }
  };
  one(S.dyn_dims, S.Ld, dyn_W, dyn_b, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B], 0);
  one(S.cost_dims, S.Lc, cost_W, cost_b, S.dir[DIR_COST_F], S.dir[DIR_COST_B], S.Ld);
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

inline int t128_launch(T128State& S, const PlanParams& P, cudaStream_t st, int64_t* launches) {
  TParams Q;
  memset(&Q, 0, sizeof(Q));
  for (int d = 0; d < 4; ++d) Q.dir[d] = S.dir[d];
  Q.n = P.n; Q.m = P.m; Q.T = P.T; Q.K = P.K;
  Q.fout = P.fout; Q.mode = P.mode; Q.method = P.method; Q.iters = P.iters;
  Q.use_cost = P.use_cost; Q.final_fwd = P.final_fwd;
  Q.nslot = S.nslot; Q.slot_bytes = S.slot_bytes;
  Q.nbias = S.nbias; Q.nscale = S.nscale;
  Q.NQ = P.NQ;
  Q.ntiles = (int)((P.NQ + T_NB - 1) / T_NB);
  Q.lr = P.lr; Q.b1 = P.b1; Q.b2 = P.b2; Q.eps = P.eps;
  Q.x0 = P.x0; Q.U_in = P.U_in; Q.goal = P.goal; Q.mpcw = P.mpcw;
  Q.bias = S.d_bias; Q.inv_scale = S.d_scale;
  Q.U_out = P.U_out; Q.X_out = P.X_out; Q.J_out = P.J_out; Q.dU_out = P.dU_out; Q.lam_out = P.lam_out;
  Q.ws_X = S.ws_X; Q.ws_G = S.ws_G; Q.ws_U = S.ws_U; Q.ws_M = S.ws_M; Q.ws_V = S.ws_V;
  Q.ws_mask = S.ws_mask;
  Q.ovf = S.d_ovf;
  if (Q.ntiles <= 0) return GMPC_OK;
  const int grid = std::min(Q.ntiles, S.num_sms);
  t128_kernel_ptr(S.maxks)<<<grid, T_THREADS, S.smem_bytes, st>>>(Q);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
