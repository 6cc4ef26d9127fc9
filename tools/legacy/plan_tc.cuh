// plan_tc.cuh -- fused rollout + cost + adjoint + update kernel on the 5th-gen tensor cores
// (tcgen05.mma kind::tf32, accumulators in TMEM, weights streamed by the bulk-copy (TMA) engine).
//
// Formulation (per CTA, one tile of NB=32 trajectories at a time, persistent over tiles):
//     out^T[features x NB] = W^T[features x K] * act^T[K x NB]
// i.e. the WEIGHTS are the A operand (M = output features, 128 per block, <= 2 blocks) and the
// batch tile is the N dimension, so a 32-trajectory tile still feeds full 128-row MMAs and all 148
// SMs get work at B=4096.  Everything is K-major / SWIZZLE_NONE (pinned on hardware by
// tests/test_gpu_tc_probe.py).
//
// Precision: 3xTF32 error-compensated split, fp32 accumulation in TMEM:
//     W = Wh + Wl,  a = ah + al   (Wh, ah tf32-exact)      W a ~= Wh ah + Wh al + Wl ah
// issued as two MMAs per k-step and block: [D1 | D2] (+)= Wh x [ah ; al] (N = 64) and
// D2 += Wl x ah (N = 32); the epilogue adds the two column halves.  A single-pass TF32 product
// (10-bit mantissa) misses the 1e-4 parity bar after 32 residual steps x 20 iterations.
//
// Warp roles (320 threads): warp 0 lane 0 = weight-stage producer (cp.async.bulk + mbarrier ring),
// warp 1 lane 0 = MMA issuer, warps 2..9 = epilogue/elementwise (TMEM -> registers -> bias/ReLU/
// mask -> hi/lo split -> next layer's B operand in shared memory; per-trajectory costs, adjoint
// and Adam/gradient update).  The layer chain is strictly sequential (layer l+1 needs all of layer
// l), so one accumulator buffer and one operand buffer suffice; the weight ring runs ahead.
//
// Restates the same reference lines as plan_ffma.cuh.
#pragma once
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <vector>

#include "../../include/gmpc.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace gmpc {

constexpr int TC_NB = 32;              // trajectories per tile
constexpr int TC_THREADS = 320;        // producer warp + MMA warp + 8 compute warps
constexpr int TC_COMPUTE = 256;
constexpr int TC_SLOT_BYTES = 26624;   // weight ring slot: two k-steps of a 200-row layer (2 x 12800 B)
constexpr int TC_NSLOT = 4;
constexpr int TC_RING_PAD = 4096;      // MMA row blocks may read (never use) past a short image
constexpr int TC_SB_FEATS = 32;        // small operand buffer: q=[x;u], lambda, dy (<= 32 features)
constexpr uint32_t TC_LBO_B = 2 * TC_NB * 16 + 16;  // B operand k-chunk slab: 32 hi + 32 lo rows, +16 B pad
constexpr int TC_SROW = TC_NB + 1;     // padded row of the small fp32 arrays

struct TcLayer {
  const uint8_t* gsrc;  // packed stage stream of this layer (hi/lo images, see tc_pack_kernel)
  const float* bias;    // forward only
  int M_true;           // output features of this (possibly transposed) layer
  int red_steps;        // reduction length / 8
  int kps;              // k-steps per ring stage
  int nstages;
  int nblk;             // 128-row blocks
  int next_kpad;        // features the epilogue must define in the next operand (round_up(M_true, 8))
  uint32_t kstep_bytes, hi_bytes, lbo;
  uint32_t pad_;
};
struct TcDir {
  TcLayer layer[MAXL];
  int L;
  int pad_;
};

struct TcParams {
  TcDir dir[4];
  int n, m, T, K;
  int fout, mode, method, iters, use_cost, final_fwd, ntiles, hb_chunks;
  long long NQ;
  float lr, b1, b2, eps;
  const float *x0, *U_in, *goal, *mpcw;
  float *U_out, *X_out, *J_out, *dU_out, *lam_out;
  float *ws_X, *ws_G, *ws_U, *ws_M, *ws_V;
  uint32_t* ws_mask;
  long long* dbg;  // optional [grid][16] cycle counters (phase breakdown), nullptr = off
};

__device__ __forceinline__ int tc_pass_kind(const TcParams& P, int p) {
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

// Shared-memory carve-up (byte offsets from the 128-aligned dynamic base).
struct TcSmem {
  uint32_t ring, hb, sb, small, bars, total;
};
__host__ __device__ inline TcSmem tc_smem_layout(int hb_chunks) {
  TcSmem s;
  s.ring = 0;
  s.hb = TC_NSLOT * TC_SLOT_BYTES + TC_RING_PAD;
  s.sb = s.hb + (uint32_t)hb_chunks * TC_LBO_B;
  s.small = s.sb + (TC_SB_FEATS / 4) * TC_LBO_B;
  s.bars = s.small + 11 * TC_SB_FEATS * TC_SROW * 4;  // x_s, lam_s, dq_s, y_s, x0_s, 5 staging arrays, scalars
  s.total = s.bars + 256;
  return s;
}

// Store one activation value (already split) into a B operand buffer: trajectory nb, feature f.
__device__ __forceinline__ void tc_store_op(uint8_t* buf, int f, int nb, float v) {
  float hi, lo;
  split_tf32(v, hi, lo);
  uint8_t* p = buf + (f >> 2) * TC_LBO_B + (nb >> 3) * 128 + (nb & 7) * 16 + (f & 3) * 4;
  *reinterpret_cast<float*>(p) = hi;
  *reinterpret_cast<float*>(p + (TC_NB / 8) * 128) = lo;  // lo rows follow the 32 hi rows
}

__global__ void __launch_bounds__(TC_THREADS, 1) plan_tc_kernel(const __grid_constant__ TcParams P) {
  extern __shared__ __align__(128) uint8_t tsm[];
  const TcSmem L = tc_smem_layout(P.hb_chunks);
  uint8_t* ring = tsm + L.ring;
  uint8_t* HB = tsm + L.hb;
  uint8_t* SB = tsm + L.sb;
  float* x_s = reinterpret_cast<float*>(tsm + L.small);
  float* lam_s = x_s + TC_SB_FEATS * TC_SROW;
  float* dq_s = lam_s + TC_SB_FEATS * TC_SROW;
  float* y_s = dq_s + TC_SB_FEATS * TC_SROW;
  float* x0_s = y_s + TC_SB_FEATS * TC_SROW;
  // per-step staging of the L2-resident trajectory scratch (prefetched during the MMA phase)
  float* pu_s = x0_s + TC_SB_FEATS * TC_SROW;  // U[t]
  float* pg_s = pu_s + TC_SB_FEATS * TC_SROW;  // goal[t]
  float* px_s = pg_s + TC_SB_FEATS * TC_SROW;  // X[t]
  float* pm_s = px_s + TC_SB_FEATS * TC_SROW;  // Adam first moment [t]
  float* pv_s = pm_s + TC_SB_FEATS * TC_SROW;  // Adam second moment [t]
  float* scal_s = pv_s + TC_SB_FEATS * TC_SROW;  // per-trajectory scalars of the current step
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tsm + L.bars);
  uint64_t* empty_bar = full_bar + TC_NSLOT;
  uint64_t* acc_bar = empty_bar + TC_NSLOT;
  uint64_t* act_bar = acc_bar + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = P.n, m = P.m, T = P.T;
  // cluster: the CTAs of a cluster share every weight stage (each loads 1/C and multicasts it)
  const uint32_t C = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);
  const int n_iter = (P.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;  // uniform per cluster

  // zero all operand / scratch memory once: padded features must stay finite
  for (uint32_t i = tid * 4; i < L.bars; i += TC_THREADS * 4) *reinterpret_cast<uint32_t*>(tsm + i) = 0u;
  if (tid == 0) {
    for (int s = 0; s < TC_NSLOT; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], C);  // one tcgen05.commit arrival from every CTA of the cluster
    }
    mbar_init(acc_bar, 1);
    mbar_init(act_bar, TC_COMPUTE / 32);  // one arrival per compute warp
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_holder, 128);  // 2 blocks x (32 + 32) fp32 columns
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();  // peers' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================== weight-stage producer
    {
      uint32_t cnt = 0;
      for (int ti = 0; ti < n_iter; ++ti) {  // every CTA of a cluster streams every iteration
        for (int p = 0;; ++p) {
          const int kind = tc_pass_kind(P, p);
          if (kind == DIR_END) break;
          const TcDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TcLayer& Y = D.layer[l];
            for (int s = 0; s < Y.nstages; ++s, ++cnt) {
              const uint32_t slot = cnt % TC_NSLOT, ph = (cnt / TC_NSLOT) & 1;
              const int ks = min(Y.kps, Y.red_steps - s * Y.kps);
              const uint32_t bytes = (uint32_t)ks * Y.kstep_bytes;
              const uint32_t part = bytes / C;  // kstep_bytes is a multiple of 64
              mbar_wait(&empty_bar[slot], ph ^ 1);  // all C CTAs released the slot
              const uint8_t* src = Y.gsrc + (size_t)s * Y.kps * Y.kstep_bytes + crank * part;
              uint8_t* dst = ring + slot * TC_SLOT_BYTES + crank * part;
              if (elect_one()) {
                mbar_arrive_expect_tx(&full_bar[slot], bytes);
                if (C > 1)
                  bulk_copy_g2s_mc(dst, src, part, &full_bar[slot], cmask);
                else
                  bulk_copy_g2s(dst, src, part, &full_bar[slot]);
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    {
      uint32_t cnt = 0, act_ph = 0;
      long long t_act = 0, t_full = 0, t_issue = 0, tt;
      const uint32_t idesc64 = umma_idesc_tf32(2 * TC_NB, 0), idesc32 = umma_idesc_tf32(TC_NB, 0);
      const uint32_t ring_a = smem_u32(ring), hb_a = smem_u32(HB), sb_a = smem_u32(SB);
      for (int ti = 0; ti < n_iter; ++ti) {
        const bool live = (int)blockIdx.x + ti * (int)gridDim.x < P.ntiles;
        for (int p = 0;; ++p) {
          const int kind = tc_pass_kind(P, p);
          if (kind == DIR_END) break;
          const TcDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TcLayer& Y = D.layer[l];
            if (!live) {  // no tile this round: keep the cluster's ring protocol going
              for (int s = 0; s < Y.nstages; ++s, ++cnt) {
                const uint32_t slot = cnt % TC_NSLOT, ph = (cnt / TC_NSLOT) & 1;
                mbar_wait(&full_bar[slot], ph);
                if (elect_one()) umma_commit_mc(&empty_bar[slot], cmask);
                __syncwarp();
              }
              continue;
            }
            const uint32_t b_base = (l == 0) ? sb_a : hb_a;
            tt = clock64();
            mbar_wait(act_bar, act_ph);
            t_act += clock64() - tt;
            act_ph ^= 1;
            tc_fence_after();
            int kstep = 0;
            for (int s = 0; s < Y.nstages; ++s, ++cnt) {
              const uint32_t slot = cnt % TC_NSLOT, ph = (cnt / TC_NSLOT) & 1;
              tt = clock64();
              mbar_wait(&full_bar[slot], ph);
              t_full += clock64() - tt;
              tc_fence_after();
              const int ks = min(Y.kps, Y.red_steps - s * Y.kps);
              tt = clock64();
              if (elect_one()) {
                // descriptors advance by adding (bytes >> 4) to the low word: a handful of
                // uniform adds per MMA instead of rebuilding the 64-bit descriptor
                uint64_t ah = umma_smem_desc(ring_a + slot * TC_SLOT_BYTES, Y.lbo, 128);
                uint64_t bd = umma_smem_desc(b_base + kstep * 2 * TC_LBO_B, TC_LBO_B, 128);
                const uint64_t lo_off = Y.hi_bytes >> 4, a_step = Y.kstep_bytes >> 4;
                for (int j = 0; j < ks; ++j) {
                  const uint32_t acc = (kstep + j) > 0 ? 1u : 0u;
                  umma_tf32(tmem_base, ah, bd, idesc64, acc);                   // [D1|D2] (+)= Wh x [ah;al]
                  umma_tf32(tmem_base + TC_NB, ah + lo_off, bd, idesc32, 1u);   // D2 += Wl x ah
                  if (Y.nblk > 1) {
                    umma_tf32(tmem_base + 2 * TC_NB, ah + (2048 >> 4), bd, idesc64, acc);
                    umma_tf32(tmem_base + 3 * TC_NB, ah + (2048 >> 4) + lo_off, bd, idesc32, 1u);
                  }
                  ah += a_step;
                  bd += (2 * TC_LBO_B) >> 4;
                }
                if (C > 1)
                  umma_commit_mc(&empty_bar[slot], cmask);
                else
                  umma_commit(&empty_bar[slot]);
              }
              __syncwarp();
              t_issue += clock64() - tt;
              kstep += ks;
            }
            if (elect_one()) umma_commit(acc_bar);
            __syncwarp();
          }
        }
      }
      if (P.dbg != nullptr && lane == 0) {
        P.dbg[blockIdx.x * 16 + 0] = t_act;
        P.dbg[blockIdx.x * 16 + 1] = t_full;
        P.dbg[blockIdx.x * 16 + 2] = t_issue;
      }
    }
  } else {
    // ================================================================== epilogue / elementwise
    const int ct = tid - 64;               // 0..255
    const int q = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;      // which 16 of the 32 trajectory columns
    const int c0 = half * (TC_NB / 2);
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t acc_ph = 0;
    long long t_acc = 0, t_epi = 0, t_fin = 0, t_total0 = clock64(), tq;
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
    float* wsX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * TC_NB;
    float* wsG = P.ws_G + (size_t)blockIdx.x * (T + 1) * n * TC_NB;
    float* wsU = P.ws_U + (size_t)blockIdx.x * T * m * TC_NB;
    float* wsM = P.ws_M + (size_t)blockIdx.x * T * m * TC_NB;
    float* wsV = P.ws_V + (size_t)blockIdx.x * T * m * TC_NB;
    uint32_t* wsMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)T * (Ld - 1) + (Lc - 1)) * TC_COMPUTE;
    uint32_t* costMask = wsMask + (size_t)T * (Ld - 1) * TC_COMPUTE;

    // signal "operand of the next layer is in shared memory": every writer fences its own
    // generic-proxy stores for the async proxy, then one lane per warp arrives (count 8)
    auto publish = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(act_bar);
    };
    // write `cnt` features of trajectory ct from a [f][TC_SROW] array into the small operand
    auto sb_from = [&](const float* src, int cnt) {
      if (ct < TC_NB)
        for (int f = 0; f < cnt; ++f) tc_store_op(SB, f, ct, src[f * TC_SROW + ct]);
    };
    // same, spread over all 256 threads (caller guarantees src is visible to all of them)
    auto sb_from_all = [&](const float* src, int cnt) {
      for (int e = ct; e < cnt * TC_NB; e += TC_COMPUTE) {
        const int f = e / TC_NB, r = e - f * TC_NB;
        tc_store_op(SB, f, r, src[f * TC_SROW + r]);
      }
    };
    // hidden layer: TMEM -> (+bias, relu, mask) or (mask gate) -> hi/lo -> HB
    const int f0 = q * 32 + lane;  // this thread's feature in row block 0 (block 1: +128)
    auto hidden_epilogue = [&](const TcLayer& Y, bool fwd, uint32_t* maskp) {
      // operands that do not depend on the accumulator are fetched before the wait
      uint32_t mw = fwd ? 0u : maskp[ct];
      float bias[2] = {0.f, 0.f};
      if (fwd) {
        if (f0 < Y.M_true) bias[0] = Y.bias[f0];
        if (f0 + 128 < Y.M_true) bias[1] = Y.bias[f0 + 128];
      }
      tq = clock64();
      mbar_wait(acc_bar, acc_ph);
      t_acc += clock64() - tq;
      tq = clock64();
      acc_ph ^= 1;
      tc_fence_after();
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (b < Y.nblk) {
          uint32_t d[1][2][16];
          tmem_ld16_issue(tmem_base + t_lane + b * (2 * TC_NB) + c0, d[0][0]);
          tmem_ld16_issue(tmem_base + t_lane + b * (2 * TC_NB) + TC_NB + c0, d[0][1]);
          tmem_ld_wait();
          const int f = b * 128 + f0;
          if (f < Y.next_kpad) {
            const bool live = f < Y.M_true;
            uint8_t* base = HB + (f >> 2) * TC_LBO_B + (f & 3) * 4 + c0 * 16;
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float z = live ? (__uint_as_float(d[0][0][c]) + __uint_as_float(d[0][1][c])) + bias[b] : 0.f;
              if (fwd) {
                if (z > 0.f) mw |= 1u << (b * 16 + c);
                z = fmaxf(z, 0.f);
              } else {
                z = ((mw >> (b * 16 + c)) & 1u) ? z : 0.f;
              }
              float hi, lo;
              split_tf32(z, hi, lo);
              uint8_t* p = base + (c >> 3) * 128 + (c & 7) * 16;
              *reinterpret_cast<float*>(p) = hi;
              *reinterpret_cast<float*>(p + (TC_NB / 8) * 128) = lo;
            }
          }
        }
      }
      if (fwd) maskp[ct] = mw;
      publish();
      t_epi += clock64() - tq;
    };
    // last layer of a pass: <= 32 output features, lanes of quadrant 0 only -> small fp32 array
    auto final_epilogue = [&](const TcLayer& Y, bool fwd, float* out, bool resid, float* gout) {
      float bias = 0.f;
      if (q == 0 && fwd && lane < Y.M_true) bias = Y.bias[lane];
      tq = clock64();
      mbar_wait(acc_bar, acc_ph);
      t_fin += clock64() - tq;
      acc_ph ^= 1;
      tc_fence_after();
      if (q == 0) {
        uint32_t d1[16], d2[16];
        tmem_ld16_issue(tmem_base + t_lane + c0, d1);
        tmem_ld16_issue(tmem_base + t_lane + TC_NB + c0, d2);
        tmem_ld_wait();
        if (lane < Y.M_true) {
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float v = (__uint_as_float(d1[c]) + __uint_as_float(d2[c])) + bias;
            if (resid) v += out[lane * TC_SROW + c0 + c];
            out[lane * TC_SROW + c0 + c] = v;
            if (gout) gout[lane * TC_NB + c0 + c] = v;
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1, TC_COMPUTE);
    };

    const bool adam = (P.mode == MODE_PLAN && P.method == 1);
    const bool need_goal = cost_mode || P.mode == MODE_L2GRAD;
    // stage step t's slices of the per-CTA scratch in shared memory (all 256 threads, coalesced);
    // issued while the tensor pipe works so the per-trajectory code never waits on L2
    auto prefetch = [&](int t, bool bwd) {
      for (int e = ct; e < m * TC_NB; e += TC_COMPUTE) {
        const int j = e / TC_NB, r = e - j * TC_NB;
        pu_s[j * TC_SROW + r] = wsU[t * m * TC_NB + e];
        if (bwd && adam) {
          pm_s[j * TC_SROW + r] = wsM[t * m * TC_NB + e];
          pv_s[j * TC_SROW + r] = wsV[t * m * TC_NB + e];
        }
      }
      if (need_goal) {
        for (int e = ct; e < n * TC_NB; e += TC_COMPUTE) {
          const int i = e / TC_NB, r = e - i * TC_NB;
          pg_s[i * TC_SROW + r] = wsG[t * n * TC_NB + e];
          if (bwd) px_s[i * TC_SROW + r] = wsX[t * n * TC_NB + e];
        }
      }
    };

    for (int ti = 0; ti < n_iter; ++ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      if (tile >= P.ntiles) break;
      const long long q0 = (long long)tile * TC_NB;
      named_bar_sync(1, TC_COMPUTE);
      // ---------------------------------------------------------------- stage the tile
      for (int e = ct; e < TC_NB * n; e += TC_COMPUTE) {
        const int r = e / n, i = e - r * n;
        const long long qq = q0 + r;
        x0_s[i * TC_SROW + r] = (qq < P.NQ) ? P.x0[(qq / P.K) * n + i] : 0.f;
      }
      if (P.goal != nullptr) {
        const int per = (T + 1) * n;
        for (int e = ct; e < TC_NB * per; e += TC_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          wsG[rest * TC_NB + r] = (qq < P.NQ) ? P.goal[(qq / P.K) * per + rest] : 0.f;
        }
      }
      {
        const int per = T * m;
        for (int e = ct; e < TC_NB * per; e += TC_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          wsU[rest * TC_NB + r] = (qq < P.NQ) ? P.U_in[qq * per + rest] : 0.f;
          if (P.mode == MODE_PLAN && P.method == 1) {
            wsM[rest * TC_NB + r] = 0.f;
            wsV[rest * TC_NB + r] = 0.f;
          }
        }
      }
      named_bar_sync(1, TC_COMPUTE);
      const long long qr = q0 + ct;
      const bool rvalid = (ct < TC_NB) && (qr < P.NQ);
      float Jr = 0.f;

      for (int it = 0;; ++it) {
        const bool last = (it == P.iters);
        if (last && !P.final_fwd) break;
        // -------------------------------------------------------------- forward rollout
        named_bar_sync(1, TC_COMPUTE);  // the previous sweep's last update (warp 2) is complete
        for (int e = ct; e < n * TC_NB; e += TC_COMPUTE) {
          const int i = e / TC_NB, r = e - i * TC_NB;
          x_s[i * TC_SROW + r] = x0_s[i * TC_SROW + r];
          wsX[e] = x0_s[i * TC_SROW + r];
        }
        Jr = 0.f;
        prefetch(0, false);
        named_bar_sync(1, TC_COMPUTE);
        for (int t = 0; t < T; ++t) {
          // q = [x ; u_t] -> operand buffer, all threads; staging cost by one thread per trajectory
          for (int e = ct; e < (n + m) * TC_NB; e += TC_COMPUTE) {
            const int f = e / TC_NB, r = e - f * TC_NB;
            tc_store_op(SB, f, r, f < n ? x_s[f * TC_SROW + r] : pu_s[(f - n) * TC_SROW + r]);
          }
          if (ct < TC_NB && need_goal) {
            const int r = ct;
            float uu = 0.f, dd = 0.f;
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
              const float u = pu_s[j * TC_SROW + r];
              uu = fmaf(u, u, uu);
            }
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
              const float d = x_s[i * TC_SROW + r] - pg_s[i * TC_SROW + r];
              dd = fmaf(d, d, dd);
            }
            if (cost_mode)
              Jr += w0 * (sqrtf(uu + a2) - ALPHA) + w1 * (sqrtf(dd + a2) - ALPHA);
            else
              Jr += dd;
          }
          publish();
          const TcDir& D = P.dir[DIR_DYN_F];
          if (D.L == 1) named_bar_sync(1, TC_COMPUTE);  // staging is free once warp 2 is past the pre-step
          for (int l = 0; l < D.L - 1; ++l) {
            hidden_epilogue(D.layer[l], true, wsMask + ((size_t)t * (Ld - 1) + l) * TC_COMPUTE);
            // every warp has published layer 0's operand => the pre-step reads are done: refill
            if (l == 0 && t + 1 < T) prefetch(t + 1, false);
          }
          if (D.L == 1 && t + 1 < T) prefetch(t + 1, false);
          final_epilogue(D.layer[D.L - 1], true, x_s, true, wsX + (size_t)(t + 1) * n * TC_NB);
        }
        // -------------------------------------------------------------- terminal cost
        if (P.use_cost) {
          sb_from_all(x_s, n);
          publish();
          const TcDir& D = P.dir[DIR_COST_F];
          for (int l = 0; l < D.L - 1; ++l)
            hidden_epilogue(D.layer[l], true, costMask + (size_t)l * TC_COMPUTE);
          final_epilogue(D.layer[D.L - 1], true, y_s, false, nullptr);
          if (ct < TC_NB) {
            float yy = 0.f;
            const float s = 2.f * w2;
            for (int o = 0; o < P.fout; ++o) {
              const float y = y_s[o * TC_SROW + ct];
              yy = fmaf(y, y, yy);
              y_s[o * TC_SROW + ct] = s * y;
            }
            Jr += w2 * yy;
          }
        } else if (P.mode == MODE_L2GRAD) {
          if (ct < TC_NB) {
            float dd = 0.f;
            for (int i = 0; i < n; ++i) {
              const float d = x_s[i * TC_SROW + ct] - wsG[(T * n + i) * TC_NB + ct];
              dd = fmaf(d, d, dd);
            }
            Jr = (Jr + dd) / (float)(T + 1);
          }
        }
        if (last) break;
        // -------------------------------------------------------------- adjoint seed lambda_T
        if (P.use_cost) {
          sb_from(y_s, P.fout);
          publish();
          const TcDir& D = P.dir[DIR_COST_B];
          for (int lb = 0; lb < D.L - 1; ++lb)
            hidden_epilogue(D.layer[lb], false, costMask + (size_t)(D.L - 2 - lb) * TC_COMPUTE);
          final_epilogue(D.layer[D.L - 1], false, lam_s, false, nullptr);
        } else if (ct < TC_NB) {
          for (int i = 0; i < n; ++i)
            lam_s[i * TC_SROW + ct] = l2scale * (x_s[i * TC_SROW + ct] - wsG[(T * n + i) * TC_NB + ct]);
        }
        if (!P.use_cost) named_bar_sync(1, TC_COMPUTE);  // L2 seed was written by warp 2 only
        for (int e = ct; e < n * TC_NB; e += TC_COMPUTE) {
          const int i = e / TC_NB, r = e - i * TC_NB;
          const float lam = lam_s[i * TC_SROW + r];
          tc_store_op(SB, i, r, lam);
          if (P.lam_out != nullptr && q0 + r < P.NQ) P.lam_out[((q0 + r) * (T + 1) + T) * n + i] = lam;
        }
        float bc1 = 1.f, bc2 = 1.f;
        if (adam) {
          bc1 = (float)(1.0 - pow((double)P.b1, (double)(it + 1)));
          bc2 = (float)(1.0 - pow((double)P.b2, (double)(it + 1)));
        }
        // -------------------------------------------------------------- adjoint sweep + update
        for (int t = T - 1; t >= 0; --t) {
          publish();  // lambda_{t+1} is in the operand buffer
          const TcDir& D = P.dir[DIR_DYN_B];
          if (D.L == 1) named_bar_sync(1, TC_COMPUTE);
          for (int lb = 0; lb < D.L - 1; ++lb) {
            hidden_epilogue(D.layer[lb], false,
                            wsMask + ((size_t)t * (Ld - 1) + (D.L - 2 - lb)) * TC_COMPUTE);
            if (lb == 0) prefetch(t, true);  // every warp is past the previous step's update
          }
          if (D.L == 1) prefetch(t, true);
          final_epilogue(D.layer[D.L - 1], false, dq_s, false, nullptr);
          // phase 1: per-trajectory norms of the staging cost (one thread per trajectory)
          if (ct < TC_NB && cost_mode) {
            const int r = ct;
            float uu = 0.f, dd = 0.f;
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
              const float u = pu_s[j * TC_SROW + r];
              uu = fmaf(u, u, uu);
            }
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
              const float d = px_s[i * TC_SROW + r] - pg_s[i * TC_SROW + r];
              dd = fmaf(d, d, dd);
            }
            scal_s[r] = sqrtf(uu + a2);
            scal_s[TC_SROW + r] = sqrtf(dd + a2);
          }
          named_bar_sync(1, TC_COMPUTE);
          // phase 2: action gradient + update and adjoint update, one (feature, trajectory) per thread
          for (int e = ct; e < m * TC_NB; e += TC_COMPUTE) {
            const int j = e / TC_NB, r = e - j * TC_NB;
            const int ix = t * m * TC_NB + e;
            float u = pu_s[j * TC_SROW + r];
            float g = dq_s[(n + j) * TC_SROW + r];
            if (cost_mode) g = (w0 * u) / scal_s[r] + g;
            if (P.mode == MODE_PLAN) {
              if (P.method == 0) {
                u = u - P.lr * g;
              } else {
                const float mo = P.b1 * pm_s[j * TC_SROW + r] + (1.f - P.b1) * g;
                const float ve = P.b2 * pv_s[j * TC_SROW + r] + (1.f - P.b2) * g * g;
                wsM[ix] = mo;
                wsV[ix] = ve;
                u = u - P.lr * (mo / bc1) / (sqrtf(ve / bc2) + P.eps);
              }
              wsU[ix] = u;
            } else if (P.dU_out != nullptr && q0 + r < P.NQ) {
              P.dU_out[((q0 + r) * T + t) * m + j] = g;
            }
          }
          for (int e = ct; e < n * TC_NB; e += TC_COMPUTE) {
            const int i = e / TC_NB, r = e - i * TC_NB;
            const float d = px_s[i * TC_SROW + r] - pg_s[i * TC_SROW + r];
            const float c = cost_mode ? (w1 * d) / scal_s[TC_SROW + r] : l2scale * d;
            const float lam = (c + lam_s[i * TC_SROW + r]) + dq_s[i * TC_SROW + r];
            lam_s[i * TC_SROW + r] = lam;
            tc_store_op(SB, i, r, lam);  // operand of the next adjoint step
            if (P.lam_out != nullptr && q0 + r < P.NQ) P.lam_out[((q0 + r) * (T + 1) + t) * n + i] = lam;
          }
        }
        if (P.mode != MODE_PLAN) break;
      }
      named_bar_sync(1, TC_COMPUTE);
      // ---------------------------------------------------------------- write the tile out
      if (P.J_out != nullptr && rvalid) P.J_out[qr] = Jr;
      if (P.U_out != nullptr) {
        const int per = T * m;
        for (int e = ct; e < TC_NB * per; e += TC_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          if (qq < P.NQ) P.U_out[qq * per + rest] = wsU[rest * TC_NB + r];
        }
      }
      if (P.X_out != nullptr) {
        const int per = (T + 1) * n;
        for (int e = ct; e < TC_NB * per; e += TC_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          if (qq < P.NQ) P.X_out[qq * per + rest] = wsX[rest * TC_NB + r];
        }
      }
    }
    if (P.dbg != nullptr && ct == 0) {
      P.dbg[blockIdx.x * 16 + 4] = t_acc;
      P.dbg[blockIdx.x * 16 + 5] = t_epi;
      P.dbg[blockIdx.x * 16 + 6] = t_fin;
      P.dbg[blockIdx.x * 16 + 7] = clock64() - t_total0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 128);
  }
}

// Pack one Dense kernel W[K][N] (flax layout) into the stage stream of one direction.
//   transposed == 0 (forward):  A rows r = output feature n, reduction kk = input feature k.
//   transposed == 1 (adjoint):  A rows r = input feature k,  reduction kk = output feature n.
// k-step image: hi = [2 k-chunks][rows_pad][4 floats], lo follows at +hi_bytes.  Pre-zeroed.
__global__ void tc_pack_kernel(const float* __restrict__ W, int K, int N, int transposed,
                               uint8_t* dst, uint32_t kstep_bytes, uint32_t hi_bytes,
                               uint32_t lbo) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int k = idx / N, o = idx - k * N;
  const int r = transposed ? k : o, kk = transposed ? o : k;
  float hi, lo;
  split_tf32(W[idx], hi, lo);
  uint8_t* p = dst + (size_t)(kk >> 3) * kstep_bytes + ((kk & 7) >> 2) * lbo + (r >> 3) * 128 +
               (r & 7) * 16 + (kk & 3) * 4;
  *reinterpret_cast<float*>(p) = hi;
  *reinterpret_cast<float*>(p + hi_bytes) = lo;
}

// ------------------------------------------------------------------------------------- host side
struct TcState {
  bool supported = false;
  std::string why = "not initialised";
  int dyn_dims[MAXL + 1], cost_dims[MAXL + 1], Ld = 0, Lc = 0;
  uint8_t* d_stream = nullptr;
  size_t stream_bytes = 0;
  TcDir dir[4];
  float* d_bias = nullptr;  // forward biases, packed
  int hb_chunks = 0;
  size_t smem_bytes = 0;
  int num_sms = 0;
  int cluster = 4;                 // requested CTAs per cluster (1, 2 or 4)
  int max_clusters[5] = {0, 0, 0, 0, 0};  // co-resident clusters per cluster size
  int last_cluster = 1;
  long long* d_dbg = nullptr;  // GMPC_DEBUG: per-CTA phase cycle counters
};

inline int rup(int v, int a) { return (v + a - 1) / a * a; }

inline void tc_layer_geom(TcLayer& Y, int M_true, int red_true) {
  const int rows_pad = rup(M_true, 8);
  Y.M_true = M_true;
  Y.red_steps = rup(red_true, 8) / 8;
  Y.lbo = (uint32_t)rows_pad * 16;
  Y.hi_bytes = 2 * Y.lbo;
  Y.kstep_bytes = 2 * Y.hi_bytes;
  Y.kps = std::max(1, (int)(TC_SLOT_BYTES / Y.kstep_bytes));
  Y.nstages = (Y.red_steps + Y.kps - 1) / Y.kps;
  Y.nblk = (M_true + 127) / 128;
  Y.next_kpad = rup(M_true, 8);
  Y.bias = nullptr;
  Y.gsrc = nullptr;
  Y.pad_ = 0;
}

// Build geometry for both MLPs; returns total stream bytes.  dir[DIR_*_B].layer[i] is the
// transposed layer L-1-i.
inline size_t tc_build_geometry(TcState& S) {
  size_t off = 0;
  auto one = [&](const int* dims, int Ln, TcDir& F, TcDir& Bw) {
    F.L = Bw.L = Ln;
    F.pad_ = Bw.pad_ = 0;
    for (int l = 0; l < Ln; ++l) {
      tc_layer_geom(F.layer[l], dims[l + 1], dims[l]);
      F.layer[l].gsrc = reinterpret_cast<const uint8_t*>(off);
      off += (size_t)F.layer[l].red_steps * F.layer[l].kstep_bytes;
    }
    for (int i = 0; i < Ln; ++i) {
      const int lt = Ln - 1 - i;
      tc_layer_geom(Bw.layer[i], dims[lt], dims[lt + 1]);
      Bw.layer[i].gsrc = reinterpret_cast<const uint8_t*>(off);
      off += (size_t)Bw.layer[i].red_steps * Bw.layer[i].kstep_bytes;
    }
  };
  one(S.dyn_dims, S.Ld, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B]);
  one(S.cost_dims, S.Lc, S.dir[DIR_COST_F], S.dir[DIR_COST_B]);
  return off;
}

inline int tc_create(TcState& S, const gmpc_config& c, const int* dyn_dims, const int* cost_dims,
                     const cudaDeviceProp& prop) {
  S.Ld = c.dyn_layers;
  S.Lc = c.cost_layers;
  for (int i = 0; i <= S.Ld; ++i) S.dyn_dims[i] = dyn_dims[i];
  for (int i = 0; i <= S.Lc; ++i) S.cost_dims[i] = cost_dims[i];
  S.num_sms = prop.multiProcessorCount;
  int hmax = 8;
  for (int i = 1; i < S.Ld; ++i) hmax = std::max(hmax, dyn_dims[i]);
  for (int i = 1; i < S.Lc; ++i) hmax = std::max(hmax, cost_dims[i]);
  S.supported = false;
  if (hmax > 256) { S.why = "hidden width > 256 (two 128-row MMA blocks)"; return GMPC_OK; }
  if (c.n + c.m > TC_SB_FEATS || c.cost_fout > TC_SB_FEATS) {
    S.why = "n+m or fout > 32";
    return GMPC_OK;
  }
  S.hb_chunks = rup(hmax, 8) / 4;
  const TcSmem L = tc_smem_layout(S.hb_chunks);
  S.smem_bytes = L.total + 128;
  if (S.smem_bytes > (size_t)prop.sharedMemPerBlockOptin) { S.why = "shared memory"; return GMPC_OK; }
  S.stream_bytes = tc_build_geometry(S);
  size_t nbias = 0;
  for (int l = 0; l < S.Ld; ++l) nbias += rup(dyn_dims[l + 1], 4);
  for (int l = 0; l < S.Lc; ++l) nbias += rup(cost_dims[l + 1], 4);
  if (cudaMalloc(&S.d_stream, S.stream_bytes + TC_RING_PAD) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_stream, 0, S.stream_bytes + TC_RING_PAD) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_bias, nbias * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  // turn offsets into pointers, assign biases
  float* bp = S.d_bias;
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < S.dir[d].L; ++l)
      S.dir[d].layer[l].gsrc = S.d_stream + reinterpret_cast<size_t>(S.dir[d].layer[l].gsrc);
  for (int l = 0; l < S.Ld; ++l) { S.dir[DIR_DYN_F].layer[l].bias = bp; bp += rup(dyn_dims[l + 1], 4); }
  for (int l = 0; l < S.Lc; ++l) { S.dir[DIR_COST_F].layer[l].bias = bp; bp += rup(cost_dims[l + 1], 4); }
  if (cudaFuncSetAttribute(plan_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)S.smem_bytes) != cudaSuccess)
    return GMPC_E_CUDA;
  for (int C = 2; C <= 4; C *= 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(prop.multiProcessorCount / C * C);
    cfg.blockDim = dim3(TC_THREADS);
    cfg.dynamicSmemBytes = S.smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, plan_tc_kernel, &cfg) != cudaSuccess) nc = 0;
    S.max_clusters[C] = nc;
  }
  if (const char* env = getenv("GMPC_TC_CLUSTER")) S.cluster = atoi(env);
  if (S.cluster != 1 && S.cluster != 2 && S.cluster != 4) S.cluster = 4;
  while (S.cluster > 1 && S.max_clusters[S.cluster] <= 0) S.cluster >>= 1;
  if (getenv("GMPC_DEBUG")) {
    cudaMalloc(&S.d_dbg, sizeof(long long) * 16 * 1024);
    cudaMemset(S.d_dbg, 0, sizeof(long long) * 16 * 1024);
  }
  if (getenv("GMPC_DEBUG"))
    fprintf(stderr, "[gmpc] tc: smem %zu B, max co-resident clusters: x2=%d x4=%d, using cluster=%d (%s)\n",
            S.smem_bytes, S.max_clusters[2], S.max_clusters[4], S.cluster, cudaGetErrorString(cudaGetLastError()));
  S.supported = true;
  S.why = "";
  return GMPC_OK;
}

inline void tc_destroy(TcState& S) {
  cudaFree(S.d_stream);
  cudaFree(S.d_bias);
  S.d_stream = nullptr;
  S.d_bias = nullptr;
}

inline int tc_set_weights(TcState& S, const float* const* dyn_W, const float* const* dyn_b,
                          const float* const* cost_W, const float* const* cost_b, cudaStream_t st,
                          int64_t* launches) {
  if (!S.supported) return GMPC_OK;
  auto one = [&](const int* dims, int Ln, const float* const* W, const float* const* b, TcDir& F,
                 TcDir& Bw) {
    for (int l = 0; l < Ln; ++l) {
      const int K = dims[l], N = dims[l + 1], blocks = (K * N + 255) / 256;
      const TcLayer& f = F.layer[l];
      const TcLayer& r = Bw.layer[Ln - 1 - l];
      tc_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 0, const_cast<uint8_t*>(f.gsrc),
                                             f.kstep_bytes, f.hi_bytes, f.lbo);
      tc_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 1, const_cast<uint8_t*>(r.gsrc),
                                             r.kstep_bytes, r.hi_bytes, r.lbo);
      *launches += 2;
      cudaMemcpyAsync(const_cast<float*>(f.bias), b[l], sizeof(float) * N, cudaMemcpyDeviceToDevice, st);
    }
  };
  one(S.dyn_dims, S.Ld, dyn_W, dyn_b, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B]);
  one(S.cost_dims, S.Lc, cost_W, cost_b, S.dir[DIR_COST_F], S.dir[DIR_COST_B]);
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

// The tensor-core path needs a real dense contraction: at least two full tiles of trajectories
// (batch tile >= 64) and hidden width >= 64 (north star).
inline bool tc_worthwhile(const TcState& S, int64_t NQ) {
  int hmin = 1 << 30;
  for (int i = 1; i < S.Ld; ++i) hmin = std::min(hmin, S.dyn_dims[i]);
  return NQ >= 64 && S.Ld > 1 && hmin >= 64;
}

inline int tc_launch(TcState& S, const PlanParams& P, cudaStream_t st, int64_t* launches) {
  TcParams Q;
  memset(&Q, 0, sizeof(Q));
  for (int d = 0; d < 4; ++d) Q.dir[d] = S.dir[d];
  Q.n = P.n; Q.m = P.m; Q.T = P.T; Q.K = P.K;
  Q.fout = P.fout; Q.mode = P.mode; Q.method = P.method; Q.iters = P.iters;
  Q.use_cost = P.use_cost; Q.final_fwd = P.final_fwd;
  Q.hb_chunks = S.hb_chunks;
  Q.NQ = P.NQ;
  Q.ntiles = (int)((P.NQ + TC_NB - 1) / TC_NB);
  Q.lr = P.lr; Q.b1 = P.b1; Q.b2 = P.b2; Q.eps = P.eps;
  Q.x0 = P.x0; Q.U_in = P.U_in; Q.goal = P.goal; Q.mpcw = P.mpcw;
  Q.U_out = P.U_out; Q.X_out = P.X_out; Q.J_out = P.J_out; Q.dU_out = P.dU_out; Q.lam_out = P.lam_out;
  Q.ws_X = P.ws_X; Q.ws_G = P.ws_G; Q.ws_U = P.ws_U; Q.ws_M = P.ws_M; Q.ws_V = P.ws_V;
  Q.ws_mask = P.ws_mask;
  Q.dbg = S.d_dbg;
  if (Q.ntiles <= 0) return GMPC_OK;
  // cluster size: share each weight stage among up to 4 CTAs (multicast) when there are tiles for them
  int C = S.cluster;
  while (C > 1 && Q.ntiles < C) C >>= 1;
  const int max_ctas = C > 1 ? S.max_clusters[C] * C : S.num_sms;
  int grid = std::min((Q.ntiles + C - 1) / C * C, max_ctas);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(TC_THREADS);
  cfg.dynamicSmemBytes = S.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, plan_tc_kernel, Q);
  ++*launches;
  S.last_cluster = C;
  if (S.d_dbg != nullptr && e == cudaSuccess) {
    long long hdbg[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hdbg, S.d_dbg, sizeof(hdbg), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gmpc] tc CTA0 cycles: mma-warp wait_act %lld wait_full %lld issue %lld | compute wait_acc(hidden) %lld epilogue %lld wait_acc(final) %lld total %lld\n",
            hdbg[0], hdbg[1], hdbg[2], hdbg[4], hdbg[5], hdbg[6], hdbg[7]);
  }
  return e == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
