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

#include <string>
#include <vector>

#include "../../include/gmpc.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace gmpc {

constexpr int TC_NB = 32;              // trajectories per tile
constexpr int TC_THREADS = 320;        // producer warp + MMA warp + 8 compute warps
constexpr int TC_COMPUTE = 256;
constexpr int TC_SLOT_BYTES = 16384;   // weight ring slot
constexpr int TC_NSLOT = 6;
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
  s.bars = s.small + 5 * TC_SB_FEATS * TC_SROW * 4;  // x_s, lam_s, dq_s, y_s, x0_s
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
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(tsm + L.bars);
  uint64_t* empty_bar = full_bar + TC_NSLOT;
  uint64_t* acc_bar = empty_bar + TC_NSLOT;
  uint64_t* act_bar = acc_bar + 1;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_bar + 1);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = P.n, m = P.m, T = P.T;

  // zero all operand / scratch memory once: padded features must stay finite
  for (uint32_t i = tid * 4; i < L.bars; i += TC_THREADS * 4) *reinterpret_cast<uint32_t*>(tsm + i) = 0u;
  if (tid == 0) {
    for (int s = 0; s < TC_NSLOT; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(acc_bar, 1);
    mbar_init(act_bar, TC_COMPUTE);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_holder, 128);  // 2 blocks x (32 + 32) fp32 columns
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0) {
    // ================================================================== weight-stage producer
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
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
              mbar_wait(&empty_bar[slot], ph ^ 1);
              mbar_arrive_expect_tx(&full_bar[slot], bytes);
              bulk_copy_g2s(ring + slot * TC_SLOT_BYTES,
                            Y.gsrc + (size_t)s * Y.kps * Y.kstep_bytes, bytes, &full_bar[slot]);
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================================================== MMA issuer
    if (lane == 0) {
      uint32_t cnt = 0, act_ph = 0;
      const uint32_t idesc64 = umma_idesc_tf32(2 * TC_NB, 0), idesc32 = umma_idesc_tf32(TC_NB, 0);
      const uint32_t ring_a = smem_u32(ring), hb_a = smem_u32(HB), sb_a = smem_u32(SB);
      for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
        for (int p = 0;; ++p) {
          const int kind = tc_pass_kind(P, p);
          if (kind == DIR_END) break;
          const TcDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TcLayer& Y = D.layer[l];
            const uint32_t b_base = (l == 0) ? sb_a : hb_a;
            mbar_wait(act_bar, act_ph);
            act_ph ^= 1;
            tc_fence_after();
            int kstep = 0;
            for (int s = 0; s < Y.nstages; ++s, ++cnt) {
              const uint32_t slot = cnt % TC_NSLOT, ph = (cnt / TC_NSLOT) & 1;
              mbar_wait(&full_bar[slot], ph);
              tc_fence_after();
              const int ks = min(Y.kps, Y.red_steps - s * Y.kps);
              for (int j = 0; j < ks; ++j, ++kstep) {
                const uint32_t a0 = ring_a + slot * TC_SLOT_BYTES + j * Y.kstep_bytes;
                const uint64_t bd = umma_smem_desc(b_base + kstep * 2 * TC_LBO_B, TC_LBO_B, 128);
                for (int b = 0; b < Y.nblk; ++b) {
                  const uint64_t ah = umma_smem_desc(a0 + b * 2048, Y.lbo, 128);
                  const uint64_t al = umma_smem_desc(a0 + Y.hi_bytes + b * 2048, Y.lbo, 128);
                  const uint32_t d = tmem_base + b * (2 * TC_NB);
                  umma_tf32(d, ah, bd, idesc64, kstep > 0 ? 1u : 0u);  // [D1|D2] (+)= Wh x [ah;al]
                  umma_tf32(d + TC_NB, al, bd, idesc32, 1u);           // D2 += Wl x ah
                }
              }
              umma_commit(&empty_bar[slot]);
            }
            umma_commit(acc_bar);
          }
        }
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

    // signal "operand of the next layer is in shared memory" (all 256 compute threads)
    auto publish = [&]() {
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(act_bar);
    };
    // write `cnt` features of trajectory ct from a [f][TC_SROW] array into the small operand
    auto sb_from = [&](const float* src, int cnt) {
      if (ct < TC_NB)
        for (int f = 0; f < cnt; ++f) tc_store_op(SB, f, ct, src[f * TC_SROW + ct]);
    };
    // hidden layer: TMEM -> (+bias, relu, mask) or (mask gate) -> hi/lo -> HB
    auto hidden_epilogue = [&](const TcLayer& Y, bool fwd, uint32_t* maskp) {
      mbar_wait(acc_bar, acc_ph);
      acc_ph ^= 1;
      tc_fence_after();
      uint32_t mw = fwd ? 0u : maskp[ct];
      for (int b = 0; b < Y.nblk; ++b) {
        const int f = b * 128 + q * 32 + lane;
        float d1[16], d2[16];
        tmem_ld16(tmem_base + t_lane + b * (2 * TC_NB) + c0, d1);
        tmem_ld16(tmem_base + t_lane + b * (2 * TC_NB) + TC_NB + c0, d2);
        if (f < Y.next_kpad) {
          const bool live = f < Y.M_true;
          const float bias = (fwd && live) ? Y.bias[f] : 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float z = live ? (d1[c] + d2[c]) + bias : 0.f;
            if (fwd) {
              if (z > 0.f) mw |= 1u << (b * 16 + c);
              z = fmaxf(z, 0.f);
            } else {
              z = ((mw >> (b * 16 + c)) & 1u) ? z : 0.f;
            }
            tc_store_op(HB, f, c0 + c, z);
          }
        }
      }
      if (fwd) maskp[ct] = mw;
      publish();
    };
    // last layer of a pass: <= 32 output features, lanes of quadrant 0 only -> small fp32 array
    auto final_epilogue = [&](const TcLayer& Y, bool fwd, float* out, bool resid, float* gout) {
      mbar_wait(acc_bar, acc_ph);
      acc_ph ^= 1;
      tc_fence_after();
      if (q == 0) {
        float d1[16], d2[16];
        tmem_ld16(tmem_base + t_lane + c0, d1);
        tmem_ld16(tmem_base + t_lane + TC_NB + c0, d2);
        if (lane < Y.M_true) {
          const float bias = fwd ? Y.bias[lane] : 0.f;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float v = (d1[c] + d2[c]) + bias;
            if (resid) v += out[lane * TC_SROW + c0 + c];
            out[lane * TC_SROW + c0 + c] = v;
            if (gout) gout[lane * TC_NB + c0 + c] = v;
          }
        }
      }
      tc_fence_before();
      named_bar_sync(1, TC_COMPUTE);
    };

    for (int tile = blockIdx.x; tile < P.ntiles; tile += gridDim.x) {
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
        for (int e = ct; e < n * TC_NB; e += TC_COMPUTE) {
          const int i = e / TC_NB, r = e - i * TC_NB;
          x_s[i * TC_SROW + r] = x0_s[i * TC_SROW + r];
          wsX[e] = x0_s[i * TC_SROW + r];
        }
        Jr = 0.f;
        named_bar_sync(1, TC_COMPUTE);
        for (int t = 0; t < T; ++t) {
          if (ct < TC_NB) {
            const int r = ct;
            float uu = 0.f;
#pragma unroll 4
            for (int j = 0; j < m; ++j) {
              const float u = wsU[(t * m + j) * TC_NB + r];
              tc_store_op(SB, n + j, r, u);
              uu = fmaf(u, u, uu);
            }
            float dd = 0.f;
#pragma unroll 4
            for (int i = 0; i < n; ++i) {
              const float x = x_s[i * TC_SROW + r];
              tc_store_op(SB, i, r, x);
              if (cost_mode || P.mode == MODE_L2GRAD) {
                const float d = x - wsG[(t * n + i) * TC_NB + r];
                dd = fmaf(d, d, dd);
              }
            }
            if (cost_mode)
              Jr += w0 * (sqrtf(uu + a2) - ALPHA) + w1 * (sqrtf(dd + a2) - ALPHA);
            else if (P.mode == MODE_L2GRAD)
              Jr += dd;
          }
          publish();
          const TcDir& D = P.dir[DIR_DYN_F];
          for (int l = 0; l < D.L - 1; ++l)
            hidden_epilogue(D.layer[l], true, wsMask + ((size_t)t * (Ld - 1) + l) * TC_COMPUTE);
          final_epilogue(D.layer[D.L - 1], true, x_s, true, wsX + (size_t)(t + 1) * n * TC_NB);
        }
        // -------------------------------------------------------------- terminal cost
        if (P.use_cost) {
          sb_from(x_s, n);
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
        if (P.lam_out != nullptr && rvalid)
          for (int i = 0; i < n; ++i) P.lam_out[(qr * (T + 1) + T) * n + i] = lam_s[i * TC_SROW + ct];
        float bc1 = 1.f, bc2 = 1.f;
        if (P.mode == MODE_PLAN && P.method == 1) {
          bc1 = (float)(1.0 - pow((double)P.b1, (double)(it + 1)));
          bc2 = (float)(1.0 - pow((double)P.b2, (double)(it + 1)));
        }
        // -------------------------------------------------------------- adjoint sweep + update
        for (int t = T - 1; t >= 0; --t) {
          sb_from(lam_s, n);
          publish();
          const TcDir& D = P.dir[DIR_DYN_B];
          for (int lb = 0; lb < D.L - 1; ++lb)
            hidden_epilogue(D.layer[lb], false,
                            wsMask + ((size_t)t * (Ld - 1) + (D.L - 2 - lb)) * TC_COMPUTE);
          final_epilogue(D.layer[D.L - 1], false, dq_s, false, nullptr);
          if (ct < TC_NB) {
            const int r = ct;
            float su = 0.f, sd = 0.f;
            if (cost_mode) {
              float uu = 0.f, dd = 0.f;
#pragma unroll 4
              for (int j = 0; j < m; ++j) {
                const float u = wsU[(t * m + j) * TC_NB + r];
                uu = fmaf(u, u, uu);
              }
#pragma unroll 4
              for (int i = 0; i < n; ++i) {
                const float d = wsX[(t * n + i) * TC_NB + r] - wsG[(t * n + i) * TC_NB + r];
                dd = fmaf(d, d, dd);
              }
              su = sqrtf(uu + a2);
              sd = sqrtf(dd + a2);
            }
#pragma unroll 2
            for (int j = 0; j < m; ++j) {
              const int ix = (t * m + j) * TC_NB + r;
              float u = wsU[ix];
              float g = dq_s[(n + j) * TC_SROW + r];
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
#pragma unroll 2
            for (int i = 0; i < n; ++i) {
              const float d = wsX[(t * n + i) * TC_NB + r] - wsG[(t * n + i) * TC_NB + r];
              const float c = cost_mode ? (w1 * d) / sd : l2scale * d;
              const float lam = (c + lam_s[i * TC_SROW + r]) + dq_s[i * TC_SROW + r];
              lam_s[i * TC_SROW + r] = lam;
              if (P.lam_out != nullptr && rvalid) P.lam_out[(qr * (T + 1) + t) * n + i] = lam;
            }
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
  }
  tc_fence_before();
  __syncthreads();
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
  const int grid = std::min(Q.ntiles, S.num_sms);
  if (grid <= 0) return GMPC_OK;
  plan_tc_kernel<<<grid, TC_THREADS, S.smem_bytes, st>>>(Q);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
