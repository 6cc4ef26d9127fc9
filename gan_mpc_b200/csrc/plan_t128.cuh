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
constexpr int T_FRONT = 3;                         // producer warp + two MMA issuer warps (N-part 0 / N-part 1)
constexpr int T_THREADS = T_EPI + 32 * T_FRONT;
constexpr int T_MAXKS = 16;                        // k-steps of the widest operand (hidden <= 256)
constexpr int T_PKS = 8;                           // k-steps one N-part of a layer defines (half of T_MAXKS)
constexpr int T_MAX_SLOTS = 32;
constexpr uint32_t T_D_COL = 0, T_AH_COL = 256, T_AL_COL = 384;  // TMEM columns: D | A hi | A lo

// A layer's MMAs go out as two N-parts (output columns [0, n0) for every k-step, then [n0, npad)): the
// accumulator of part 0 is complete, read out and processed while the tensor pipe works on part 1.
struct TLayer {
  uint32_t goff[2];        // byte offset of the tiles of N-part p in the image of the pass
  int ncol[2];             // output columns (MMA N) of N-part p; ncol[1] == 0: single part (<= 32 columns)
  int kpg[2];              // k-steps per ring group (one bulk copy, one slot) of part p
  int M_true;              // output features of this (possibly transposed) layer
  int npad;                // round_up(M_true, 16) = ncol[0] + ncol[1]
  int nks;                 // reduction k-steps (round_up(K, 16) / 16)
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
  long long* dbg;          // cycle counters of CTA 0 (builds with -DGMPC_T128_TIMED only)
};

// The pass schedule: iters x {T dyn fwd, [cost fwd, cost bwd], T dyn bwd}, then the final evaluation.
#ifdef GMPC_T128_TIMED
#define T128_TRACE(layer_no, slot) do { if (blockIdx.x == 0 && P.dbg != nullptr && (layer_no) >= 2000 && (layer_no) < 2003) \
    P.dbg[256 + ((layer_no) - 2000) * 128 + (slot)] = clock64(); } while (0)
#else
#define T128_TRACE(layer_no, slot)
#endif

#ifdef GMPC_T128_TIMED
__device__ __forceinline__ void t_wait_dbg(uint32_t bar, uint32_t parity, int tag, int layer, long long* dbg) {
  if (mbar_try_wait_a(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_a(bar, parity)) {
    if (dbg != nullptr && *((volatile long long*)(dbg + 701)) > 40) return;   // the run is lost: let everybody fall through
    if (clock64() - t0 > 20000000LL) {
      if (dbg != nullptr) {
        const unsigned long long i = atomicAdd((unsigned long long*)(dbg + 700), 1ULL);
        if (i < 96) dbg[704 + i] = ((long long)tag << 40) | ((long long)(threadIdx.x >> 5) << 32) | (unsigned)layer;
        atomicAdd((unsigned long long*)(dbg + 701), 1ULL);
      }
      return;
    }
  }
}
#define T_WAIT(bar, par, tag, layer) t_wait_dbg(bar, par, tag, layer, P.dbg)
#else
#define T_WAIT(bar, par, tag, layer) mbar_wait_a(bar, par)
#endif

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
  s.total = s.bars + 8 * (2 * T_MAX_SLOTS + 5 + 2 * T_MAXKS) + 16;
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
  uint64_t* acc_bar = empty_bar + T_MAX_SLOTS;   // [2] all MMAs of N-part p of a layer are complete
  uint64_t* dr_bar = acc_bar + 2;                // [2] every epilogue warp has read part p of the accumulator out
  uint64_t* act_bar = dr_bar + 2;                // [MAXKS] k-step j of the next operand is in TMEM (4 arrivals: its owner warps)
  uint64_t* rel_bar = act_bar + T_MAXKS;         // [MAXKS] the layer's last MMAs that read k-step j of the current operand are complete
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(rel_bar + T_MAXKS);

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
    mbar_init(&acc_bar[0], 1);
    mbar_init(&acc_bar[1], 1);
    mbar_init(&dr_bar[0], T_EPI_WARPS);
    mbar_init(&dr_bar[1], T_EPI_WARPS);
    for (int j = 0; j < T_MAXKS; ++j) {
      mbar_init(&act_bar[j], 4);
      mbar_init(&rel_bar[j], 1);
    }
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
            for (int pt = 0; pt < 2; ++pt) {
              if (Y.ncol[pt] == 0) continue;
              const uint32_t kb = (uint32_t)Y.ncol[pt] * 64u;
              uint32_t off = Y.goff[pt];
              for (int j = 0; j < Y.nks; j += Y.kpg[pt]) {
                const uint32_t bytes = (uint32_t)min(Y.kpg[pt], Y.nks - j) * kb;
                if ((gc & 3u) == (uint32_t)lane) {
                  T_WAIT(empty_a + slot * 8, ph ^ 1, 30, (int)gc);
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
    }
    __syncwarp();
  } else if (warp == 1 || warp == 2) {
    // ================================================================== MMA issuers (one thread each)
    // A single thread issues dependent instructions ~5 cycles apart: with descriptor arithmetic, R2UR moves
    // and mbarrier probes one tcgen05.mma costs it 40-130 cycles, a split MMA only 40-64 tensor cycles.
    // Issuer 0 therefore issues N-part 0 of every layer (and the whole of a single-part layer), issuer 1
    // N-part 1, concurrently: both wait for the operand k-step by k-step (act_bar), accumulator columns of the
    // two parts are disjoint and every other ordering is carried by the mbarriers, so the two instruction
    // streams may interleave freely on the tensor pipe.
    const int which = warp - 1;
    if (elect_one()) {
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), ring_a = smem_u32(tsm + L.ring);
      const uint32_t acc_a = smem_u32(acc_bar) + 8 * which, dr_a = smem_u32(dr_bar), act_a = smem_u32(act_bar);
      const uint32_t rel_a = smem_u32(rel_bar);
      const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 0, 128) >> 32);
      const uint32_t d_t = tmem_base + T_D_COL, ah_t = tmem_base + T_AH_COL, al_t = tmem_base + T_AL_COL;
      const uint32_t slot16 = P.slot_bytes >> 4, ring16 = ring_a >> 4;
      uint32_t slot = 0, ph = 0, act_par = 0, dr_par = 0;
      int lno = -1, prev_n0 = 0x7fffffff;   // prev_n0: accumulator columns the previous layer's part 0 covers
      auto skip_slots = [&](int ng) {
        slot += (uint32_t)ng;
        while (slot >= (uint32_t)NS) { slot -= (uint32_t)NS; ph ^= 1; }
      };
      for (int ti = 0; ti < my_tiles; ++ti) {
        TPassWalk walk;
        for (;;) {
          const int kind = walk.next(P);
          if (kind == DIR_END) break;
          const TDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TLayer& Y = D.layer[l];
            const int nks = Y.nks, n0 = Y.ncol[0], n1 = Y.ncol[1];
            const int ng0 = (nks + Y.kpg[0] - 1) / Y.kpg[0], ng1 = n1 ? (nks + Y.kpg[1] - 1) / Y.kpg[1] : 0;
            const bool single = (n1 == 0);
            const uint32_t lay_mask = (1u << nks) - 1u;
            ++lno;
            T128_TRACE(lno, which * 40 + 0);
            if (which == 1) skip_slots(ng0);
            // Every epilogue warp has read the previous layer's accumulator out (both parts).  Besides the
            // write-after-read hazard on the accumulator columns this keeps every mbarrier at most ONE phase
            // ahead of its slowest waiter: a warp arrives here only after it has passed both acc_bar waits of the
            // previous layer, and nothing of this layer is committed before all sixteen have.
            // Issuer 0 needs part 1 of the previous layer drained only when it writes those columns (its part 0
            // reaches past the previous part 0) or commits acc_bar[1] itself (single-part layer); otherwise it
            // observes that phase after issuing, so part 0 starts while part 1 is still being read out.
            const bool dr1_first = which == 1 || single || n0 > prev_n0;
            prev_n0 = single ? 0x7fffffff : n0;
            T_WAIT(dr_a, dr_par, 1 + 10 * which, lno);
            if (dr1_first) T_WAIT(dr_a + 8, dr_par, 2 + 10 * which, lno);
            T128_TRACE(lno, which * 40 + 1);
            if (which == 1 && single) {
              dr_par ^= 1;
              // nothing to issue, but every phase of every barrier must be OBSERVED: a waiter that only flipped its
              // parity bits could run two phases ahead and then pass a parity wait on a stale phase
              for (int j = 0; j < nks; ++j) T_WAIT(act_a + 8 * j, (act_par >> j) & 1u, 15, lno * 16 + j);
              act_par ^= lay_mask;
              continue;
            }
            const int ncol = which ? n1 : n0, kpg = Y.kpg[which];
            const uint32_t idesc = t_idesc(ncol);
            const uint32_t tile16 = ((uint32_t)ncol * 32u) >> 4;          // one (hi or lo) tile, in 16-byte units
            const uint32_t lbo_f = (((uint32_t)ncol * 16u) >> 4) << 16;   // LBO field of the descriptor
            const uint32_t d_p = d_t + (which ? (uint32_t)n0 : 0u);
            const bool commit_rel = which == 1 || single;
            uint32_t a_off = 0;           // 8 j: TMEM column offset of k-step j of the operand, byte offset of its barriers
            uint32_t par = act_par;       // bit 0 = parity of the barrier of the k-step about to be issued
            int left = nks;
            while (left > 0) {
              const int nk = left < kpg ? left : kpg;
              left -= nk;
              T_WAIT(full_a + slot * 8, ph, 3 + 10 * which, lno);
              uint32_t b_lo = (ring16 + slot * slot16) | lbo_f;
#pragma unroll 1
              for (int jj = 0; jj < nk; ++jj) {
                T_WAIT(act_a + a_off, par & 1u, 4 + 10 * which, lno * 16 + (int)(a_off >> 3));
                par >>= 1;
                tc_fence_after();
                const uint64_t bh = ((uint64_t)desc_hi << 32) | b_lo;
                const uint64_t bl = ((uint64_t)desc_hi << 32) | (b_lo + tile16);
                t_mma(d_p, ah_t + a_off, bh, idesc, a_off);          // accumulate = (j > 0)
                t_mma(d_p, al_t + a_off, bh, idesc, 1u);
                t_mma(d_p, ah_t + a_off, bl, idesc, 1u);
                // the layer's last read of k-step j of the operand: the epilogue may overwrite it from here on
                if (commit_rel) umma_commit_a(rel_a + a_off);
                T128_TRACE(lno, which * 40 + 2 + (int)(a_off >> 3));
                a_off += 8;
                b_lo += 2 * tile16;
              }
              umma_commit_a(empty_a + slot * 8);
              if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
            }
            act_par ^= lay_mask;
            if (!dr1_first) T_WAIT(dr_a + 8, dr_par, 5, lno);
            dr_par ^= 1;
            umma_commit_a(acc_a);
            if (which == 0) {
              if (single) umma_commit_a(acc_a + 8);          // both accumulator phases complete together
              skip_slots(ng1);
            }
          }
        }
      }
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue / per-trajectory work
    const int ew = warp - T_FRONT;         // 0..15
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int sub = ew >> 2;               // which 4 of every 16 features; sub 0 also owns the trajectory's state,
                                           // sub 1 its action update
    const int r = q * 32 + lane;           // trajectory (TMEM lane) in the tile
    const int et = ew * 32 + lane;         // 0..511
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc_a = smem_u32(acc_bar), dr_a = smem_u32(dr_bar), act_a = smem_u32(act_bar), rel_a = smem_u32(rel_bar);
    uint32_t acc_ph = 0, rel_par = 0;
    int cur_kind = 0, elno = -1, elno2 = 0;
    const bool tr_on = (q == 0 && lane == 0);
    const int tr_base = 80 + sub * 12;
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

#ifdef GMPC_T128_TIMED
    long long e_acc = 0, e_hid = 0, e_bnd = 0, e_ld = 0, e_alu0 = 0, e_acc1 = 0, e_pub0 = 0, e_p1 = 0, eq, eh, ex;
    long long e_last = clock64(), e_lay[32];
    int e_prev = -1;
    for (int i = 0; i < 32; ++i) e_lay[i] = 0;
    auto lay_mark = [&](int id) {
      const long long now = clock64();
      if (e_prev >= 0) e_lay[e_prev] += now - e_last;
      e_last = now;
      e_prev = id;
    };
    const long long e_begin = clock64();
#define T128_E0(v) v = clock64()
#define T128_E1(acc, v) acc += clock64() - v
#else
#define T128_E0(v)
#define T128_E1(acc, v)
#endif
    // all MMAs of the layer (both N-parts) are complete
    auto wait_acc = [&]() {
      T128_E0(eq);
      T_WAIT(acc_a, acc_ph, 20, elno2);
      T_WAIT(acc_a + 8, acc_ph, 21, elno2);
      ++elno2;
      T128_E1(e_acc, eq);
#ifdef GMPC_T128_TIMED
      lay_mark(cur_kind * 8 + 7);
      ++elno;
      if (tr_on) T128_TRACE(elno, tr_base + 0);
#endif
      acc_ph ^= 1;
      tc_fence_after();
    };
    // "my reads of the accumulator (both parts) are complete" (after tcgen05.wait::ld)
    auto arrive_drained = [&]() {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive_a(dr_a); mbar_arrive_a(dr_a + 8); }
    };
    // sub 0 (the four warps that wrote it): "k-steps [0, nk) of the next operand are in TMEM" (after tcgen05.wait::st)
    auto arrive_act_range = [&](int nk) {
      if (sub != 0) return;
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
        cur_kind = kind;
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
          // k-step j of the next operand (features [16 j, 16 j + 16)) belongs to sub j % 4.  Part 0 (k-steps
          // [0, nk0)) is read out, processed AND stored while the tensor pipe works on part 1: k-step j of the
          // operand in flight may be overwritten as soon as part 1's MMAs of k-step j are complete (rel_bar[j]).
          const int nk0 = Y.ncol[0] >> 4, nk1 = Y.ncol[1] >> 4;
          const float inv = inv_s[Y.scale_idx];
          const float* bp = bias_s + Y.bias_off;
          float s = 1.f;
          // 16 accumulator values (one k-step of the next operand) of this thread's trajectory -> packed hi / lo halfs
          auto process16 = [&](const uint32_t (&raw)[16], int jg, uint32_t& mw, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
            float z[16];
            if (fwd) {
#pragma unroll
              for (int c4 = 0; c4 < 4; ++c4) {
                const float4 b4 = *reinterpret_cast<const float4*>(bp + 16 * jg + 4 * c4);
                z[4 * c4 + 0] = fmaf(__uint_as_float(raw[4 * c4 + 0]), inv, b4.x * s);
                z[4 * c4 + 1] = fmaf(__uint_as_float(raw[4 * c4 + 1]), inv, b4.y * s);
                z[4 * c4 + 2] = fmaf(__uint_as_float(raw[4 * c4 + 2]), inv, b4.z * s);
                z[4 * c4 + 3] = fmaf(__uint_as_float(raw[4 * c4 + 3]), inv, b4.w * s);
              }
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                mw = __funnelshift_l(__float_as_uint(z[c]), mw, 1);   // collects the SIGN bits (complemented below)
                z[c] = fmaxf(z[c], 0.f);
              }
            } else {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float v = __uint_as_float(raw[c]) * inv;
                z[c] = ((int)mw < 0) ? v : 0.f;
                mw <<= 1;
              }
            }
#pragma unroll
            for (int c = 0; c < 16; c += 4)
              opmax = fmaxf(opmax, fmaxf(fmaxf(fabsf(z[c]), fabsf(z[c + 1])), fmaxf(fabsf(z[c + 2]), fabsf(z[c + 3]))));
#pragma unroll
            for (int k = 0; k < 8; ++k) split_h2(z[2 * k], z[2 * k + 1], hi[k], lo[k]);
          };
          T128_E0(eh);
#pragma unroll 1
          for (int pt = 0; pt < 2; ++pt) {
            // part 0 runs under the MMAs of part 1, part 1 after the layer's last MMA
            T128_E0(eq);
            T_WAIT(acc_a + 8 * pt, acc_ph, 22 + pt, elno2);
            if (pt == 1) ++elno2;
            T128_E1(e_acc, eq);
#ifdef GMPC_T128_TIMED
            if (pt == 0) { lay_mark(kind * 8 + l); ++elno; }
            if (tr_on) T128_TRACE(elno, tr_base + pt * 6 + 0);
#endif
            tc_fence_after();
            if (pt == 0 && fwd) s = ssc_s[r];
            const int jbeg = pt ? nk0 : 0, jend = pt ? nk0 + nk1 : nk0, dcol = pt ? Y.ncol[0] : 0;
            const bool need_rel = (pt == 0 && nk1 > 0);
            uint32_t mw = fwd ? 0u : maskp[pt * T_EPI];
            int nb = 0;
            const int jfirst = jbeg + ((sub - jbeg) & 3);   // my first k-step of this part (k-step j belongs to sub j % 4)
            if (jfirst >= jend) {                           // none: nothing of this part to read
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_a(dr_a + 8 * pt);
            }
            uint32_t raw[16];
            if (jfirst < jend) {
              tmem_ld16_issue(tl + T_D_COL + dcol + 16 * (jfirst - jbeg), raw);
              tmem_ld_wait();
            }
#pragma unroll 1
            for (int j = jfirst; j < jend; j += 4) {
              uint32_t nxt[16], hi[8], lo[8];
              const bool more = j + 4 < jend;
              // the read of my next k-step is in flight under the arithmetic of this one (16 warps reading at once
              // are bound by the ~64 B/clk TMEM read path)
              if (more) tmem_ld16_issue(tl + T_D_COL + dcol + 16 * (j + 4 - jbeg), nxt);
              else {   // my last read of this part of the accumulator is complete
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive_a(dr_a + 8 * pt);
              }
              if (tr_on) T128_TRACE(elno, tr_base + pt * 6 + 1 + 2 * ((j - jfirst) >> 2));
              T128_E0(ex);
              process16(raw, j, mw, hi, lo);
              nb += 16;
              T128_E1(e_alu0, ex);
              T128_E0(ex);
              if (need_rel && j < Y.nks) T_WAIT(rel_a + 8 * j, (rel_par >> j) & 1u, 24, elno2 * 16 + j);
              T128_E1(e_acc1, ex);
              T128_E0(ex);
              t_st8(tl + T_AH_COL + 8 * j, hi);
              t_st8(tl + T_AL_COL + 8 * j, lo);
              t_st_wait();
              tc_fence_before();
              __syncwarp();
              if (lane == 0) mbar_arrive_a(act_a + 8 * j);
              T128_E1(e_pub0, ex);
              if (more) {
                tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) raw[c] = nxt[c];
              }
            }
            // sign bits -> (z >= +0) bits, first element at bit 31 (the adjoint sweep shifts them out in order)
            if (fwd) maskp[pt * T_EPI] = nb == 0 ? 0u : (~mw) << (32 - nb);
          }
          acc_ph ^= 1;
          rel_par ^= (1u << Y.nks) - 1u;
          T128_E1(e_p1, ex);
          T128_E1(e_hid, eh);
        }
        T128_E0(eh);
        // -------------------------------------------------------------- last layer of the pass (<= 32 outputs)
        const TLayer& Yf = D.layer[D.L - 1];
        const float invf = inv_s[Yf.scale_idx];
        rel_par ^= (1u << Yf.nks) - 1u;   // (the last layer's epilogues below start after all of its MMAs)
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
        T128_E1(e_bnd, eh);
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
#ifdef GMPC_T128_TIMED
    if (blockIdx.x == 0 && P.dbg != nullptr && lane == 0 && (ew == 0 || ew == 4)) {
      long long* o = P.dbg + 8 + (ew >> 2) * 8;   // sub 0 and sub 1 of one lane quarter
      o[0] = clock64() - e_begin; o[1] = e_acc; o[2] = e_hid; o[3] = e_bnd; o[4] = e_ld;
      if (ew == 0) for (int i = 0; i < 32; ++i) P.dbg[64 + i] = e_lay[i];
      P.dbg[24 + (ew >> 2) * 8 + 0] = e_alu0; P.dbg[24 + (ew >> 2) * 8 + 1] = e_acc1; P.dbg[24 + (ew >> 2) * 8 + 2] = e_pub0; P.dbg[24 + (ew >> 2) * 8 + 3] = e_p1;
    }
#endif
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
// Image of a layer: N-part 0 (operand rows [0, n0)), then N-part 1 (rows [n0, n0 + n1)); per part, per k-step:
// [hi tile | lo tile], tile = [2 k-chunks][rows of the part][8 halfs]
// (K-major SWIZZLE_NONE core matrices: LBO = rows * 16, SBO = 128; pinned by tools/t128_probe.cu).
__global__ void t128_pack_kernel(const float* __restrict__ W, int K, int N, int transposed, uint8_t* dst, int n0, int n1,
                                 int nks, const uint32_t* absmax, float* inv_scale) {
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
  int row = transposed ? k : o;
  const int kk = transposed ? o : k;
  __half hi, lo;
  split_h1(W[idx] * sc, hi, lo);
  const int part = row >= n0 ? 1 : 0;
  const int rows = part ? n1 : n0;
  if (part) { row -= n0; dst += (size_t)nks * n0 * 64; }
  const size_t tile_b = (size_t)rows * 32;
  const int j = kk >> 4, k16 = kk & 15;
  uint8_t* p = dst + (size_t)j * 2 * tile_b + (size_t)(k16 >> 3) * rows * 16 + (row >> 3) * 128 + (row & 7) * 16 + (k16 & 7) * 2;
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
  long long* d_dbg = nullptr;  // tools/t128_bench.cu with -DGMPC_T128_TIMED
};

using T128Kernel = void (*)(const TParams);
inline T128Kernel t128_kernel_ptr(int) { return plan_t128_kernel; }

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
  S.maxks = 13;
  // geometry: N-parts, biases, scales, images.  A layer with more than 32 output columns is issued as two
  // N-parts (k-steps of the next operand split ceil/floor), the narrow last layer of a pass as one.
  auto parts = [&](TLayer& Y) {
    const int nko = Y.npad / 16;
    if (Y.npad <= 32) { Y.ncol[0] = Y.npad; Y.ncol[1] = 0; }
    else { Y.ncol[0] = 16 * ((nko + 1) / 2); Y.ncol[1] = Y.npad - Y.ncol[0]; }
  };
  uint32_t slot = 0;
  auto widest = [&](const int* dims, int Ln) {
    for (int l = 0; l < Ln; ++l)
      for (int side = 0; side < 2; ++side) {
        TLayer Y;
        Y.npad = t_rup(dims[l + (side ? 0 : 1)], 16);
        parts(Y);
        slot = std::max(slot, (uint32_t)Y.ncol[0] * 64u);
      }
  };
  widest(S.dyn_dims, S.Ld);
  widest(S.cost_dims, S.Lc);
  S.slot_bytes = std::max(slot, 16384u);  // ring group: >= one k-step of the widest N-part, 16 KB when it is smaller
  int nbias = 0, nscale = 0;
  auto one_layer = [&](TLayer& Y, int M_true, int red, size_t& off) {
    Y.M_true = M_true; Y.npad = t_rup(M_true, 16); Y.nks = t_rup(red, 16) / 16;
    parts(Y);
    for (int pt = 0; pt < 2; ++pt) {
      Y.kpg[pt] = Y.ncol[pt] ? std::max(1, (int)(S.slot_bytes / (Y.ncol[pt] * 64))) : 1;
      Y.goff[pt] = (uint32_t)off;
      off += (size_t)Y.nks * Y.ncol[pt] * 64;
    }
    Y.pad_ = 0;
  };
  auto geom = [&](const int* dims, int Ln, TDir& F, TDir& Bw, size_t& off_f, size_t& off_b) {
    F.L = Bw.L = Ln; F.pad_ = Bw.pad_ = 0;
    for (int l = 0; l < Ln; ++l) {
      TLayer& Y = F.layer[l];
      one_layer(Y, dims[l + 1], dims[l], off_f);
      Y.bias_off = nbias; nbias += Y.npad;
      Y.scale_idx = nscale + l;
    }
    for (int i = 0; i < Ln; ++i) {
      const int lt = Ln - 1 - i;  // the adjoint pass visits the transposed layers L-1 .. 0
      TLayer& Y = Bw.layer[i];
      one_layer(Y, dims[lt], dims[lt + 1], off_b);
      Y.bias_off = 0; Y.scale_idx = nscale + lt;
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
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 0, const_cast<uint8_t*>(F.gsrc) + f.goff[0], f.ncol[0], f.ncol[1],
                                               f.nks, S.d_absmax + sbase + l, S.d_scale + sbase + l);
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 1, const_cast<uint8_t*>(Bw.gsrc) + rv.goff[0], rv.ncol[0],
                                               rv.ncol[1], rv.nks, S.d_absmax + sbase + l, nullptr);
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
  Q.dbg = S.d_dbg;
  if (Q.ntiles <= 0) return GMPC_OK;
  const int grid = std::min(Q.ntiles, S.num_sms);
  t128_kernel_ptr(S.maxks)<<<grid, T_THREADS, S.smem_bytes, st>>>(Q);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
