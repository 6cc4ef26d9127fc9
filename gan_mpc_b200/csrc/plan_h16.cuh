// plan_h16.cuh -- fused rollout + cost + adjoint + update kernel on the 5th-gen tensor cores,
// kind::f16 split-precision operands (see h16_common.cuh), software-pipelined layer chain.
//
// One persistent CTA per SM plans a tile of 32 trajectories at a time; the whole planning loop
// (N x {T forward steps, terminal-cost MLP, adjoint seed, T adjoint steps with the fused
// gradient/Adam update}, final evaluation) runs inside the kernel.  The contraction is issued as
//     out^T[features x 32 traj] = W^T[features x K] * act^T[K x 32 traj]
// with the WEIGHTS as the A operand (128 output features per block, <= 2 blocks) and the
// trajectories as the N dimension, so B = 4096 start states fill 128 SMs with full-height MMAs.
//
// What is different from the first, 3xTF32 formulation (tools/legacy/plan_tc.cuh), each justified by tools/mma_rate.cu on B200:
//  * an M=128 MMA with a small N is bound by the 4 KB read of A from shared memory (>= 40
//    cycles), not by the tensor pipe: fp16 operands cover K = 16 per MMA instead of 8, halving the
//    MMA count and the streamed weight bytes for the same 22-bit effective mantissa;
//  * the weight stream is block-major (all k-steps of output block 0, then block 1) with one
//    commit per block, TMEM accumulators and the hidden operand buffer are double-buffered, so the
//    epilogue of block 0 runs under the MMAs of block 1 and the next layer starts on the features
//    block 0 produced while block 1's epilogue is still running;
//  * the B operand is MN-major: an epilogue thread (one feature, 16 trajectories) stores four
//    16-byte vectors per block instead of 32 scalar stores;
//  * the step boundary is handled by the two warps that can read the <= 32 valid accumulator
//    lanes: they add the residual / adjoint terms and write the next step's operand directly.
//
// Restates the same reference lines as plan_ffma.cuh (dynamics/nn.py:27-34, cost/nn.py:23-29,
// cost/cost_model.py:20-42, policy/optimizers.py:24-31 and :78-83; optax adam of norm/runner.py:53).
#pragma once
#include <cuda_runtime.h>

#include <stdio.h>
#include <stdlib.h>

#include <string>
#include <type_traits>
#include <vector>

#include "../../include/gmpc.h"
#include "common.cuh"
#include "h16_common.cuh"

namespace gmpc {

inline int rup(int v, int a) { return (v + a - 1) / a * a; }

constexpr int H_THREADS = 384;  // producer, issuer, 8 compute warps, second producer, second issuer
constexpr int H_PRODUCER2 = 10; // warp index of the second producer
constexpr int H_ISSUER2 = 11;   // warp index of the second MMA issuer (Wl x ah products)
constexpr int H_TMEM_BLK = 96;  // accumulator columns per block: D1 | D2 | D3 (32 each)
constexpr int H_TMEM_BUF = 2 * H_TMEM_BLK;  // per layer buffer (two blocks); two buffers ping-pong
constexpr int H_COMPUTE = 256;
constexpr int H_BK_BYTES = 8192;     // one block-k-step: hi unit + lo unit
constexpr int H_GROUP_BYTES = 2 * H_BK_BYTES;  // ring slot = one bulk copy = two k-steps of one block
constexpr int H_SLOT_BYTES = H_GROUP_BYTES;
constexpr int H_MAX_SLOTS = 12;
constexpr int H_SROW = H_NB + 1;     // padded row of the small fp32 arrays
constexpr int H_SB_BYTES = 4096;     // small operand: <= 32 features

struct HLayer {
  const uint8_t* gsrc;     // [block][k-step][hi unit | lo unit]
  const float* bias;       // forward only
  const float* inv_scale;  // device scalar: 1 / (power-of-two weight scale of this layer)
  int M_true;              // output features of this (possibly transposed) layer
  int nblk;                // 128-row blocks
  int ksteps;              // reduction length / 16 (may be odd: the last group of a block is then one k-step)
  int next_kpad;           // round_up(M_true, 16): features the epilogue defines in the next operand
  int rows[4];             // stored A rows of block b (multiple of 16, >= 32, <= 128): the MMA reads
                           // 128 rows, the rows past rows[b] alias later bytes and land in unused lanes
};
struct HDir {
  HLayer layer[MAXL];
  int L;
  int pad_;
};

struct HParams {
  HDir dir[4];
  int n, m, T, K;
  int fout, mode, method, iters, use_cost, final_fwd, ntiles, nslot;
  int wide, pad1_;  // wide: hidden width in (256, 512] -> 3-4 row blocks, serial layer schedule (see kernel)
  uint32_t hb_bytes, exp_;  // exp_: timing experiments (GMPC_H16_EXP with GMPC_DEBUG: 1 = no weight stream,
                            // 2 = no operand stores); honoured by the TIMED instantiation only, results are garbage
  const uint2* gtab[4];     // per pass: {byte offset in the pass image, bytes} of every ring group
  uint32_t ngroups[4];
  long long NQ;
  float lr, b1, b2, eps;
  float acc_comp;   // relative gain per accumulate event that undoes the truncation of the fp32 accumulation (see H16State)
  int pad2_;
  const float *x0, *U_in, *goal, *mpcw;
  float *U_out, *X_out, *J_out, *dU_out, *lam_out;
  float *ws_X, *ws_G, *ws_U, *ws_M, *ws_V, *ws_S;
  uint32_t* ws_mask;
  uint32_t* ovf;   // incremented by a CTA that clamped an fp16 operand (|scaled value| > 65000)
  long long* dbg;
};

__device__ __forceinline__ int h_pass_kind(const HParams& P, int p) {
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

// The same schedule as h_pass_kind, walked incrementally (no divisions in the issuers' path).
struct HPassWalk {
  int pp = 0, itc = 0;
  __device__ __forceinline__ int next(const HParams& P) {
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

// Shared-memory carve-up (byte offsets from the 128-aligned dynamic base).
struct HSmem {
  uint32_t ring, hb0, hb1, sb, small, gtab, bars, total;
  // rows of the small fp32 arrays, in units of H_SROW floats
  int r_x, r_lam, r_dq, r_y, r_st, st_rows, r_part, r_sc, rows;
  // offsets inside one staging buffer (two buffers ping-pong, filled one step ahead by cp.async)
  int o_pu, o_pg, o_px, o_pm, o_pv, o_su, o_sd;
};
__host__ __device__ inline HSmem h_smem_layout(int nslot, uint32_t hb_bytes, int n, int m, int fout, int ngroups_total,
                                               int wide) {
  HSmem s;
  s.ring = 0;
  s.hb0 = (uint32_t)nslot * H_SLOT_BYTES;
  s.hb1 = wide ? s.hb0 : s.hb0 + hb_bytes;  // wide: one buffer, rewritten in place between layers
  s.sb = s.hb1 + hb_bytes;
  s.small = s.sb + H_SB_BYTES;
  int r = 0;
  s.r_x = r; r += n;
  s.r_lam = r; r += n;
  s.r_dq = r; r += 2 * (n + m);  // two generations: the update runs one step behind the boundary
  s.r_y = r; r += fout;
  int o = 0;
  s.o_pu = o; o += m;   // U[t]
  s.o_pg = o; o += n;   // goal[t]
  s.o_px = o; o += n;   // X[t]
  s.o_pm = o; o += m;   // Adam first moment [t]
  s.o_pv = o; o += m;   // Adam second moment [t]
  s.o_su = o; o += 1;   // sqrt(|u_t|^2 + a^2)   (saved by the forward sweep)
  s.o_sd = o; o += 1;   // sqrt(|x_t - goal_t|^2 + a^2)
  s.st_rows = o;
  s.r_st = r; r += 3 * o;   // three staging buffers (t % 3)
  s.r_part = r; r += 32;  // per-warp partial sums of the two staging-cost norms, two generations
  s.r_sc = r; r += 5;     // 1/scale of the adjoint operand in flight, scale for the next one,
                          // two generations (step parity) of the forward operand scale
  s.rows = r;
  s.gtab = s.small + (uint32_t)r * H_SROW * 4;
  s.gtab = (s.gtab + 15u) & ~15u;
  s.bars = s.gtab + (uint32_t)ngroups_total * 8;
  s.bars = (s.bars + 15u) & ~15u;
  s.total = s.bars + 256;
  return s;
}

// FSCALE: per-trajectory power-of-two scaling of the FORWARD operands as well (states of any
// magnitude); without it the forward pass assumes |state|, |action|, |activation| < 65000 and
// counts clamps (HParams::ovf).  The adjoint sweep is always rescaled.
template <bool TIMED, bool WIDE, bool FSCALE>
__global__ void __launch_bounds__(H_THREADS, 1) plan_h16_kernel(const __grid_constant__ HParams P) {
  extern __shared__ __align__(128) uint8_t hsm[];
  const HSmem L = h_smem_layout(P.nslot, P.hb_bytes, P.n, P.m, P.fout,
                                (int)(P.ngroups[0] + P.ngroups[1] + P.ngroups[2] + P.ngroups[3]), WIDE);
  uint8_t* ring = hsm + L.ring;
  uint8_t* HB0 = hsm + L.hb0;
  uint8_t* HB1 = hsm + L.hb1;
  uint8_t* SB = hsm + L.sb;
  float* small = reinterpret_cast<float*>(hsm + L.small);
  float* x_s = small + L.r_x * H_SROW;
  float* lam_s = small + L.r_lam * H_SROW;
  float* dq_s = small + L.r_dq * H_SROW;
  float* y_s = small + L.r_y * H_SROW;
  float* st_s = small + L.r_st * H_SROW;      // two staging buffers of L.st_rows rows
  float* part_s = small + L.r_part * H_SROW;  // [8] partial |x-goal|^2, then [8] partial |u|^2
  float* isc_s = small + L.r_sc * H_SROW;     // 1 / scale of the adjoint operand in flight
  float* snx_s = isc_s + H_SROW;              // scale for the next adjoint operand
  // [2][32] forward operand scale of step t in row (t & 1), 16-byte aligned for vector loads
  float* fsc_s = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(snx_s + H_SROW) + 15) & ~(uintptr_t)15);
  constexpr int FSR = H_NB;                   // row stride of fsc_s
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(hsm + L.bars);
  uint64_t* empty_bar = full_bar + H_MAX_SLOTS;
  uint64_t* acc_bar = empty_bar + H_MAX_SLOTS;  // [2]: accumulator block b complete
  uint64_t* act_bar = acc_bar + 2;              // [2]: operand part b (features [128b, 128b+128)) ready
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(act_bar + 2);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n = P.n, m = P.m, T = P.T, NS = P.nslot;
  const uint32_t C = cluster_nctarank(), crank = cluster_ctarank();
  const uint16_t cmask = (uint16_t)((1u << C) - 1u);
  const int n_iter = (P.ntiles + (int)gridDim.x - 1) / (int)gridDim.x;  // uniform per cluster

  // zero all operand / scratch memory once: padded features must stay finite
  for (uint32_t i = tid * 4; i < L.gtab; i += H_THREADS * 4) *reinterpret_cast<uint32_t*>(hsm + i) = 0u;
  {  // group tables of the four passes -> shared memory (the producers index them every copy)
    uint2* gt = reinterpret_cast<uint2*>(hsm + L.gtab);
    uint32_t o = 0;
    for (int k = 0; k < 4; ++k) {
      for (uint32_t i = tid; i < P.ngroups[k]; i += H_THREADS) gt[o + i] = P.gtab[k][i];
      o += P.ngroups[k];
    }
  }
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 2 * C);  // one tcgen05.commit arrival per issuer from every CTA of the cluster
    }
    mbar_init(&acc_bar[0], 2);  // both issuers commit
    mbar_init(&acc_bar[1], 2);
    mbar_init(&act_bar[0], H_COMPUTE / 32);  // one arrival per compute warp
    mbar_init(&act_bar[1], H_COMPUTE / 32);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_holder, 512);  // 2 buffers x 2 blocks x 96 fp32 columns = 384
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();  // peers' barriers exist before anything is multicast to them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == 0 || warp == H_PRODUCER2) {
    // ================================================================== weight-stream producers
    // tools/bulk_copy_rate.cu: one cp.async.bulk blocks its issuing thread ~460 cycles whatever the
    // size, but copies issued by several lanes of one instruction cost only ~67 cycles each.  The
    // stream is therefore cut into 16 KB groups (two k-steps of one block, hi+lo units), and two
    // producer warps x LP lanes issue R = 2 LP group copies per round: lane k of producer w owns
    // the groups G = R r + LP w + k.  Every pass (all layers of one direction) is one contiguous
    // image.  R <= ring slots is required: a lane waits on a slot by phase PARITY only, which is
    // sound only if it can never be two refills ahead of the slot; its warp's previous round
    // guarantees the releases up to G - R - NS, and G - R - NS >= G - 2 NS iff NS >= R.
    const int w = (warp == 0) ? 0 : 1;
    const int LP = NS >= 8 ? 4 : (NS >= 6 ? 3 : 2), R = 2 * LP;
    if (lane < LP && !(TIMED && (P.exp_ & 1))) {
      uint32_t ngk[4];
      const uint8_t* basek[4];
      const uint2* gtk[4];
      {
        const uint2* gt = reinterpret_cast<const uint2*>(hsm + L.gtab);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          ngk[k] = P.ngroups[k];
          basek[k] = P.dir[k].layer[0].gsrc;
          gtk[k] = gt;
          gt += P.ngroups[k];
        }
      }
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), ring_a = smem_u32(ring);
      const uint32_t G0 = (uint32_t)(2 * lane + w);  // consecutive groups alternate between the two producer warps
      uint32_t slot = G0 % (uint32_t)NS, ph = (G0 / (uint32_t)NS) & 1u;
      uint32_t gi = G0;  // group index relative to the start of pass p (may run past its end)
      int ti = 0, p = 0, kind = h_pass_kind(P, 0);
      const long long t_start = clock64();
      while (true) {
        while (kind != DIR_END && gi >= ngk[kind]) {
          gi -= ngk[kind];
          kind = h_pass_kind(P, ++p);
        }
        if (kind == DIR_END) {
          if (++ti == n_iter) break;
          p = 0;
          kind = h_pass_kind(P, 0);
          continue;
        }
        const uint32_t bar = full_a + slot * 8;
        // poll instead of blocking: a lane whose slot is free issues its copy at once, it is not held
        // back by the slower slots of the other lanes of its warp
        if (!mbar_try_wait_a(empty_a + slot * 8, ph ^ 1)) {  // all C CTAs released the slot?
          if (clock64() - t_start > 60000000000LL) __trap();
          continue;
        }
        const uint2 ge = gtk[kind][gi];         // {offset, bytes}; bytes is a multiple of 64 * C
        const uint32_t part = ge.y / C;
        const uint32_t dst = ring_a + slot * H_GROUP_BYTES + crank * part;
        const uint8_t* src = basek[kind] + ge.x + crank * part;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(ge.y) : "memory");
        if (C > 1)
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster "
              "[%0], [%1], %2, [%3], %4;" ::"r"(dst), "l"(src), "r"(part), "r"(bar), "h"(cmask) : "memory");
        else
          asm volatile(
              "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
              ::"r"(dst), "l"(src), "r"(part), "r"(bar) : "memory");
        gi += R;
        slot += R;
        while (slot >= (uint32_t)NS) { slot -= (uint32_t)NS; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1 || warp == H_ISSUER2) {
    // ================================================================== MMA issuers
    // ONE thread per issuer runs the whole role, and a single instruction stream is the limit
    // (ncu: ~25 dependent SASS instructions per block-k-step even when stripped): the two products
    // of a k-step therefore go out from two warps -- issuer 0: [D1|D2] (+)= Wh x [ah|al] (N = 64),
    // issuer 1: D3 (+)= Wl x ah (N = 32) into its own accumulator columns, so the result does not
    // depend on how the two streams interleave.  Work is issued in fully unrolled chunks of two
    // ring groups (four k-steps): full barriers probed back to back, descriptors are adds of
    // running 32-bit words, the split at k-step 8 (operand part 1) is hoisted out of the loop, the
    // pass schedule is walked without divisions, timers exist only in the TIMED instantiation.
    const int which = (warp == 1) ? 0 : 1;
    // GMPC_H16_WARP_ISSUER (experiment): the whole warp walks the schedule and computes the descriptors
    // (warp-uniform values can then live in uniform registers, no R2UR in front of every MMA); only the
    // elected lane issues.  Default: the elected lane runs the role alone.
    const bool leader = elect_one();
#ifdef GMPC_H16_WARP_ISSUER
    const bool run_role = true;
#define H16_LEAD(stmt) do { if (leader) { stmt; } } while (0)
#else
    const bool run_role = leader;
#define H16_LEAD(stmt) do { stmt; } while (0)
#endif
    if (run_role) {
      uint32_t act_ph0 = 0, act_ph1 = 0, lc = 0;
      long long t_act = 0, t_full = 0, tt = 0, t_actl[8] = {0, 0, 0, 0, 0, 0, 0, 0}, t_iss1 = 0, n_iss1 = 0, tl0 = 0, t_b0a = 0, t_b0b = 0, t_p1 = 0, t_b1 = 0, tl1 = 0;
      const uint32_t idesc = which == 0 ? h16_idesc(2 * H_NB, 0, 1) : h16_idesc(H_NB, 0, 1);
      const uint32_t hb_a[2] = {smem_u32(HB0), smem_u32(HB1)}, sb_a = smem_u32(SB);
      const uint64_t a_desc0 = umma_smem_desc(smem_u32(ring), 0, H_A_SBO);
      const uint32_t a_hi = (uint32_t)(a_desc0 >> 32), a_lo0 = (uint32_t)a_desc0;
      const uint32_t full_a = smem_u32(full_bar);
      constexpr uint32_t EMPTY_OFF = H_MAX_SLOTS * 8;  // empty_bar[s] sits EMPTY_OFF bytes after full_bar[s]
      constexpr uint32_t KS = H_B_KSTEP >> 4;          // B descriptor advance per k-step
      const uint32_t acc_a = smem_u32(acc_bar), act_a = smem_u32(act_bar);
      const uint32_t d_off = which == 0 ? 0u : 2u * H_NB;  // issuer 1 accumulates into columns [64, 96)
      // ring cursor: barrier address, A descriptor low word of the slot start, slots left, parity
      uint32_t fb = full_a, a_lo = a_lo0, left = (uint32_t)NS, ph = 0;
      const bool nostream = TIMED && (P.exp_ & 1);  // timing experiment: no weight stream (garbage results)
      // loop invariants the compiler would otherwise re-derive from special registers / the
      // constant bank inside every chunk (S2UR SR_CgaSize -> UIMAD -> LDCU chains in the SASS)
      uint32_t NSr = (uint32_t)NS, mc = C > 1 ? 1u : 0u, cmask_r = cmask, full_r = full_a, alo0_r = a_lo0;
#ifndef GMPC_H16_WARP_ISSUER
      asm volatile("" : "+r"(NSr), "+r"(mc), "+r"(cmask_r), "+r"(full_r), "+r"(alo0_r));
#endif
      // per-block A geometry (set by the layer loop): low-word addend that selects this issuer's unit
      // and carries the LBO field, and the descriptor advance from the first to the second k-step
      uint32_t a_blk = 0, a_k2 = 0;
      // issue U consecutive groups of one block, fully unrolled; the last one holds NKL k-steps
      auto chunk = [&](auto Utag, auto NKLtag, uint32_t d0, uint32_t b_lo, uint32_t b_hi, uint32_t acc) {
        constexpr int U = decltype(Utag)::value;
        constexpr int NKL = decltype(NKLtag)::value;
        uint32_t fbu[U], au[U], phu[U];
        bool oku[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          fbu[u] = fb; au[u] = a_lo; phu[u] = ph;
          fb += 8;
          a_lo += (H_GROUP_BYTES >> 4);
          if (--left == 0) { fb = full_r; a_lo = alo0_r; left = NSr; ph ^= 1; }
        }
        if (!nostream) {
#pragma unroll
          for (int u = 0; u < U; ++u) oku[u] = mbar_try_wait_a(fbu[u], phu[u]);
#pragma unroll
          for (int u = 0; u < U; ++u)
            if (!oku[u]) {
              if (TIMED) tt = clock64();
              mbar_wait_a(fbu[u], phu[u]);
              if (TIMED) t_full += clock64() - tt;
            }
        }
        tc_fence_after();
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint64_t ad = ((uint64_t)a_hi << 32) | (au[u] + a_blk);
          const uint64_t bd = ((uint64_t)b_hi << 32) | (b_lo + u * 2 * KS);
          H16_LEAD(umma_f16(d0, ad, bd, idesc, (u == 0) ? acc : 1u));
          if (u + 1 < U || NKL == 2) H16_LEAD(umma_f16(d0, ad + a_k2, bd + KS, idesc, 1u));  // second k-step of the group
          if (mc)
            H16_LEAD(umma_commit_mc_a(fbu[u] + EMPTY_OFF, (uint16_t)cmask_r));
          else
            H16_LEAD(umma_commit_a(fbu[u] + EMPTY_OFF));
        }
      };
      using I1 = std::integral_constant<int, 1>;
      using I2 = std::integral_constant<int, 2>;
      // issue groups [g0, g1) of one block; `half_tail`: the last of them holds a single k-step
      auto issue = [&](int g0, int g1, bool half_tail, uint32_t d0, uint32_t b_lo, uint32_t b_hi, uint32_t acc_first) {
        uint32_t acc = acc_first;
        int g = g0;
        const int gfull = half_tail ? g1 - 1 : g1;
        for (; g + 2 <= gfull; g += 2, b_lo += 4 * KS, acc = 1u) chunk(I2{}, I2{}, d0, b_lo, b_hi, acc);
        if (g < gfull) { chunk(I1{}, I2{}, d0, b_lo, b_hi, acc); b_lo += 2 * KS; acc = 1u; }
        if (half_tail) chunk(I1{}, I1{}, d0, b_lo, b_hi, acc);
      };
      for (int ti = 0; ti < n_iter; ++ti) {
        const bool live = (int)blockIdx.x + ti * (int)gridDim.x < P.ntiles;
        HPassWalk walk;
        for (;;) {
          const int kind = walk.next(P);
          if (kind == DIR_END) break;
          const HDir& D = P.dir[kind];
          int prev_nblk = 1;
          for (int l = 0; l < D.L; ++l, ++lc) {
            const HLayer& Y = D.layer[l];
            const int ngrp = (Y.ksteps + 1) >> 1, nblk = Y.nblk;  // groups per block
            const bool odd = Y.ksteps & 1;
            if (!live) {  // no tile this round: keep the cluster's ring protocol going
              for (int i = 0; i < nblk * ngrp; ++i) {
                mbar_wait_a(fb, ph);
                H16_LEAD(umma_commit_mc_a(fb + EMPTY_OFF, cmask));
                fb += 8;
                a_lo += (H_GROUP_BYTES >> 4);
                if (--left == 0) { fb = full_a; a_lo = a_lo0; left = (uint32_t)NS; ph ^= 1; }
              }
              continue;
            }
            const bool two_parts = !WIDE && (l > 0) && (prev_nblk > 1);
            prev_nblk = nblk;
            const uint64_t b_desc0 = umma_smem_desc((l == 0) ? sb_a : hb_a[(l - 1) & 1], H_B_LBO, H_B_SBO);
            const uint32_t b_hi = (uint32_t)(b_desc0 >> 32), b_lo0 = (uint32_t)b_desc0;
            const uint32_t d_base = tmem_base + (WIDE ? 0u : (lc & 1) * H_TMEM_BUF) + d_off;
            if (TIMED) tt = clock64();
            mbar_wait_a(act_a, act_ph0);
            if (TIMED) { const long long dt = clock64() - tt; t_act += dt; if (kind == DIR_DYN_F || kind == DIR_DYN_B) t_actl[l & 3] += dt; }
            act_ph0 ^= 1;
            if (TIMED) tl0 = clock64();
            // block geometry: unit = rows x 32 B; this issuer's unit (hi or lo), LBO = rows x 16 B
            auto set_block = [&](int rows) {
              a_blk = (uint32_t)(which * rows * 32) >> 4 | ((uint32_t)(rows * 16) >> 4) << 16;
              a_k2 = (uint32_t)(rows * 64) >> 4;
            };
            if (WIDE) {
              // serial schedule: the operand is complete, all blocks go out back to back into the
              // single accumulator buffer, one commit for the layer
              for (int b = 0; b < nblk; ++b) {
                set_block(Y.rows[b]);
                issue(0, ngrp, odd, d_base + b * H_TMEM_BLK, b_lo0, b_hi, 0u);
              }
              H16_LEAD(umma_commit_a(acc_a));
              continue;
            }
            set_block(Y.rows[0]);
            // block 0: groups [0, 4) (k-steps 0..7) need operand part 0 only; the rest need part 1
            const bool probe = TIMED && kind == DIR_DYN_F && l == 2;
            if (two_parts) {
              issue(0, 4, false, d_base, b_lo0, b_hi, 0u);
              if (probe) { tl1 = clock64(); t_b0a += tl1 - tl0; }
              if (TIMED) tt = clock64();
              mbar_wait_a(act_a + 8, act_ph1);
              if (TIMED) { const long long dt = clock64() - tt; t_act += dt; if (kind == DIR_DYN_F || kind == DIR_DYN_B) t_actl[4 + (l & 3)] += dt; }
              act_ph1 ^= 1;
              if (probe) { const long long t1 = clock64(); t_p1 += t1 - tl1; tl1 = t1; }
              issue(4, ngrp, odd, d_base, b_lo0 + 8 * KS, b_hi, 1u);
              if (probe) { const long long t1 = clock64(); t_b0b += t1 - tl1; tl1 = t1; }
            } else {
              issue(0, ngrp, odd, d_base, b_lo0, b_hi, 0u);
            }
            H16_LEAD(umma_commit_a(acc_a));
            if (nblk > 1) {
              set_block(Y.rows[1]);
              issue(0, ngrp, odd, d_base + H_TMEM_BLK, b_lo0, b_hi, 0u);
              H16_LEAD(umma_commit_a(acc_a + 8));
            }
            if (probe) { const long long t1 = clock64(); t_b1 += t1 - tl1; t_iss1 += t1 - tl0; ++n_iss1; }
          }
        }
      }
      if (TIMED && leader) {
        P.dbg[blockIdx.x * 16 + 0 + 9 * which] = t_act;
        P.dbg[blockIdx.x * 16 + 1 + 9 * which] = t_full;
        if (blockIdx.x == 0)
          printf("[gmpc] h16 issuer %d: dyn-fwd layer 2 issue spans: k0-7 %lld | wait part1 %lld | k8-12 %lld | block1 %lld | total %lld (avg cycles over %lld)\n",
                 which, t_b0a / max(n_iss1, 1LL), t_p1 / max(n_iss1, 1LL), t_b0b / max(n_iss1, 1LL), t_b1 / max(n_iss1, 1LL), t_iss1 / max(n_iss1, 1LL), n_iss1);
        if (which == 0 && blockIdx.x == 0)
          printf("[gmpc] h16 issuer wait_act by dyn layer (part0 | part1): %lld %lld %lld %lld | %lld %lld %lld %lld\n",
                 t_actl[0], t_actl[1], t_actl[2], t_actl[3], t_actl[4], t_actl[5], t_actl[6], t_actl[7]);
      }
    }
#undef H16_LEAD
    __syncwarp();
  } else {
    // ================================================================== epilogue / elementwise
    const int ct = tid - 64;               // 0..255
    const int q = warp & 3;                // TMEM lane quadrant this warp may access
    const int half = (warp - 2) >> 2;      // which 16 of the 32 trajectory columns
    const int c0 = half * (H_NB / 2);
    const bool fwarp = (q == 0);           // owns accumulator lanes 0..31: the <= 32-feature outputs
    const uint32_t t_lane = (uint32_t)(q * 32) << 16;
    uint32_t acc_ph0 = 0, acc_ph1 = 0, lc = 0;
    long long t_acc = 0, t_epi = 0, t_fin = 0, t_bnd = 0, t_total0 = clock64(), tq = 0;
    long long t_o1 = 0, t_o2 = 0, t_o3 = 0, t_o4 = 0, t_b0done = 0;
    constexpr bool timed = TIMED;
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
    float* wsX = P.ws_X + (size_t)blockIdx.x * (T + 1) * n * H_NB;
    float* wsG = P.ws_G + (size_t)blockIdx.x * (T + 1) * n * H_NB;
    float* wsU = P.ws_U + (size_t)blockIdx.x * T * m * H_NB;
    float* wsM = P.ws_M + (size_t)blockIdx.x * T * m * H_NB;
    float* wsV = P.ws_V + (size_t)blockIdx.x * T * m * H_NB;
    float* wsS = P.ws_S + (size_t)blockIdx.x * T * 2 * H_NB;  // saved staging-cost norms
    const int MSTR = (WIDE ? 2 : 1) * H_COMPUTE;  // mask words per (step, layer): 32 bits cover 2 blocks x 16 columns
    uint32_t* wsMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)T * (Ld - 1) + (Lc - 1)) * MSTR;
    uint32_t* costMask = wsMask + (size_t)T * (Ld - 1) * MSTR;
    const bool adam = (P.mode == MODE_PLAN && P.method == 1);
    const bool need_goal = cost_mode || P.mode == MODE_L2GRAD;

    // "operand part `part` of the next layer is in shared memory": every writer fences its own
    // generic-proxy stores for the async proxy, then one lane per warp arrives (count 8)
    const uint32_t acc_sa = smem_u32(acc_bar), act_sa = smem_u32(act_bar);
    auto publish = [&](int part) {
      fence_proxy_async_smem();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(act_sa + part * 8);
    };
    // 16 values of feature f (trajectories c0 .. c0+15), already scaled -> operand buffer `dst`
    float opmax = 0.f;  // largest operand magnitude this thread has written (range check)
    auto store_row16 = [&](uint8_t* dst, int f, const float (&v)[16]) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int k = 0; k < 16; k += 2) opmax = fmaxf(opmax, fmaxf(fabsf(v[k]), fabsf(v[k + 1])));
#pragma unroll
      for (int k = 0; k < 8; ++k) split_h2(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
      uint8_t* p = dst + (f >> 3) * H_B_LBO + (f & 7) * 16 + (c0 >> 3) * H_B_SBO;
      *reinterpret_cast<uint4*>(p) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
      *reinterpret_cast<uint4*>(p + H_B_SBO) = make_uint4(hi[4], hi[5], hi[6], hi[7]);
      *reinterpret_cast<uint4*>(p + 4 * H_B_SBO) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
      *reinterpret_cast<uint4*>(p + 5 * H_B_SBO) = make_uint4(lo[4], lo[5], lo[6], lo[7]);
    };
    // hidden layer li of a pass: TMEM -> (+bias, relu, mask) or (mask gate) -> hi/lo -> HB[li & 1]
    const int f0 = q * 32 + lane;  // this thread's feature in row block 0 (block 1: +128)
    bool probe_layer = false;
    // wide layers (3-4 row blocks): serial schedule.  All MMAs of the layer are complete before the
    // single operand buffer is rewritten in place; one publish when every block has been written.
    auto hidden_epilogue_wide = [&](const HLayer& Y, bool fwd, uint32_t* maskp, const float* fsc) {
      const uint32_t d_base = tmem_base + t_lane + c0;
      const float inv = *Y.inv_scale * fmaf(P.acc_comp, (float)Y.ksteps, 1.f);   // (un-biasing of the truncating accumulation)
      if (timed) tq = clock64();
      mbar_wait_a(acc_sa, acc_ph0);
      acc_ph0 ^= 1;
      if (timed) { const long long t1 = clock64(); t_acc += t1 - tq; tq = t1; }
      tc_fence_after();
      float sc[16];
#pragma unroll
      for (int c4 = 0; c4 < 4; ++c4) {
        const float4 q4 = (FSCALE && fwd) ? *reinterpret_cast<const float4*>(fsc + c0 + 4 * c4) : make_float4(1.f, 1.f, 1.f, 1.f);
        sc[4 * c4] = q4.x; sc[4 * c4 + 1] = q4.y; sc[4 * c4 + 2] = q4.z; sc[4 * c4 + 3] = q4.w;
      }
      for (int b = 0; b < Y.nblk; ++b) {
        const int f = b * 128 + f0;
        uint32_t* mp = maskp + (b >> 1) * H_COMPUTE + ct;
        uint32_t mw = (fwd && !(b & 1)) ? 0u : *mp;
        const float bias = (fwd && f < Y.M_true) ? Y.bias[f] : 0.f;
        uint32_t d1[16], d2[16], d3[16];
        tmem_ld16_issue(d_base + b * H_TMEM_BLK, d1);
        tmem_ld16_issue(d_base + b * H_TMEM_BLK + H_NB, d2);
        tmem_ld16_issue(d_base + b * H_TMEM_BLK + 2 * H_NB, d3);
        tmem_ld_wait();
        if (f < Y.next_kpad) {
          const bool live = f < Y.M_true;
          float v[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) {
            float z = live ? fmaf(__uint_as_float(d1[c]) + (__uint_as_float(d2[c]) + __uint_as_float(d3[c])), inv, bias * sc[c]) : 0.f;
            if (fwd) {
              if (z > 0.f) mw |= 1u << ((b & 1) * 16 + c);
              z = fmaxf(z, 0.f);
            } else {
              z = ((mw >> ((b & 1) * 16 + c)) & 1u) ? z : 0.f;
            }
            v[c] = z;
          }
          store_row16(HB0, f, v);
        }
        if (fwd) *mp = mw;
      }
      publish(0);
      if (timed) t_epi += clock64() - tq;
      ++lc;
    };
    // `fsc` (forward only): the 32 per-trajectory scales of the operand in flight.  The forward pass
    // runs in scaled units a' = a * s (ReLU is positively homogeneous), so the bias enters as b * s.
    auto hidden_epilogue = [&](const HLayer& Y, int li, bool fwd, uint32_t* maskp, const float* fsc) {
      if (WIDE) { hidden_epilogue_wide(Y, fwd, maskp, fsc); return; }
      uint8_t* dst = (li & 1) ? HB1 : HB0;
      const uint32_t d_base = tmem_base + (lc & 1) * H_TMEM_BUF + t_lane + c0;
      uint32_t mw = fwd ? 0u : maskp[ct];
      const float inv = *Y.inv_scale * fmaf(P.acc_comp, (float)Y.ksteps, 1.f);   // (un-biasing of the truncating accumulation)
      float bias[2] = {0.f, 0.f};
      if (fwd) {
        if (f0 < Y.M_true) bias[0] = Y.bias[f0];
        if (f0 + 128 < Y.M_true) bias[1] = Y.bias[f0 + 128];
      }
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        if (b < Y.nblk) {
          if (timed) tq = clock64();
          if (b == 0) { mbar_wait_a(acc_sa, acc_ph0); acc_ph0 ^= 1; }
          else        { mbar_wait_a(acc_sa + 8, acc_ph1); acc_ph1 ^= 1; }
          if (timed) {
            const long long t1 = clock64();
            t_acc += t1 - tq;
            tq = t1;
            if (probe_layer) { if (b == 0) t_b0done = t1; else { t_o1 += t1 - t_b0done; ++t_o2; } }
          }
          tc_fence_after();
          uint32_t d1[16], d2[16], d3[16];
          tmem_ld16_issue(d_base + b * H_TMEM_BLK, d1);
          tmem_ld16_issue(d_base + b * H_TMEM_BLK + H_NB, d2);
          tmem_ld16_issue(d_base + b * H_TMEM_BLK + 2 * H_NB, d3);
          float sc[16];  // fetched while the TMEM loads are in flight, not live across the wait above
#pragma unroll
          for (int c4 = 0; c4 < 4; ++c4) {
            const float4 q4 = (FSCALE && fwd) ? *reinterpret_cast<const float4*>(fsc + c0 + 4 * c4) : make_float4(1.f, 1.f, 1.f, 1.f);
            sc[4 * c4] = q4.x; sc[4 * c4 + 1] = q4.y; sc[4 * c4 + 2] = q4.z; sc[4 * c4 + 3] = q4.w;
          }
          tmem_ld_wait();
          const int f = b * 128 + f0;
          if (f < Y.next_kpad) {
            const bool live = f < Y.M_true;
            float v[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) {
              float z = live ? fmaf(__uint_as_float(d1[c]) + (__uint_as_float(d2[c]) + __uint_as_float(d3[c])), inv, bias[b] * sc[c]) : 0.f;
              if (fwd) {
                if (z > 0.f) mw |= 1u << (b * 16 + c);
                z = fmaxf(z, 0.f);
              } else {
                z = ((mw >> (b * 16 + c)) & 1u) ? z : 0.f;
              }
              v[c] = z;
            }
            if (!(TIMED && (P.exp_ & 2))) store_row16(dst, f, v);
          }
          publish(b);
          if (timed) t_epi += clock64() - tq;
        }
      }
      if (fwd) maskp[ct] = mw;
      ++lc;
    };
    // last layer of a pass (<= 32 output features): every warp keeps the barrier phase, the two
    // f-warps load their 16 columns of the accumulator: out[c] = (d1 + d2) * inv_scale
    auto final_load = [&](const HLayer& Y, float (&out)[16]) {
      const uint32_t d_base = tmem_base + (WIDE ? 0u : (lc & 1) * H_TMEM_BUF) + t_lane + c0;
      const float inv = *Y.inv_scale * fmaf(P.acc_comp, (float)Y.ksteps, 1.f);   // (un-biasing of the truncating accumulation)
      if (timed) tq = clock64();
      mbar_wait_a(acc_sa, acc_ph0);
      acc_ph0 ^= 1;
      if (timed) t_fin += clock64() - tq;
      tc_fence_after();
      if (fwarp) {
        uint32_t d1[16], d2[16], d3[16];
        tmem_ld16_issue(d_base, d1);
        tmem_ld16_issue(d_base + H_NB, d2);
        tmem_ld16_issue(d_base + 2 * H_NB, d3);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 16; ++c)
          out[c] = (__uint_as_float(d1[c]) + (__uint_as_float(d2[c]) + __uint_as_float(d3[c]))) * inv;
      }
      ++lc;
    };
    // ---- staging buffers: step t's slices of the per-CTA scratch, copied one step ahead with
    // cp.async into buffer (t & 1) (every source was written at least one sweep earlier)
    auto stbuf = [&](int t) { return st_s + (t % 3) * L.st_rows * H_SROW; };
    auto cp4 = [&](float* dst, const float* src) {
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(smem_u32(dst)), "l"(src) : "memory");
    };
    auto prefetch = [&](int t, bool bwd) {
      float* B = stbuf(t);
      for (int e = ct; e < m * H_NB; e += H_COMPUTE) {
        const int j = e / H_NB, r = e - j * H_NB;
        cp4(B + (L.o_pu + j) * H_SROW + r, wsU + t * m * H_NB + e);
        if (bwd && adam) {
          cp4(B + (L.o_pm + j) * H_SROW + r, wsM + t * m * H_NB + e);
          cp4(B + (L.o_pv + j) * H_SROW + r, wsV + t * m * H_NB + e);
        }
      }
      if (need_goal) {
        for (int e = ct; e < n * H_NB; e += H_COMPUTE) {
          const int i = e / H_NB, r = e - i * H_NB;
          cp4(B + (L.o_pg + i) * H_SROW + r, wsG + t * n * H_NB + e);
          if (bwd) cp4(B + (L.o_px + i) * H_SROW + r, wsX + t * n * H_NB + e);
        }
        if (bwd && cost_mode && ct < H_NB) {  // the lanes of warp 2 wrote these (norms_finish)
          cp4(B + L.o_su * H_SROW + ct, wsS + (t * 2 + 0) * H_NB + ct);
          cp4(B + L.o_sd * H_SROW + ct, wsS + (t * 2 + 1) * H_NB + ct);
        }
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
    };
    // every step commits exactly one group (possibly empty): "all but the newest group landed"
    auto prefetch_none = [&]() { asm volatile("cp.async.commit_group;" ::: "memory"); };
    auto prefetch_wait1 = [&]() { asm volatile("cp.async.wait_group 1;" ::: "memory"); };
    // u rows (features n .. n+m-1) of the forward operand q = [x ; u], from the staged U[t]
    auto sb_u_rows = [&](int t) {
      const float* pu = stbuf(t) + L.o_pu * H_SROW;
      const float* fs = fsc_s + (t & 1) * FSR;
      for (int e = ct; e < m * H_NB; e += H_COMPUTE) {
        const int j = e / H_NB, r = e - j * H_NB;
        const float uv = FSCALE ? pu[j * H_SROW + r] * fs[r] : pu[j * H_SROW + r];
        opmax = fmaxf(opmax, fabsf(uv));
        h16_store_op(SB, n + j, r, uv);
      }
    };
    // staging-cost norms of step t, phase A: every warp sums the features i = wi, wi+8, ... of its
    // lane's trajectory (balanced: no warp does a serial per-trajectory loop on the critical path)
    const int wi = warp - 2;
    auto norms_partial = [&](int t) {
      const float* B = stbuf(t);
      float dd = 0.f, uu = 0.f;
      for (int i = wi; i < n; i += 8) {
        const float d = x_s[i * H_SROW + lane] - B[(L.o_pg + i) * H_SROW + lane];
        dd = fmaf(d, d, dd);
      }
      for (int j = wi; j < m; j += 8) {
        const float u = B[(L.o_pu + j) * H_SROW + lane];
        uu = fmaf(u, u, uu);
      }
      float* pt = part_s + (t & 1) * 16 * H_SROW;
      pt[wi * H_SROW + lane] = dd;
      pt[(8 + wi) * H_SROW + lane] = uu;
    };
    // phase B (one warp, after a barrier): combine the 8 partials; returns this lane's cost term
    auto norms_finish = [&](int t, float& Jr) {
      const float* pt = part_s + (t & 1) * 16 * H_SROW;
      float dd = 0.f, uu = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        dd += pt[k * H_SROW + lane];
        uu += pt[(8 + k) * H_SROW + lane];
      }
      if (cost_mode) {
        const float su = sqrtf(uu + a2), sd = sqrtf(dd + a2);
        Jr += w0 * (su - ALPHA) + w1 * (sd - ALPHA);
        wsS[(t * 2 + 0) * H_NB + lane] = su;  // reused by the adjoint sweep of this iteration
        wsS[(t * 2 + 1) * H_NB + lane] = sd;
      } else {
        Jr += dd;
      }
    };
    // per-column maximum over the 32 lanes of v[16] (transpose-reduce, 16 shuffles): lane l ends
    // up with the maximum of column (l >> 1) & 15
    auto colmax16 = [&](const float (&v)[16]) -> float {
      float a8[8], a4[4], a2_[2], a1;
      const bool h4 = lane & 16, h3 = lane & 8, h2 = lane & 4, h1 = lane & 2;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const float keep = h4 ? v[8 + k] : v[k], give = h4 ? v[k] : v[8 + k];
        a8[k] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 16));
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const float keep = h3 ? a8[4 + k] : a8[k], give = h3 ? a8[k] : a8[4 + k];
        a4[k] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 8));
      }
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const float keep = h2 ? a4[2 + k] : a4[k], give = h2 ? a4[k] : a4[2 + k];
        a2_[k] = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 4));
      }
      {
        const float keep = h1 ? a2_[1] : a2_[0], give = h1 ? a2_[0] : a2_[1];
        a1 = fmaxf(keep, __shfl_xor_sync(0xFFFFFFFFu, give, 2));
      }
      return fmaxf(a1, __shfl_xor_sync(0xFFFFFFFFu, a1, 1));
    };

    for (int ti = 0; ti < n_iter; ++ti) {
      const int tile = (int)blockIdx.x + ti * (int)gridDim.x;
      if (tile >= P.ntiles) break;
      const long long q0 = (long long)tile * H_NB;
      named_bar_sync(1, H_COMPUTE);
      // ---------------------------------------------------------------- stage the tile
      for (int e = ct; e < H_NB * n; e += H_COMPUTE) {  // x0 is row 0 of the tile's state scratch
        const int r = e / n, i = e - r * n;
        const long long qq = q0 + r;
        wsX[i * H_NB + r] = (qq < P.NQ) ? P.x0[(qq / P.K) * n + i] : 0.f;
      }
      if (P.goal != nullptr) {
        const int per = (T + 1) * n;
        for (int e = ct; e < H_NB * per; e += H_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          wsG[rest * H_NB + r] = (qq < P.NQ) ? P.goal[(qq / P.K) * per + rest] : 0.f;
        }
      }
      {
        const int per = T * m;
        for (int e = ct; e < H_NB * per; e += H_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          wsU[rest * H_NB + r] = (qq < P.NQ) ? P.U_in[qq * per + rest] : 0.f;
          if (adam) {
            wsM[rest * H_NB + r] = 0.f;
            wsV[rest * H_NB + r] = 0.f;
          }
        }
      }
      __threadfence_block();
      named_bar_sync(1, H_COMPUTE);
      const long long qr = q0 + ct;
      const bool rvalid = (ct < H_NB) && (qr < P.NQ);
      float Jr = 0.f;

      for (int it = 0;; ++it) {
        const bool last = (it == P.iters);
        if (last && !P.final_fwd) break;
        // -------------------------------------------------------------- forward rollout
        // One CTA barrier per step (mid-step).  Everything that is not needed to publish the next
        // operand is scheduled into the shadow of the big layers' MMAs.
        named_bar_sync(1, H_COMPUTE);  // the previous sweep's last update is complete
        prefetch(0, false);
        if (T > 1) prefetch(1, false); else prefetch_none();
        if (ct < H_NB) {  // operand scale of step 0 from x0 (>= 1 so small states never scale the actions up)
          float mx = 1.f;
          for (int i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(wsX[i * H_NB + ct]));
          const float s0 = FSCALE ? pow2_scale_to_8(mx) : 1.f;
          fsc_s[ct] = s0;
          fsc_s[FSR + ct] = s0;  // step 1 uses the scale of x_0 as well (scales lag one step)
        }
        named_bar_sync(1, H_COMPUTE);
        for (int e = ct; e < n * H_NB; e += H_COMPUTE) {
          const int i = e / H_NB, r = e - i * H_NB;
          const float v = wsX[e];
          x_s[i * H_SROW + r] = v;
          const float vs = FSCALE ? v * fsc_s[r] : v;
          opmax = fmaxf(opmax, fabsf(vs));
          h16_store_op(SB, i, r, vs);
        }
        Jr = 0.f;
        prefetch_wait1();  // slices of step 0 landed (step 1 may still be in flight)
        named_bar_sync(1, H_COMPUTE);
        sb_u_rows(0);
        publish(0);
        for (int t = 0; t < T; ++t) {
          const HDir& D = P.dir[DIR_DYN_F];
          for (int l = 0; l < D.L - 1; ++l) {
            probe_layer = timed && (l == 1);
            hidden_epilogue(D.layer[l], l, true, wsMask + ((size_t)t * (Ld - 1) + l) * MSTR, fsc_s + (t & 1) * FSR);
            probe_layer = false;
            if (l == 0) {
              // layer 0 has consumed SB.  Fetch step t+1's slices, then (barrier) x_t and the
              // slices of step t are visible to every warp: partial norms of step t, finish the
              // norms of step t-1, and write the u rows of the next operand.
              if (t + 2 < T) prefetch(t + 2, false); else prefetch_none();
              prefetch_wait1();  // slices of step t+1 landed
              named_bar_sync(1, H_COMPUTE);
              if (need_goal) {
                if (warp == 2 && t > 0) norms_finish(t - 1, Jr);
                norms_partial(t);
              }
              if (t + 1 < T) sb_u_rows(t + 1);
            }
          }
          // step boundary: x_{t+1} = x_t + Dense(h) ; the f-warps write the next operand rows
          float o[16];
          const HLayer& Yf = D.layer[D.L - 1];
          const float bias = (fwarp && lane < n) ? Yf.bias[lane] : 0.f;
          float xo[16];
          if (fwarp && lane < n) {
#pragma unroll
            for (int c = 0; c < 16; ++c) xo[c] = x_s[lane * H_SROW + c0 + c];
          }
          final_load(Yf, o);
          const bool more = (t + 1 < T) || P.use_cost;
          if (fwarp && lane < n) {
            // scales of the operand in flight (step t) and of the next one, as vector loads
            const float4* fcur = reinterpret_cast<const float4*>(fsc_s + (t & 1) * FSR + c0);
            const float4* fnxt = reinterpret_cast<const float4*>(fsc_s + ((t + 1) & 1) * FSR + c0);
            float v[16];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 one4 = make_float4(1.f, 1.f, 1.f, 1.f);
              const float4 a = FSCALE ? fcur[c4] : one4, b = FSCALE ? fnxt[c4] : one4;
              const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                const int c = 4 * c4 + k;
                // the layer ran in scaled units; 1/s is exact (power of two)
                xo[c] = fmaf(o[c], pow2_recip(av[k]), bias) + xo[c];
                v[c] = xo[c] * bv[k];
              }
            }
            if (more) store_row16(SB, lane, v);
          }
          if (more) publish(0);
          if (fwarp) {  // off the critical path: keep x_{t+1} for the costs and the adjoint
            float xabs[16];
#pragma unroll
            for (int c = 0; c < 16; ++c) xabs[c] = lane < n ? fabsf(xo[c]) : 0.f;
            if (lane < n) {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                x_s[lane * H_SROW + c0 + c] = xo[c];
                wsX[(size_t)(t + 1) * n * H_NB + lane * H_NB + c0 + c] = xo[c];
              }
            }
            // scale of step t+2 from x_{t+1} (one step of lag), written over the row of step t:
            // its last readers were this warp's loads above and the epilogues of step t
            if (FSCALE) {
              const float mx = colmax16(xabs);
              __syncwarp();
              if (!(lane & 1)) fsc_s[(t & 1) * FSR + c0 + ((lane >> 1) & 15)] = pow2_scale_to_8(fmaxf(mx, 1.f));
              __syncwarp();
            }
          }
        }
        __threadfence_block();  // wsS / wsX of this sweep are read back through cp.async
        named_bar_sync(1, H_COMPUTE);  // x_T and the last partial norms are visible
        if (warp == 2 && need_goal) norms_finish(T - 1, Jr);
        // the adjoint sweep starts at t = T-1: fetch its slices under the cost MLP passes
        // (su/sd of step T-1 are written by warp 2 just above: it fetches those itself)
        if (!last) prefetch(T - 1, true);
        // -------------------------------------------------------------- terminal cost
        if (P.use_cost) {
          const HDir& D = P.dir[DIR_COST_F];
          for (int l = 0; l < D.L - 1; ++l)
            hidden_epilogue(D.layer[l], l, true, costMask + (size_t)l * MSTR, fsc_s + (T & 1) * FSR);
          float o[16];
          const HLayer& Yf = D.layer[D.L - 1];
          const float bias = (fwarp && lane < P.fout) ? Yf.bias[lane] : 0.f;
          final_load(Yf, o);
          if (fwarp && lane < P.fout) {
            const float* fcur = fsc_s + (T & 1) * FSR;
#pragma unroll
            for (int c = 0; c < 16; ++c) y_s[lane * H_SROW + c0 + c] = fmaf(o[c], FSCALE ? pow2_recip(fcur[c0 + c]) : 1.f, bias);
          }
          named_bar_sync(1, H_COMPUTE);
          if (ct < H_NB) {
            float yy = 0.f, mx = 0.f;
            const float s2 = 2.f * w2;
            for (int oo = 0; oo < P.fout; ++oo) {
              const float y = y_s[oo * H_SROW + ct];
              yy = fmaf(y, y, yy);
              mx = fmaxf(mx, fabsf(s2 * y));
            }
            Jr += w2 * yy;
            if (!last) {  // adjoint seed dJ/dy = 2 w2 y, scaled per trajectory into fp16 range
              const float sc = pow2_scale_to_8(mx);
              isc_s[ct] = 1.f / sc;
              for (int oo = 0; oo < P.fout; ++oo) h16_store_op(SB, oo, ct, s2 * y_s[oo * H_SROW + ct] * sc);
            }
          }
        } else if (P.mode == MODE_L2GRAD) {
          if (ct < H_NB) {
            float dd = 0.f;
            for (int i = 0; i < n; ++i) {
              const float d = x_s[i * H_SROW + ct] - wsG[(T * n + i) * H_NB + ct];
              dd = fmaf(d, d, dd);
            }
            Jr = (Jr + dd) / (float)(T + 1);
          }
        }
        if (last) break;
        // -------------------------------------------------------------- adjoint seed lambda_T
        if (P.use_cost) {
          named_bar_sync(1, H_COMPUTE);  // isc_s of the seed is visible to the f-warps
          publish(0);  // seed operand of the cost MLP's backward pass
          const HDir& D = P.dir[DIR_COST_B];
          for (int lb = 0; lb < D.L - 1; ++lb)
            hidden_epilogue(D.layer[lb], lb, false, costMask + (size_t)(D.L - 2 - lb) * MSTR, nullptr);
          float o[16];
          final_load(D.layer[D.L - 1], o);
          if (fwarp && lane < n) {
#pragma unroll
            for (int c = 0; c < 16; ++c) lam_s[lane * H_SROW + c0 + c] = o[c] * isc_s[c0 + c];
          }
        } else if (ct < H_NB) {
          for (int i = 0; i < n; ++i)
            lam_s[i * H_SROW + ct] = l2scale * (x_s[i * H_SROW + ct] - wsG[(T * n + i) * H_NB + ct]);
        }
        named_bar_sync(1, H_COMPUTE);
        if (ct < H_NB) {  // exact per-trajectory scale of lambda_T
          float mx = 0.f;
          for (int i = 0; i < n; ++i) mx = fmaxf(mx, fabsf(lam_s[i * H_SROW + ct]));
          snx_s[ct] = pow2_scale_to_8(mx);
        }
        named_bar_sync(1, H_COMPUTE);
        for (int e = ct; e < n * H_NB; e += H_COMPUTE) {
          const int i = e / H_NB, r = e - i * H_NB;
          const float lam = lam_s[i * H_SROW + r];
          h16_store_op(SB, i, r, lam * snx_s[r]);
          if (P.lam_out != nullptr && q0 + r < P.NQ) P.lam_out[((q0 + r) * (T + 1) + T) * n + i] = lam;
        }
        named_bar_sync(1, H_COMPUTE);
        if (ct < H_NB) isc_s[ct] = 1.f / snx_s[ct];  // scale in flight; snx_s stays the lagged estimate
        float bc1 = 1.f, bc2 = 1.f;
        if (adam) {
          bc1 = (float)(1.0 - pow((double)P.b1, (double)(it + 1)));
          bc2 = (float)(1.0 - pow((double)P.b2, (double)(it + 1)));
        }
        named_bar_sync(1, H_COMPUTE);  // isc_s visible to the f-warps of the first boundary
        publish(0);  // lambda_T operand of the first adjoint step
        // -------------------------------------------------------------- adjoint sweep + update
        // action gradient + update of step tu, one (feature, trajectory) per thread; runs one step
        // late (after the next step's mid barrier) so it never sits between a boundary and the
        // first epilogue of the following step
        auto update = [&](int tu) {
          const float* Bu = stbuf(tu);
          for (int e = ct; e < m * H_NB; e += H_COMPUTE) {
            const int j = e / H_NB, r = e - j * H_NB;
            const int ix = tu * m * H_NB + e;
            float u = Bu[(L.o_pu + j) * H_SROW + r];
            float g = dq_s[((tu & 1) * (n + m) + n + j) * H_SROW + r];
            if (cost_mode) g = (w0 * u) / Bu[L.o_su * H_SROW + r] + g;
            if (P.mode == MODE_PLAN) {
              if (P.method == 0) {
                u = u - P.lr * g;
              } else {
                const float mo = P.b1 * Bu[(L.o_pm + j) * H_SROW + r] + (1.f - P.b1) * g;
                const float ve = P.b2 * Bu[(L.o_pv + j) * H_SROW + r] + (1.f - P.b2) * g * g;
                wsM[ix] = mo;
                wsV[ix] = ve;
                u = u - P.lr * (mo / bc1) / (sqrtf(ve / bc2) + P.eps);
              }
              wsU[ix] = u;
            } else if (P.dU_out != nullptr && q0 + r < P.NQ) {
              P.dU_out[((q0 + r) * T + tu) * m + j] = g;
            }
          }
        };
        for (int t = T - 1; t >= 0; --t) {
          const HDir& D = P.dir[DIR_DYN_B];
          const float* B = stbuf(t);
          for (int lb = 0; lb < D.L - 1; ++lb) {
            hidden_epilogue(D.layer[lb], lb, false,
                            wsMask + ((size_t)t * (Ld - 1) + (D.L - 2 - lb)) * MSTR, nullptr);
            if (lb == 0) {
              if (t > 0) prefetch(t - 1, true); else prefetch_none();
              prefetch_wait1();  // slices of step t landed
              named_bar_sync(1, H_COMPUTE);  // staging buffer t and dq of step t+1 are visible
              if (t + 1 < T) update(t + 1);
            }
          }
          // step boundary: lambda_t = l_x + lambda_{t+1} + dq_x ; u gradient kept for the update
          float base[16], isc[16], snx[16];
          if (fwarp && lane < n + m) {
#pragma unroll
            for (int c = 0; c < 16; ++c) isc[c] = isc_s[c0 + c];
            if (lane < n) {
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float d = B[(L.o_px + lane) * H_SROW + c0 + c] - B[(L.o_pg + lane) * H_SROW + c0 + c];
                const float cc = cost_mode ? (w1 * d) / B[L.o_sd * H_SROW + c0 + c] : l2scale * d;
                base[c] = cc + lam_s[lane * H_SROW + c0 + c];
                snx[c] = snx_s[c0 + c];
              }
            }
          }
          float o[16];
          final_load(D.layer[D.L - 1], o);
          float lamabs[16];
#pragma unroll
          for (int c = 0; c < 16; ++c) lamabs[c] = 0.f;
          if (fwarp && lane < n + m) {
            if (lane < n) {
              float v[16];
#pragma unroll
              for (int c = 0; c < 16; ++c) {
                const float lam = base[c] + o[c] * isc[c];
                base[c] = lam;
                lamabs[c] = fabsf(lam);
                v[c] = lam * snx[c];
              }
              if (t > 0) store_row16(SB, lane, v);
            }
          }
          if (t > 0) publish(0);
          if (fwarp) {  // off the critical path
            if (lane < n) {
#pragma unroll
              for (int c = 0; c < 16; ++c) lam_s[lane * H_SROW + c0 + c] = base[c];
              if (P.lam_out != nullptr) {
#pragma unroll
                for (int c = 0; c < 16; ++c)
                  if (q0 + c0 + c < P.NQ) P.lam_out[((q0 + c0 + c) * (T + 1) + t) * n + lane] = base[c];
              }
            } else if (lane < n + m) {  // u part of the input adjoint, consumed by update(t) one step later
#pragma unroll
              for (int c = 0; c < 16; ++c) dq_s[((t & 1) * (n + m) + lane) * H_SROW + c0 + c] = o[c] * isc[c];
            }
            // the operand just written used snx_s; the next one uses the scale of lambda_t
            // (columns c0..c0+15 belong to this warp alone: a warp-level sync orders the update)
            const float mx = colmax16(lamabs);
            __syncwarp();
            if (t > 0 && !(lane & 1)) {
              const int col = c0 + ((lane >> 1) & 15);
              isc_s[col] = pow2_recip(snx_s[col]);
              snx_s[col] = pow2_scale_to_8(mx);
            }
            __syncwarp();
          }
        }
        named_bar_sync(1, H_COMPUTE);  // dq of step 0 visible
        update(0);
        if (P.mode != MODE_PLAN) break;
        __threadfence_block();  // this sweep's wsU/wsM/wsV are read back through cp.async next sweep
      }
      named_bar_sync(1, H_COMPUTE);
      // ---------------------------------------------------------------- write the tile out
      if (P.J_out != nullptr && rvalid) P.J_out[qr] = Jr;
      if (P.U_out != nullptr) {
        const int per = T * m;
        for (int e = ct; e < H_NB * per; e += H_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          if (qq < P.NQ) P.U_out[qq * per + rest] = wsU[rest * H_NB + r];
        }
      }
      if (P.X_out != nullptr) {
        const int per = (T + 1) * n;
        for (int e = ct; e < H_NB * per; e += H_COMPUTE) {
          const int r = e / per, rest = e - r * per;
          const long long qq = q0 + r;
          if (qq < P.NQ) P.X_out[qq * per + rest] = wsX[rest * H_NB + r];
        }
      }
    }
    // fp16 operand range check (the epilogue clamps with satfinite; the host re-plans on the fp32
    // CUDA-core kernel when this fires, see gmpc_plan_host)
    if (!(opmax <= 65000.f) && P.ovf != nullptr) atomicAdd(P.ovf, 1u);
    if (TIMED && ct == 0) {
      P.dbg[blockIdx.x * 16 + 4] = t_acc;
      P.dbg[blockIdx.x * 16 + 5] = t_epi;
      P.dbg[blockIdx.x * 16 + 6] = t_fin;
      P.dbg[blockIdx.x * 16 + 7] = clock64() - t_total0;
      P.dbg[blockIdx.x * 16 + 8] = t_bnd;
      P.dbg[blockIdx.x * 16 + 11] = t_o1;
      P.dbg[blockIdx.x * 16 + 12] = t_o2;
      P.dbg[blockIdx.x * 16 + 13] = t_o3;
      P.dbg[blockIdx.x * 16 + 14] = t_o4;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (C > 1) cluster_sync_all();  // no CTA leaves while a peer may still multicast into it
  if (warp == 0) {
    __syncwarp();
    tmem_dealloc(tmem_base, 512);
  }
}

// Pack one Dense kernel W[K][N] (flax layout) into the block-major unit stream of one direction,
// scaled by the power of two that puts max |W| in [2^10, 2^11).
//   transposed == 0 (forward):  A rows r = output feature n, reduction kk = input feature k.
//   transposed == 1 (adjoint):  A rows r = input feature k,  reduction kk = output feature n.
// Block b stores rows_b rows: k-step image = [hi unit | lo unit], unit = [2 k-chunks][rows_b][8 halfs].
__global__ void h16_pack_kernel(const float* __restrict__ W, int K, int N, int transposed,
                                uint8_t* dst, int ksteps, int rows0, int rows1, int rows2, int rows3,
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
  if (idx == 0) *inv_scale = 1.f / sc;
  if (idx >= K * N) return;
  const int k = idx / N, o = idx - k * N;
  const int r = transposed ? k : o, kk = transposed ? o : k;
  __half hi, lo;
  split_h1(W[idx] * sc, hi, lo);
  const int b = r >> 7, rr = r & 127, j = kk >> 4, k16 = kk & 15;
  const int rowsv[4] = {rows0, rows1, rows2, rows3};
  const int rows = rowsv[b];
  const size_t unit = (size_t)rows * 32;
  size_t blk_base = 0;
  for (int bb = 0; bb < b; ++bb) blk_base += (size_t)ksteps * 2 * rowsv[bb] * 32;
  uint8_t* p = dst + blk_base + (size_t)j * 2 * unit + (size_t)(k16 >> 3) * rows * 16 + (rr >> 3) * H_A_SBO +
               (rr & 7) * 16 + (k16 & 7) * 2;
  *reinterpret_cast<__half*>(p) = hi;
  *reinterpret_cast<__half*>(p + unit) = lo;
}

// ------------------------------------------------------------------------------------- host side
using H16Kernel = void (*)(const HParams);
// wide layers always run with forward scaling (wide random dynamics over long horizons diverge)
inline H16Kernel h16_kernel_ptr(bool timed, bool wide, bool fscale) {
  if (wide) return timed ? plan_h16_kernel<true, true, true> : plan_h16_kernel<false, true, true>;
  if (fscale) return timed ? plan_h16_kernel<true, false, true> : plan_h16_kernel<false, false, true>;
  return timed ? plan_h16_kernel<true, false, false> : plan_h16_kernel<false, false, false>;
}

struct H16State {
  bool supported = false;
  std::string why = "not initialised";
  // The fp32 accumulation of tcgen05.mma truncates (rounds towards zero): every accumulate event shrinks the running
  // sum, which shows as a uniform relative shrink of the rollout (C2 dims: 6.4e-6 on X after 32 steps, fp32 FMA:
  // 1.9e-7).  The epilogue multiplies the accumulator by 1 + acc_comp * k-steps; calibrated on the hardware
  // (tools/acc_comp_calib.py, profiles/r2_acc_comp_calib.txt); GMPC_H16_ACC_COMP (units of 2^-24) overrides, 0 disables.
  float acc_comp = 0.30f * 5.9604644775390625e-08f;   // 0.30 x 2^-24 per accumulate event (measured)
  int dyn_dims[MAXL + 1], cost_dims[MAXL + 1], Ld = 0, Lc = 0;
  uint8_t* d_stream = nullptr;
  size_t stream_bytes = 0;
  HDir dir[4];
  float* d_bias = nullptr;     // forward biases, packed
  float* d_scale = nullptr;    // [Ld + Lc] inverse weight scales
  float* d_wsS = nullptr;      // [num_sms][T][2][32] staging-cost norms saved by the forward sweep
  uint2* d_gtab = nullptr;     // group tables of the four passes, concatenated
  uint32_t* d_ovf = nullptr;   // operand range-check counter (see HParams::ovf)
  uint32_t gtab_off[4] = {0, 0, 0, 0}, ngroups[4] = {0, 0, 0, 0};
  uint32_t* d_absmax = nullptr;
  uint32_t hb_bytes = 0;
  int nslot = 0;
  int wide = 0;                // hidden width in (256, 512]: serial layer schedule
  size_t smem_bytes = 0;
  int num_sms = 0;
  int n = 0, m = 0, fout = 0;
  uint32_t exp_ = 0;  // GMPC_H16_EXP, read once at create (timing experiments of the TIMED build)
  int cluster = 1;  // B200: multicast did not pay for this stream (12.6 ms vs 13.0 ms on C2)
  int max_clusters[5] = {0, 0, 0, 0, 0};
  int last_cluster = 1;
  long long* d_dbg = nullptr;
};

inline void h16_layer_geom(HLayer& Y, int M_true, int red_true) {
  Y.M_true = M_true;
  Y.nblk = (M_true + 127) / 128;
  Y.ksteps = rup(red_true, 16) / 16;
  for (int b = 0; b < 4; ++b)
    Y.rows[b] = b < Y.nblk - 1 ? 128 : (b == Y.nblk - 1 ? std::max(32, rup(M_true - 128 * b, 16)) : 0);
  Y.next_kpad = rup(M_true, 16);
  Y.bias = nullptr;
  Y.gsrc = nullptr;
  Y.inv_scale = nullptr;
}

inline size_t h16_layer_bytes(const HLayer& Y) {
  return (size_t)Y.ksteps * 64 * (Y.rows[0] + Y.rows[1] + Y.rows[2] + Y.rows[3]);
}

inline size_t h16_build_geometry(H16State& S) {
  size_t off = 0;
  auto one = [&](const int* dims, int Ln, HDir& F, HDir& Bw) {
    F.L = Bw.L = Ln;
    F.pad_ = Bw.pad_ = 0;
    for (int l = 0; l < Ln; ++l) {
      h16_layer_geom(F.layer[l], dims[l + 1], dims[l]);
      F.layer[l].gsrc = reinterpret_cast<const uint8_t*>(off);
      off += h16_layer_bytes(F.layer[l]);
    }
    for (int i = 0; i < Ln; ++i) {
      const int lt = Ln - 1 - i;
      h16_layer_geom(Bw.layer[i], dims[lt], dims[lt + 1]);
      Bw.layer[i].gsrc = reinterpret_cast<const uint8_t*>(off);
      off += h16_layer_bytes(Bw.layer[i]);
    }
  };
  one(S.dyn_dims, S.Ld, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B]);
  one(S.cost_dims, S.Lc, S.dir[DIR_COST_F], S.dir[DIR_COST_B]);
  return off;
}

inline int h16_create(H16State& S, const gmpc_config& c, const int* dyn_dims, const int* cost_dims,
                      const cudaDeviceProp& prop) {
  S.Ld = c.dyn_layers;
  S.Lc = c.cost_layers;
  for (int i = 0; i <= S.Ld; ++i) S.dyn_dims[i] = dyn_dims[i];
  for (int i = 0; i <= S.Lc; ++i) S.cost_dims[i] = cost_dims[i];
  S.num_sms = prop.multiProcessorCount;
  S.n = c.n; S.m = c.m; S.fout = c.cost_fout;
  int hmax = 16;
  for (int i = 1; i < S.Ld; ++i) hmax = std::max(hmax, dyn_dims[i]);
  for (int i = 1; i < S.Lc; ++i) hmax = std::max(hmax, cost_dims[i]);
  S.supported = false;
  if (S.Ld < 2) { S.why = "dynamics MLP has no hidden layer"; return GMPC_OK; }
  if (hmax > 512) { S.why = "hidden width > 512 (four 128-row MMA blocks)"; return GMPC_OK; }
  S.wide = hmax > 256 ? 1 : 0;
  if (c.n + c.m > 32 || c.cost_fout > 32) { S.why = "n+m or fout > 32"; return GMPC_OK; }
  S.hb_bytes = (uint32_t)(rup(hmax, 16) / 8) * H_B_LBO;
  const size_t budget = (size_t)prop.sharedMemPerBlockOptin;
  S.stream_bytes = h16_build_geometry(S);
  int ng_total = 0;
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < S.dir[d].L; ++l) ng_total += S.dir[d].layer[l].nblk * ((S.dir[d].layer[l].ksteps + 1) / 2);
  const HSmem L0 = h_smem_layout(0, S.hb_bytes, c.n, c.m, c.cost_fout, ng_total, S.wide);
  int nslot = (int)((budget - std::min(budget, (size_t)L0.total)) / H_SLOT_BYTES);
  nslot = std::min(nslot, H_MAX_SLOTS);
  if (const char* env = getenv("GMPC_H16_ACC_COMP")) S.acc_comp = (float)(atof(env) * 5.9604644775390625e-08);
  if (const char* env = getenv("GMPC_H16_SLOTS")) nslot = std::min(nslot, std::max(2, atoi(env)));
  if (nslot < 4) { S.why = "shared memory"; return GMPC_OK; }
  S.nslot = nslot;
  S.smem_bytes = h_smem_layout(nslot, S.hb_bytes, c.n, c.m, c.cost_fout, ng_total, S.wide).total;
  size_t nbias = 0;
  for (int l = 0; l < S.Ld; ++l) nbias += rup(dyn_dims[l + 1], 4);
  for (int l = 0; l < S.Lc; ++l) nbias += rup(cost_dims[l + 1], 4);
  if (cudaMalloc(&S.d_stream, S.stream_bytes + 4096) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_stream, 0, S.stream_bytes + 4096) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_bias, nbias * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_scale, (S.Ld + S.Lc) * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_absmax, (S.Ld + S.Lc) * sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_wsS, (size_t)S.num_sms * c.T * 2 * H_NB * sizeof(float)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMalloc(&S.d_ovf, sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  if (cudaMemset(S.d_ovf, 0, sizeof(uint32_t)) != cudaSuccess) return GMPC_E_CUDA;
  float* bp = S.d_bias;
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < S.dir[d].L; ++l)
      S.dir[d].layer[l].gsrc = S.d_stream + reinterpret_cast<size_t>(S.dir[d].layer[l].gsrc);
  for (int l = 0; l < S.Ld; ++l) {
    S.dir[DIR_DYN_F].layer[l].bias = bp; bp += rup(dyn_dims[l + 1], 4);
    S.dir[DIR_DYN_F].layer[l].inv_scale = S.d_scale + l;
    S.dir[DIR_DYN_B].layer[S.Ld - 1 - l].inv_scale = S.d_scale + l;
  }
  for (int l = 0; l < S.Lc; ++l) {
    S.dir[DIR_COST_F].layer[l].bias = bp; bp += rup(cost_dims[l + 1], 4);
    S.dir[DIR_COST_F].layer[l].inv_scale = S.d_scale + S.Ld + l;
    S.dir[DIR_COST_B].layer[S.Lc - 1 - l].inv_scale = S.d_scale + S.Ld + l;
  }
  {  // ring groups of every pass: two k-steps of one block (one when the block's k-step count is odd)
    std::vector<uint2> tab;
    for (int d = 0; d < 4; ++d) {
      S.gtab_off[d] = (uint32_t)tab.size();
      uint32_t off = 0;
      for (int l = 0; l < S.dir[d].L; ++l) {
        const HLayer& Y = S.dir[d].layer[l];
        for (int b = 0; b < Y.nblk; ++b)
          for (int j = 0; j < Y.ksteps; j += 2) {
            const uint32_t bytes = (uint32_t)std::min(2, Y.ksteps - j) * 64 * Y.rows[b];
            tab.push_back(make_uint2(off, bytes));
            off += bytes;
          }
      }
      S.ngroups[d] = (uint32_t)tab.size() - S.gtab_off[d];
    }
    if (cudaMalloc(&S.d_gtab, tab.size() * sizeof(uint2)) != cudaSuccess) return GMPC_E_CUDA;
    if (cudaMemcpy(S.d_gtab, tab.data(), tab.size() * sizeof(uint2), cudaMemcpyHostToDevice) != cudaSuccess)
      return GMPC_E_CUDA;
  }
  for (int t = 0; t < 2; ++t)
    for (int f = 0; f < 2; ++f)
      if (cudaFuncSetAttribute(h16_kernel_ptr(t, S.wide, f), cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)prop.sharedMemPerBlockOptin) != cudaSuccess)   // (per kernel, not per handle: the device maximum)
        return GMPC_E_CUDA;
  for (int C = 2; C <= 4; C *= 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(prop.multiProcessorCount / C * C);
    cfg.blockDim = dim3(H_THREADS);
    cfg.dynamicSmemBytes = S.smem_bytes;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int nc = 0;
    if (cudaOccupancyMaxActiveClusters(&nc, h16_kernel_ptr(false, S.wide, false), &cfg) != cudaSuccess) nc = 0;
    S.max_clusters[C] = nc;
  }
  cudaGetLastError();
  if (const char* env = getenv("GMPC_TC_CLUSTER")) S.cluster = atoi(env);
  if (const char* env = getenv("GMPC_H16_EXP")) S.exp_ = (uint32_t)atoi(env);
  if (S.cluster != 1 && S.cluster != 2 && S.cluster != 4) S.cluster = 1;
  while (S.cluster > 1 && S.max_clusters[S.cluster] <= 0) S.cluster >>= 1;
  if (getenv("GMPC_DEBUG")) {
    cudaMalloc(&S.d_dbg, sizeof(long long) * 16 * 1024);
    cudaMemset(S.d_dbg, 0, sizeof(long long) * 16 * 1024);
    fprintf(stderr, "[gmpc] h16: smem %zu B, %d ring slots, max co-resident clusters: x2=%d x4=%d, using cluster=%d\n",
            S.smem_bytes, S.nslot, S.max_clusters[2], S.max_clusters[4], S.cluster);
  }
  S.supported = true;
  S.why = "";
  return GMPC_OK;
}

inline void h16_destroy(H16State& S) {
  cudaFree(S.d_stream);
  cudaFree(S.d_bias);
  cudaFree(S.d_scale);
  cudaFree(S.d_absmax);
  cudaFree(S.d_wsS);
  S.d_wsS = nullptr;
  cudaFree(S.d_gtab);
  S.d_gtab = nullptr;
  cudaFree(S.d_ovf);
  S.d_ovf = nullptr;
  cudaFree(S.d_dbg);
  S.d_stream = nullptr;
  S.d_bias = nullptr;
  S.d_scale = nullptr;
  S.d_absmax = nullptr;
  S.d_dbg = nullptr;
}

inline int h16_set_weights(H16State& S, const float* const* dyn_W, const float* const* dyn_b,
                           const float* const* cost_W, const float* const* cost_b, cudaStream_t st,
                           int64_t* launches) {
  if (!S.supported) return GMPC_OK;
  cudaMemsetAsync(S.d_absmax, 0, (S.Ld + S.Lc) * sizeof(uint32_t), st);
  cudaMemsetAsync(S.d_stream, 0, S.stream_bytes, st);
  auto one = [&](const int* dims, int Ln, const float* const* W, const float* const* b, HDir& F,
                 HDir& Bw, int sbase) {
    for (int l = 0; l < Ln; ++l) {
      const int K = dims[l], N = dims[l + 1], blocks = (K * N + 255) / 256;
      const HLayer& f = F.layer[l];
      const HLayer& r = Bw.layer[Ln - 1 - l];
      h16_absmax_kernel<<<std::min(blocks, 64), 256, 0, st>>>(W[l], K * N, S.d_absmax + sbase + l);
      h16_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 0, const_cast<uint8_t*>(f.gsrc), f.ksteps,
                                              f.rows[0], f.rows[1], f.rows[2], f.rows[3], S.d_absmax + sbase + l,
                                              S.d_scale + sbase + l);
      h16_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 1, const_cast<uint8_t*>(r.gsrc), r.ksteps,
                                              r.rows[0], r.rows[1], r.rows[2], r.rows[3], S.d_absmax + sbase + l,
                                              S.d_scale + sbase + l);
      *launches += 3;
      cudaMemcpyAsync(const_cast<float*>(f.bias), b[l], sizeof(float) * N, cudaMemcpyDeviceToDevice, st);
    }
  };
  one(S.dyn_dims, S.Ld, dyn_W, dyn_b, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B], 0);
  one(S.cost_dims, S.Lc, cost_W, cost_b, S.dir[DIR_COST_F], S.dir[DIR_COST_B], S.Ld);
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

// The tensor-core path needs a real dense contraction in the feature dimension: hidden width >= 64 (north
// star).  The batch does not matter: the weights are the 128-row MMA operand, so even one trajectory drives
// full-height MMAs, and the kernel is 4-5 x faster than the fp32 kernel at any batch (tools/plan_latency.py).
inline bool h16_worthwhile(const H16State& S, int64_t NQ) {
  int hmin = 1 << 30;
  for (int i = 1; i < S.Ld; ++i) hmin = std::min(hmin, S.dyn_dims[i]);
  return NQ >= 1 && S.Ld > 1 && hmin >= 64;
}

inline int h16_launch(H16State& S, const PlanParams& P, bool fscale, cudaStream_t st, int64_t* launches) {
  HParams Q;
  memset(&Q, 0, sizeof(Q));
  for (int d = 0; d < 4; ++d) Q.dir[d] = S.dir[d];
  Q.n = P.n; Q.m = P.m; Q.T = P.T; Q.K = P.K;
  Q.fout = P.fout; Q.mode = P.mode; Q.method = P.method; Q.iters = P.iters;
  Q.use_cost = P.use_cost; Q.final_fwd = P.final_fwd;
  Q.nslot = S.nslot;
  Q.wide = S.wide;
  Q.hb_bytes = S.hb_bytes;
  for (int d = 0; d < 4; ++d) { Q.gtab[d] = S.d_gtab + S.gtab_off[d]; Q.ngroups[d] = S.ngroups[d]; }
  Q.exp_ = S.exp_;
  Q.NQ = P.NQ;
  Q.ntiles = (int)((P.NQ + H_NB - 1) / H_NB);
  Q.lr = P.lr; Q.b1 = P.b1; Q.b2 = P.b2; Q.eps = P.eps;
  Q.acc_comp = S.acc_comp;
  Q.x0 = P.x0; Q.U_in = P.U_in; Q.goal = P.goal; Q.mpcw = P.mpcw;
  Q.U_out = P.U_out; Q.X_out = P.X_out; Q.J_out = P.J_out; Q.dU_out = P.dU_out; Q.lam_out = P.lam_out;
  Q.ws_X = P.ws_X; Q.ws_G = P.ws_G; Q.ws_U = P.ws_U; Q.ws_M = P.ws_M; Q.ws_V = P.ws_V;
  Q.ws_mask = P.ws_mask;
  Q.ws_S = S.d_wsS;
  Q.ovf = S.d_ovf;
  Q.dbg = S.d_dbg;
  if (Q.ntiles <= 0) return GMPC_OK;
  int C = S.cluster;
  while (C > 1 && Q.ntiles < C) C >>= 1;
  const int max_ctas = C > 1 ? S.max_clusters[C] * C : S.num_sms;
  int grid = std::min((Q.ntiles + C - 1) / C * C, max_ctas);
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(H_THREADS);
  cfg.dynamicSmemBytes = S.smem_bytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = C;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, h16_kernel_ptr(S.d_dbg != nullptr, S.wide, fscale), Q);
  ++*launches;
  S.last_cluster = C;
  if (S.d_dbg != nullptr && e == cudaSuccess) {
    long long hdbg[16];
    cudaStreamSynchronize(st);
    cudaMemcpy(hdbg, S.d_dbg, sizeof(hdbg), cudaMemcpyDeviceToHost);
    fprintf(stderr, "[gmpc] h16 CTA0 cycles: mma-warp wait_act %lld wait_full %lld issue %lld | compute wait_acc(hidden) %lld epilogue %lld wait_acc(final) %lld boundary %lld total %lld | dynF-L1 acc0->acc1 total %lld count %lld (x %lld %lld) | issuer2 wait_act %lld wait_full %lld\n",
            hdbg[0], hdbg[1], hdbg[2], hdbg[4], hdbg[5], hdbg[6], hdbg[8], hdbg[7], hdbg[11], hdbg[12], hdbg[13], hdbg[14], hdbg[9], hdbg[10]);
  }
  return e == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
