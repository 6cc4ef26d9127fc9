// plan_t128.cuh -- fused rollout + cost + adjoint + update kernel, 128-trajectory tiles.
//
// The contraction of a layer is issued as
//     out[128 traj x N features] = act[128 traj x K] * W[K x N]
// with the TRAJECTORIES as the M dimension of the MMA (TMEM lane = trajectory), the activations as
// the A operand IN TENSOR MEMORY and the weights streamed from L2 through a shared-memory ring as
// the B operand.  Measured on B200 (tools/t128_probe.cu, profiles/r2_t128_probe.txt): with A in TMEM
// an M=128 x N x K=16 kind::f16 MMA costs max(16, N/2) cycles, the full tensor rate, and the three
// products of the fp16 hi/lo split (ah Wh + al Wh + ah Wl, one fp32 accumulator) of one 208-column
// k-step run in 312 cycles with all 148 SMs streaming their weights from L2 at the same time.
//
// Nothing but the weights ever touches shared memory.  Tensor memory holds two 240-column regions that
// swap roles every layer, and a 32-column region for the narrow last layer of a pass:
//     layer l :  A operand = region c,  accumulator = region c^1
//     epilogue:  a thread reads 16 accumulator columns of ITS trajectory (tcgen05.ld), applies
//                scale / bias / ReLU / mask bit, splits into fp16 hi/lo and writes the 8 + 8 packed
//                columns back IN PLACE (tcgen05.st): the accumulator region turns, 16 columns at a
//                time, into the A operand of layer l+1 (feature block j: hi in columns [16 j, 16 j + 8),
//                lo in [16 j + 8, 16 j + 16)).  No thread ever touches a column another thread reads.
//     layer l+1: A operand = region c^1, accumulator = region c (all of layer l's MMAs are complete).
// Issue: ONE issuer warp walks the schedule converged, an elected lane issues; the warp-role branches are taken
// on a shuffled (provably uniform) warp index, so the loop state and the descriptors live in uniform registers
// and nothing but the three UTCHMMA of a k-step and a handful of uniform adds is executed per k-step.  (The
// accumulator columns can be split into N-parts with one issuer warp each -- T_ISSUERS -- but one, two and three
// parts run at the same speed: the k-step pace is set by the epilogue.)
// Pipelining across layers is by ROUNDS of four feature blocks: block j of the next operand belongs to epilogue
// sub-group 3 - j % 4 (4 warps, one per TMEM lane quarter; the sub-groups with step-boundary duties own the later
// blocks), every epilogue warp arrives on the round's mbarrier -- with or without a block in the round -- and the
// issuer starts k-steps 4 r .. 4 r + 3 of layer l+1 as soon as round r is stored, so the MMAs of layer l+1 run
// under the epilogue of layer l.
// Synchronisation: one mbarrier "all MMAs of the layer are complete" (one tcgen05.commit per issuer), one per
// round (16 warp arrivals), the weight ring's full / empty pairs, and one for the narrow accumulator's second
// reader.  Because all 16 warps arrive on every round of every layer no barrier can get two phases ahead of any
// of its waiters, and write-after-read on the regions needs nothing else: a round's arrival follows its owners'
// reads, and a layer completes only after every round has arrived.
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
constexpr int T_ISSUERS = 1;                       // MMA issuer warps = N-parts per layer (measured: 1, 2 and 3 parts run
                                                   // at the same speed, the k-step pace is set by the epilogue rounds)
constexpr int T_THREADS = T_EPI + 32 * (1 + T_ISSUERS);   // epilogue warps 0..15, producer warp 16, issuer warps 17..
constexpr int T_MAXP = T_ISSUERS;                  // N-parts per layer
constexpr int T_MAXKS = 16;                        // k-steps of the widest operand
constexpr int T_MAXH = 240;                        // widest operand / accumulator (hidden width, padded to 16)
constexpr int T_MAX_SLOTS = 32;
constexpr uint32_t T_R0 = 0, T_R1 = 240, T_RC = 480;  // TMEM columns: region 0 | region 1 | narrow accumulator

struct TLayer {
  uint32_t goff;           // byte offset of the layer's image in the image of the pass
  int kpg;                 // k-steps per ring group (one bulk copy, one slot)
  int ncol[T_MAXP];        // output columns (MMA N) of N-part p
  int c0[T_MAXP];          // first output column of N-part p
  int np;                  // N-parts
  int M_true;              // output features of this (possibly transposed) layer
  int npad;                // round_up(M_true, 16) = sum of ncol
  int nks;                 // reduction k-steps (round_up(K, 16) / 16)
  int bias_off;            // offset of the layer's bias in the bias table (forward layers)
  int scale_idx;           // index into the inverse weight scale table
};
struct TDir {
  TLayer layer[MAXL];
  const uint8_t* gsrc;     // image of the pass: per layer, per k-step, per N-part: [hi tile | lo tile]
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
  float acc_comp;          // un-biasing of the truncating fp32 accumulation: relative gain per accumulate event
  int pad2_;
  const float *x0, *U_in, *goal, *mpcw;
  const float* bias;       // [nbias] forward biases, padded per layer to npad
  const float* inv_scale;  // [nscale] 1 / (power-of-two weight scale) per Dense layer
  float *U_out, *X_out, *J_out, *dU_out, *lam_out;
  float *ws_X, *ws_G, *ws_U, *ws_M, *ws_V;   // per-CTA scratch, [..][128] trajectory-minor
  uint32_t* ws_mask;       // [grid][T (Ld-1) + (Lc-1)][2][512]
  uint32_t* ovf;           // incremented by a CTA that clamped an fp16 operand
  long long* dbg;          // cycle counters of CTA 0 (builds with -DGMPC_T128_TIMED only)
};

#ifdef GMPC_T128_TIMED
// A wait that gives up after 20 M cycles and records (tag, warp, layer) instead of hanging: the development build.
__device__ __forceinline__ void t_wait_dbg(uint32_t bar, uint32_t parity, int tag, int layer, long long* dbg) {
  if (mbar_try_wait_a(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait_a(bar, parity)) {
    if (dbg != nullptr && *((volatile long long*)(dbg + 701)) > 40) return;   // the run is lost: let everybody fall through
    if (clock64() - t0 > 20000000LL) {
      if (dbg != nullptr) {
        const unsigned long long i = atomicAdd((unsigned long long*)(dbg + 700), 1ULL);
        if (i < 96) dbg[704 + i] = ((long long)tag << 40) | ((long long)(threadIdx.x >> 5) << 32) | (unsigned)layer;
        if (i == 0) for (int w = 0; w < 24; ++w) dbg[840 + w] = ((volatile long long*)dbg)[800 + w];   // snapshot of the positions
        atomicAdd((unsigned long long*)(dbg + 701), 1ULL);
      }
      return;
    }
  }
}
#define T_WAIT(bar, par, tag, layer) t_wait_dbg(bar, par, tag, layer, P.dbg)
// last position of every warp of CTA 0 (layer << 8 | code), for the post-mortem of a lost run
#define T128_POS(layer, code) do { if (blockIdx.x == 0 && P.dbg != nullptr && (threadIdx.x & 31) == 0) \
    ((volatile long long*)P.dbg)[800 + (threadIdx.x >> 5)] = ((long long)(layer) << 8) | (code); } while (0)
#define T128_E0(v) v = clock64()
#define T128_E1(acc, v) acc += clock64() - v
// event trace of four consecutive layers of CTA 0 (T128_TRACE_L0 ..): cycle stamps into dbg[256 ..)
#ifndef T128_TRACE_L0
#define T128_TRACE_L0 545
#endif
#define T128_TR(layer, slot) do { if (blockIdx.x == 0 && P.dbg != nullptr && (layer) >= T128_TRACE_L0 && (layer) < T128_TRACE_L0 + 4) \
    P.dbg[(slot)] = clock64(); } while (0)
#else
#define T128_TR(layer, slot)
#define T_WAIT(bar, par, tag, layer) mbar_wait_a(bar, par)
#define T128_POS(layer, code)
#define T128_E0(v)
#define T128_E1(acc, v)
#endif

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
  s.total = s.bars + 8 * (2 * T_MAX_SLOTS + T_MAXP + T_MAXKS + 1) + 16;
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
__device__ __forceinline__ void t_st8(uint32_t taddr, const uint32_t* r) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(r[0]),
               "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]) : "memory");
}
__device__ __forceinline__ void t_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ uint32_t t_region(uint32_t c) { return c ? T_R1 : T_R0; }
// 0xFFFF in every 16-bit half of x whose top bit is set (prmt with sign replication; __byte_perm drops that bit)
__device__ __forceinline__ uint32_t t_half_signs(uint32_t x) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xBB99u));
  return d;
}

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
  uint64_t* acc_bar = empty_bar + T_MAX_SLOTS;   // all MMAs of the layer are complete (one commit per issuer)
  uint64_t* rnd_bar = acc_bar + T_MAXP;          // [4] feature blocks [4 r, 4 r + 4) of the next operand are stored; ALL 16
                                                 // epilogue warps arrive, with or without a block in the round
  uint64_t* rc_bar = rnd_bar + T_MAXKS;          // sub 1 has read the u columns of the narrow accumulator (4 warp arrivals)
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(rc_bar + 1);

  const int tid = threadIdx.x, lane = tid & 31;
  // (a shuffle from lane 0 is provably warp-uniform: the role branches below are then convergent for the compiler,
  // which lets the MMA issuer keep its loop state and descriptors in uniform registers)
  const int warp = __shfl_sync(0xFFFFFFFFu, tid >> 5, 0);
  const int n = P.n, m = P.m, T = P.T, NS = P.nslot;
  const int my_tiles = ((int)blockIdx.x < P.ntiles) ? (P.ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;

  for (int i = tid; i < P.nbias; i += T_THREADS) bias_s[i] = P.bias[i];
  for (int i = tid; i < P.nscale; i += T_THREADS) inv_s[i] = P.inv_scale[i];
  if (tid == 0) {
    for (int s = 0; s < NS; ++s) {
      mbar_init(&full_bar[s], 1);
    }
    mbar_init(acc_bar, T_MAXP);
    for (int j = 0; j < 4; ++j) mbar_init(&rnd_bar[j], T_EPI_WARPS);
    mbar_init(rc_bar, 4);
    for (int s = 0; s < NS; ++s) mbar_init(&empty_bar[s], T_MAXP);
    mbar_fence_init();
  }
  __syncwarp();
  if (warp == 0) tmem_alloc(tmem_holder, 512);
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_holder;

  if (warp == T_EPI_WARPS) {
    // ================================================================== weight-stream producer
    // One ring slot = one group = kpg k-steps of one layer (all N-parts, [hi tile | lo tile] each), one bulk
    // copy.  tools/bulk_copy_rate.cu: a cp.async.bulk holds its issuing lane ~460 cycles, so four lanes take
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
    __syncwarp();
  } else if (warp > T_EPI_WARPS) {
    // ================================================================== MMA issuers (one thread each)
    // Issuer p issues N-part p of every layer, round by round as the feature blocks of the operand arrive.
    // Accumulator columns of different parts are disjoint, a part's completion is tracked by the commit of the
    // lane that issued it, so several instruction streams may interleave freely on the tensor pipe.  Every
    // issuer observes every phase of every round barrier, also in layers in which it issues nothing (a waiter
    // that only flipped its parity bit could run two phases ahead and pass a parity wait on a stale phase).
    const int which = warp - T_EPI_WARPS - 1;
    // The whole warp walks the schedule (converged); one elected lane issues the MMAs and the commits.
    const uint32_t tmem_u = __shfl_sync(0xFFFFFFFFu, tmem_base, 0);
    {
      const uint32_t full_a = smem_u32(full_bar), empty_a = smem_u32(empty_bar), ring_a = smem_u32(tsm + L.ring);
      const uint32_t acc_a = smem_u32(acc_bar), rnd_a = smem_u32(rnd_bar), rc_a = smem_u32(rc_bar);
      const uint32_t desc_hi = (uint32_t)(umma_smem_desc(0, 0, 128) >> 32);
      const uint32_t slot16 = P.slot_bytes >> 4, ring16 = ring_a >> 4;
      uint32_t slot = 0, ph = 0, rpar = 0, rc_par = 0, cur = 0;
      int lno = -1;
#ifdef GMPC_T128_TIMED
      long long i_kst = 0, i_full = 0, i_first = 0, iq;
      long long il_first[4] = {0, 0, 0, 0}, il_rest[4] = {0, 0, 0, 0}, il_t0 = 0;   // dyn fwd, by layer
      const long long i_begin = clock64();
#endif
      for (int ti = 0; ti < my_tiles; ++ti) {
        TPassWalk walk;
        for (;;) {
          const int kind = walk.next(P);
          if (kind == DIR_END) break;
          const TDir& D = P.dir[kind];
          for (int l = 0; l < D.L; ++l) {
            const TLayer& Y = D.layer[l];
            const int nks = Y.nks, kpg = Y.kpg;
            const bool narrow = (l == D.L - 1);
            ++lno;
#ifdef GMPC_T128_TIMED
            il_t0 = clock64();
#endif
            const uint32_t a_t = tmem_u + t_region(cur);
            const uint32_t d_t = tmem_u + (narrow ? T_RC : t_region(cur ^ 1u));
            cur ^= 1u;
            const bool mine = which < Y.np;
            // the u columns of the previous adjoint boundary's narrow accumulator have been read (sub 1)
            if (which == 0 && narrow && kind == DIR_DYN_B) {
              T_WAIT(rc_a, rc_par, 2, lno);
              rc_par ^= 1;
            }
            const int ncol = mine ? Y.ncol[which] : 16;
            const uint32_t idesc = t_idesc(ncol);
            const uint32_t tile16 = ((uint32_t)ncol * 32u) >> 4;          // one (hi or lo) tile, in 16-byte units
            const uint32_t lbo_f = (((uint32_t)ncol * 16u) >> 4) << 16;   // LBO field of the descriptor
            const uint32_t kst16 = ((uint32_t)Y.npad * 64u) >> 4;         // one k-step of the layer (all parts)
            const uint32_t part16 = mine ? ((uint32_t)Y.c0[which] * 64u) >> 4 : 0u;
            const uint32_t d_p = d_t + (mine ? (uint32_t)Y.c0[which] : 0u);
            const int extra = which == 0 ? T_MAXP - Y.np : 0;   // issuer 0 releases the slot for the absent parts too
            uint32_t a_off = 0;           // 16 j: TMEM column offset of k-step j of the operand
            uint32_t b_lo = 0;
            int in_group = 0;             // k-steps left in the ring group in flight
#pragma unroll 1
            for (int j = 0; j < nks; ++j) {
              if ((j & 3) == 0) {         // feature blocks [j, j + 4) of the operand are stored
                T128_E0(iq);
                T128_POS(lno, 64 + j);
                T_WAIT(rnd_a + 2 * j, (rpar >> (j >> 2)) & 1u, 4 + 10 * which, lno * 16 + j);
                T128_TR(lno, 256 + (lno - T128_TRACE_L0) * 64 + which * 16 + j);
#ifdef GMPC_T128_TIMED
                if (j == 0) {
                  i_first += clock64() - iq;
                  if (kind == DIR_DYN_F && l < 4) { il_first[l] += clock64() - il_t0; il_t0 = clock64(); }
                } else i_kst += clock64() - iq;
#endif
                tc_fence_after();
              }
              if (in_group == 0) {
                in_group = nks - j < kpg ? nks - j : kpg;
                T128_E0(iq);
                if (mine) T_WAIT(full_a + slot * 8, ph, 3 + 10 * which, lno);
                T128_E1(i_full, iq);
                b_lo = (ring16 + slot * slot16 + part16) | lbo_f;
              }
              if (mine) {
                const uint64_t bh = ((uint64_t)desc_hi << 32) | b_lo;
                const uint64_t bl = ((uint64_t)desc_hi << 32) | (b_lo + tile16);
                if (elect_one()) {
                  t_mma(d_p, a_t + a_off, bh, idesc, a_off);          // accumulate = (j > 0)
                  t_mma(d_p, a_t + a_off + 8u, bh, idesc, 1u);
                  t_mma(d_p, a_t + a_off, bl, idesc, 1u);
                }
              }
              a_off += 16;
              b_lo += kst16;
              if (--in_group == 0) {
                if (mine && elect_one()) {
                  umma_commit_a(empty_a + slot * 8);
                  for (int x = 0; x < extra; ++x) umma_commit_a(empty_a + slot * 8);
                }
                if (++slot == (uint32_t)NS) { slot = 0; ph ^= 1; }
              }
            }
            rpar ^= (1u << ((nks + 3) >> 2)) - 1u;
            // Every issuer commits, also with nothing issued: the epilogue passes a layer only after EVERY issuer has
            // walked it, and every epilogue warp arrives on the round barriers of every layer, so no barrier can get
            // two phases ahead of any of its waiters.
            if (elect_one()) umma_commit_a(acc_a);
#ifdef GMPC_T128_TIMED
            if (kind == DIR_DYN_F && l < 4) il_rest[l] += clock64() - il_t0;
#endif
          }
        }
      }
#ifdef GMPC_T128_TIMED
      if (blockIdx.x == 0 && P.dbg != nullptr) {
        long long* o = P.dbg + 40 + which * 8;
        o[0] = clock64() - i_begin; o[1] = i_first; o[2] = i_kst; o[3] = i_full;
        if (which == 0) for (int i = 0; i < 4; ++i) { P.dbg[100 + i] = il_first[i]; P.dbg[104 + i] = il_rest[i]; }
      }
#endif
    }
    __syncwarp();
  } else {
    // ================================================================== epilogue / per-trajectory work
    const int ew = warp;                   // 0..15
    const int q = warp & 3;                // TMEM lane quarter this warp may access
    const int sub = ew >> 2;               // feature blocks j % 4 == sub; sub 0 also owns the trajectory's state,
                                           // sub 1 its action update
    const int r = q * 32 + lane;           // trajectory (TMEM lane) in the tile
    const int et = ew * 32 + lane;         // 0..511
    const uint32_t tl = tmem_base + ((uint32_t)(q * 32) << 16);
    const uint32_t acc_a = smem_u32(acc_bar), rnd_a = smem_u32(rnd_bar), rc_a = smem_u32(rc_bar);
    uint32_t acc_par = 0, cur = 0;         // cur: region of the operand of the layer in flight
    int elno = 0;
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
    constexpr int MSTR = 2 * T_EPI;  // mask words per (step, layer): 4 feature blocks x 16 bits per thread
    uint32_t* wsMask = P.ws_mask + (size_t)blockIdx.x * ((size_t)T * (Ld - 1) + (Lc - 1)) * MSTR + et;
    uint32_t* costMask = wsMask + (size_t)T * (Ld - 1) * MSTR;
    const bool adam = (P.mode == MODE_PLAN && P.method == 1);
    const bool need_goal = cost_mode || P.mode == MODE_L2GRAD;
    const int nk_dynf = P.dir[DIR_DYN_F].layer[0].nks, nk_dynb = P.dir[DIR_DYN_B].layer[0].nks;
    const int nk_costf = P.use_cost ? P.dir[DIR_COST_F].layer[0].nks : 0;
    const int nk_costb = P.use_cost ? P.dir[DIR_COST_B].layer[0].nks : 0;
    float opmax = 0.f;  // largest operand magnitude this thread has written (fp16 range check)
    __half2 opmax2 = __float2half2_rn(0.f);   // the same for the packed hidden operands (NaN-propagating maximum)

#ifdef GMPC_T128_TIMED
    long long e_acc = 0, e_hid = 0, e_bnd = 0, e_ld = 0, e_alu = 0, e_st = 0, e_fin = 0, eq, eh, ex;
    const long long e_begin = clock64();
#endif
    // all MMAs of the layer are complete (every issuer has committed, with or without a part in this layer)
    auto wait_acc = [&]() {
      T128_POS(elno, 1);
      T128_E0(eq);
      T_WAIT(acc_a, acc_par, 20, elno);
      T128_E1(e_acc, eq);
      T128_POS(elno, 2);
      if (q == 0 && lane == 0) T128_TR(elno, 520 + (elno - T128_TRACE_L0) * 20 + sub * 5);
      ++elno;
      acc_par ^= 1u;
      tc_fence_after();
    };
    // "my share of feature blocks [4 r, 4 r + 4) of the next operand is stored" (after tcgen05.wait::st), or "I have
    // none in this round": every warp arrives on every round barrier of every layer
    auto arrive_round = [&](int rnd) {
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_a(rnd_a + 8 * rnd);
    };
    // sub 0: features [16 ks, 16 ks + 16) of a small operand, already scaled -> feature block ks of region `opn`
    auto store_kstep = [&](uint32_t opn, int ks, const float (&v)[16]) {
      uint32_t hi[8], lo[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        opmax = fmaxf(opmax, fmaxf(fabsf(v[2 * k]), fabsf(v[2 * k + 1])));
        split_h2(v[2 * k], v[2 * k + 1], hi[k], lo[k]);
      }
      t_st8(opn + 16 * ks, hi);
      t_st8(opn + 16 * ks + 8, lo);
    };
    auto load16 = [&](int half, float (&o)[16]) {   // 16 columns of the narrow accumulator
      uint32_t d0[16];
      tmem_ld16_issue(tl + T_RC + 16 * half, d0);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 16; ++c) o[c] = __uint_as_float(d0[c]);
    };

    // the first adjoint boundary layer has no predecessor whose narrow accumulator is still being read
    if (sub == 1 && lane == 0) mbar_arrive_a(rc_a);
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
      // staging cost of step t for this thread's trajectory (sub 0): x from shared memory, u / goal from scratch.
      // Only the sweep whose J is reported needs it (the final evaluation, or the only sweep of a call without one).
      bool want_J = false;
      auto stage_cost = [&](int t) {
        if (!want_J) return;
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
      // sub 0: the <= 32-feature operand [xs ; us] * sc (us: nu features after the n state features) -> feature
      // blocks [0, nk) of region `opn`
      auto publish_state_operand = [&](uint32_t opn, const float* xcol, int nu, float sc, int nk) {
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
            store_kstep(opn, h, v);
          }
        }
        t_st_wait();
      };

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
            publish_state_operand(tl + t_region(cur), xs, m, s0, nk_dynf);
          }
          arrive_round(0);
          Jr = 0.f;
          want_J = last || !P.final_fwd;
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
          // accumulator -> (+bias, ReLU, mask bit) or (mask gate) -> hi/lo -> A operand of the next layer, in place.
          // The forward pass runs in scaled units a' = a s (ReLU is positively homogeneous): the bias enters as b s.
          const uint32_t dreg = tl + t_region(cur ^ 1u);
          // (the tensor core truncates every accumulation towards zero: 3 nks events per output shrink it by
          // ~acc_comp each, undone here on average -- see T128State::acc_comp)
          const float inv = inv_s[Y.scale_idx] * fmaf(P.acc_comp, (float)(3 * Y.nks), 1.f);
          const float* bp = bias_s + Y.bias_off;
          float s = 1.f;
          // ReLU bits of my feature blocks: block j -> 16 bits of word (j >> 3), at bits [0, 8) (even elements) and
          // [16, 24) (odd elements), shifted left by 8 when (j & 4); a set bit = pre-activation >= +0
          uint32_t mw0 = 0u, mw1 = 0u;
          if (!fwd) { mw0 = maskp[0]; mw1 = maskp[T_EPI]; }
          // 16 accumulator values (one feature block) of this thread's trajectory -> packed hi / lo halfs, two elements
          // at a time in the packed domain: hr = fp16x2(z); negative halves are cleared with a byte-permute sign mask
          // (forward: ReLU; the adjoint clears the halves whose stored bit is 0); lo = fp16x2(z - hi) through the
          // mixed-precision fma (fp16 x fp16 + fp32), cleared with the same mask.
          // gbits: forward: returns the NEGATIVE bits of the block; adjoint: takes the block's stored bits.
          auto process16 = [&](const uint32_t (&raw)[16], int jg, uint32_t& gbits, uint32_t (&hi)[8], uint32_t (&lo)[8]) {
            uint32_t neg = 0u;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              float z0, z1;
              if (fwd) {
                const float2 b2 = *reinterpret_cast<const float2*>(bp + 16 * jg + 2 * k);
                z0 = fmaf(__uint_as_float(raw[2 * k]), inv, b2.x * s);
                z1 = fmaf(__uint_as_float(raw[2 * k + 1]), inv, b2.y * s);
              } else {
                z0 = __uint_as_float(raw[2 * k]) * inv;
                z1 = __uint_as_float(raw[2 * k + 1]) * inv;
              }
              uint32_t hr, clr;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(hr) : "f"(z1), "f"(z0));
              if (fwd) {
                clr = t_half_signs(hr);                              // 0xFFFF in every negative half
                neg |= clr & (0x00010001u << k);
              } else {
                clr = ~t_half_signs(gbits << (15 - k));              // 0xFFFF in every half whose stored bit is 0
              }
              const uint32_t h = hr & ~clr;
              float d0, d1;
              asm("{\n\t.reg .f16 l, u, m1;\n\tmov.b32 {l, u}, %2;\n\tmov.b16 m1, 0xBC00;\n\t"
                  "fma.rn.f32.f16 %0, l, m1, %3;\n\tfma.rn.f32.f16 %1, u, m1, %4;\n\t}"
                  : "=f"(d0), "=f"(d1) : "r"(h), "f"(z0), "f"(z1));
              uint32_t lr;
              asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(lr) : "f"(d1), "f"(d0));
              hi[k] = h;
              lo[k] = lr & ~clr;
              const __half2 h2 = *reinterpret_cast<const __half2*>(&h);
              opmax2 = __hmax2_nan(opmax2, fwd ? h2 : __habs2(h2));
            }
            if (fwd) gbits = neg;
          };
          T128_E0(eh);
          wait_acc();
          if (fwd) s = ssc_s[r];
          const int nko = Y.npad >> 4;
          // one round: block j (if it exists) -> operand, then the round's arrival.  (Reading the block of the next
          // round ahead of the arithmetic needs 16 more registers than the 96 a 5-warp scheduler partition allows.)
#pragma unroll 1
          for (int j = 3 - sub; j < ((nko + 3) & ~3); j += 4) {   // the first blocks come from the subs without boundary duties
            if (j < nko) {
              uint32_t raw[16], hi[8], lo[8];
              T128_E0(ex);
              tmem_ld16_issue(dreg + 16 * j, raw);
              tmem_ld_wait();
              T128_E1(e_ld, ex);
              T128_E0(ex);
              const int sh = (j & 4) << 1;            // 0 or 8: position of block j's bits in its mask word
              uint32_t gb = 0u;
              if (!fwd) gb = ((j & 8) ? mw1 : mw0) >> sh;
              process16(raw, j, gb, hi, lo);
              if (fwd) {
                const uint32_t act = (~gb & 0x00FF00FFu) << sh;   // negative bits -> (z >= +0) bits
                if (j & 8) mw1 |= act; else mw0 |= act;
              }
              T128_E1(e_alu, ex);
              T128_E0(ex);
              t_st8(dreg + 16 * j, hi);
              t_st8(dreg + 16 * j + 8, lo);
              t_st_wait();
              T128_E1(e_st, ex);
            }
            arrive_round(j >> 2);
            T128_POS(elno, 32 + j);
            if (q == 0 && lane == 0) T128_TR(elno - 1, 520 + (elno - 1 - T128_TRACE_L0) * 20 + sub * 5 + 1 + (j >> 2));
          }
          T128_E0(ex);
          if (fwd) { maskp[0] = mw0; maskp[T_EPI] = mw1; }
          cur ^= 1u;
          T128_E1(e_fin, ex);
          T128_E1(e_hid, eh);
        }
        T128_E0(eh);
        // -------------------------------------------------------------- last layer of the pass (<= 32 outputs)
        // accumulator in the narrow region; the next operand (if a layer follows) goes to the region the
        // layer's own operand does NOT occupy
        const TLayer& Yf = D.layer[D.L - 1];
        const float invf = inv_s[Yf.scale_idx] * fmaf(P.acc_comp, (float)(3 * Yf.nks), 1.f);
        const uint32_t opn = tl + t_region(cur ^ 1u);
        cur ^= 1u;
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
            if (more) {
              const float sn = pow2_scale_to_8(mx);
              ssc_s[r] = sn;
              publish_state_operand(opn, xs, with_u ? m : 0, sn, nk_next);
            }
          }
          if (more) arrive_round(0);
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
                  publish_state_operand(opn, lam, 0, sg, nk_dynb);
                }
                arrive_round(0);
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
              store_kstep(opn, 0, y0);
              if (nk_costb > 1) store_kstep(opn, 1, y1);
              t_st_wait();
            }
          }
          if (!last) arrive_round(0);
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
            sg = pow2_scale_to_8(mx);
            sig_s[((T - 1) & 1) * T_NB + r] = sg;
            publish_state_operand(opn, lam, 0, sg, nk_dynb);
          }
          arrive_round(0);
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
            if (t > 0) {
              sg = pow2_scale_to_8(mx);
              sig_s[((t - 1) & 1) * T_NB + r] = sg;
              publish_state_operand(opn, lam, 0, sg, nk_next);
              arrive_round(0);
            }
            if (P.lam_out != nullptr && rvalid)
              for (int i = 0; i < n; ++i) P.lam_out[(qr * (T + 1) + t) * n + i] = lam[i * T_NB];
          } else if (sub == 1) {
            const float c0 = invf * pow2_recip(sig_s[(t & 1) * T_NB + r]);
            // the u features are columns [n, n + m) of the narrow accumulator: read them, then the update is off the
            // critical path
            float o0[16], o1[16];
            const bool need0 = n < 16, need1 = n + m > 16;
            if (need0) load16(0, o0);
            if (need1) load16(1, o1);
            tc_fence_before();     // the narrow accumulator may be overwritten by the next adjoint boundary layer
            __syncwarp();
            if (lane == 0) mbar_arrive_a(rc_a);
            if (t > 0) arrive_round(0);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
              if (h == 0 ? need0 : need1) {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                  const int i = 16 * h + c;
                  if (i >= n && i < n + m) {
                    const int j = i - n;
                    const size_t ix = ((size_t)t * m + j) * T_NB;
                    float u = wsU[ix], m_old = 0.f, v_old = 0.f;
                    if (adam) { m_old = wsM[ix]; v_old = wsV[ix]; }
                    float g = (h == 0 ? o0[c] : o1[c]) * c0;
                    if (cost_mode) g = (w0 * u) / su + g;
                    if (P.mode == MODE_PLAN) {
                      if (P.method == 0) {
                        u = u - P.lr * g;
                      } else {
                        const float mo = P.b1 * m_old + (1.f - P.b1) * g;
                        const float ve = P.b2 * v_old + (1.f - P.b2) * g * g;
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
              }
            }
          } else {
            if (t > 0) arrive_round(0);
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
    opmax = fmaxf(opmax, 0.f);
    if (!(fmaxf(__low2float(opmax2), __high2float(opmax2)) <= 65000.f)) opmax = __int_as_float(0x7fc00000);
    if (!(opmax <= 65000.f) && P.ovf != nullptr) atomicAdd(P.ovf, 1u);
#ifdef GMPC_T128_TIMED
    if (blockIdx.x == 0 && P.dbg != nullptr && lane == 0 && (ew & 3) == 0) {
      long long* o = P.dbg + 8 + sub * 8;   // the four subs of lane quarter 0
      o[0] = clock64() - e_begin; o[1] = e_acc; o[2] = e_hid; o[3] = e_bnd; o[4] = e_ld; o[5] = e_alu; o[6] = e_st; o[7] = e_fin;
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
// Image of a layer: per k-step, per N-part p (operand rows [c0[p], c0[p+1])): [hi tile | lo tile],
// tile = [2 k-chunks][rows of the part][8 halfs]
// (K-major SWIZZLE_NONE core matrices: LBO = rows * 16, SBO = 128; pinned by tools/t128_probe.cu).
struct TParts {
  int np;
  int c0[T_MAXP + 1];   // first row of part p; c0[np] = padded row count
};
__global__ void t128_pack_kernel(const float* __restrict__ W, int K, int N, int transposed, uint8_t* dst, TParts pp,
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
  int part = 0;
  while (part + 1 < pp.np && row >= pp.c0[part + 1]) ++part;
  const int rows = pp.c0[part + 1] - pp.c0[part];
  const int npad = pp.c0[pp.np];
  row -= pp.c0[part];
  const size_t tile_b = (size_t)rows * 32;
  const int j = kk >> 4, k16 = kk & 15;
  (void)nks;
  uint8_t* p = dst + ((size_t)j * npad + pp.c0[part]) * 64 + (size_t)(k16 >> 3) * rows * 16 + (row >> 3) * 128 + (row & 7) * 16 +
               (k16 & 7) * 2;
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
  int nbias = 0, nscale = 0, nslot = 0;
  uint32_t slot_bytes = 0;
  size_t smem_bytes = 0;
  long long* d_dbg = nullptr;  // tools/t128_bench.cu with -DGMPC_T128_TIMED
  // The fp32 accumulation of tcgen05.mma truncates (rounds towards zero), so every accumulate event shrinks the
  // running sum by half an ulp on average; with the three split products in ONE accumulator a 208-wide layer has 39
  // events per output, which shows as a uniform relative shrink of the rollout (measured at C2 dims: 1.9e-5 on X
  // after 32 steps against 6.4e-6 for the 32-trajectory kernel's 13 events and 1.9e-7 for fp32 FMA).  The epilogue
  // multiplies the accumulator by 1 + acc_comp * events; acc_comp is calibrated on the hardware
  // (tools/acc_comp_calib.py, profiles/r2_acc_comp_calib.txt), GMPC_T128_ACC_COMP overrides it (0 disables).
  float acc_comp = 0.30f * 5.9604644775390625e-08f;   // 0.30 x 2^-24 per accumulate event (measured)
};

inline int t_rup(int v, int a) { return (v + a - 1) / a * a; }

inline void t128_destroy(T128State& S) {
  cudaFree(S.d_stream); cudaFree(S.d_bias); cudaFree(S.d_scale); cudaFree(S.d_absmax); cudaFree(S.d_ovf);
  cudaFree(S.ws_X); cudaFree(S.ws_G); cudaFree(S.ws_U); cudaFree(S.ws_M); cudaFree(S.ws_V); cudaFree(S.ws_mask);
  S.d_stream = nullptr; S.d_bias = S.d_scale = nullptr; S.d_absmax = S.d_ovf = nullptr;
  S.ws_X = S.ws_G = S.ws_U = S.ws_M = S.ws_V = nullptr; S.ws_mask = nullptr;
  S.supported = false;
}

// N-parts of a layer with npad output columns: parts of `bpp` feature blocks (16 columns each), the last part
// takes the remainder; a layer of fewer than bpp blocks (and the narrow last layer of a pass) is one part.
inline void t128_parts(TLayer& Y, int bpp) {
  const int nko = Y.npad / 16;
  const int np = std::min(T_MAXP, std::max(1, nko / bpp));
  const int per = (nko / bpp > T_MAXP) ? nko / np : bpp;
  Y.np = np;
  int c = 0;
  for (int p = 0; p < T_MAXP; ++p) {
    const int nb = p >= np ? 0 : (p == np - 1 ? nko - per * (np - 1) : per);
    Y.c0[p] = c;
    Y.ncol[p] = 16 * nb;
    c += 16 * nb;
  }
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
  int hmin = 1 << 30;
  for (int i = 1; i < S.Ld; ++i) hmin = std::min(hmin, dyn_dims[i]);
  for (int i = 1; i < S.Lc; ++i) hmin = std::min(hmin, cost_dims[i]);
  if (hmin < 17) { S.why = "hidden width < 17 (every epilogue sub-group pair must own a feature block)"; return GMPC_OK; }
  if (hmax > T_MAXH) { S.why = "hidden width > 240 (two operand / accumulator regions exceed the 512 TMEM columns)"; return GMPC_OK; }
  if (c.n + c.m > 32 || c.cost_fout > 32) { S.why = "n+m or fout > 32"; return GMPC_OK; }
  int bpp = 4;
  if (const char* e = getenv("GMPC_T128_PART_BLOCKS")) bpp = std::max(1, atoi(e));
  if (const char* e = getenv("GMPC_T128_ACC_COMP")) S.acc_comp = (float)(atof(e) * 5.9604644775390625e-08);   // in units of 2^-24
  int nbias = 0, nscale = 0;
  uint32_t slot = 0;
  auto one_layer = [&](TLayer& Y, int M_true, int red, bool narrow) {
    memset(&Y, 0, sizeof(Y));
    Y.M_true = M_true; Y.npad = t_rup(M_true, 16); Y.nks = t_rup(red, 16) / 16;
    t128_parts(Y, narrow ? 1000 : bpp);
    slot = std::max(slot, (uint32_t)Y.npad * 64u);
  };
  auto geom = [&](const int* dims, int Ln, TDir& F, TDir& Bw) {
    F.L = Bw.L = Ln; F.pad_ = Bw.pad_ = 0;
    for (int l = 0; l < Ln; ++l) {
      TLayer& Y = F.layer[l];
      one_layer(Y, dims[l + 1], dims[l], l == Ln - 1);
      Y.bias_off = nbias; nbias += Y.npad;
      Y.scale_idx = nscale + l;
    }
    for (int i = 0; i < Ln; ++i) {
      const int lt = Ln - 1 - i;  // the adjoint pass visits the transposed layers L-1 .. 0
      TLayer& Y = Bw.layer[i];
      one_layer(Y, dims[lt], dims[lt + 1], i == Ln - 1);
      Y.bias_off = 0; Y.scale_idx = nscale + lt;
    }
    nscale += Ln;
  };
  geom(S.dyn_dims, S.Ld, S.dir[DIR_DYN_F], S.dir[DIR_DYN_B]);
  geom(S.cost_dims, S.Lc, S.dir[DIR_COST_F], S.dir[DIR_COST_B]);
  S.slot_bytes = std::max(slot, 16384u);  // ring group: >= one k-step of the widest layer, 16 KB when it is smaller
  size_t sz[4] = {0, 0, 0, 0};
  for (int d = 0; d < 4; ++d)
    for (int l = 0; l < S.dir[d].L; ++l) {
      TLayer& Y = S.dir[d].layer[l];
      Y.kpg = std::max(1, (int)(S.slot_bytes / (Y.npad * 64)));
      Y.goff = (uint32_t)sz[d];
      sz[d] += (size_t)Y.nks * Y.npad * 64;
    }
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
  // (the attribute belongs to the kernel, not to the handle: the device maximum, so that handles of different
  // shapes can be alive at the same time)
  if (cudaFuncSetAttribute(plan_t128_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_optin) != cudaSuccess)
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
  auto parts_of = [](const TLayer& Y) {
    TParts pp;
    pp.np = Y.np;
    for (int p = 0; p < T_MAXP; ++p) pp.c0[p] = Y.c0[p];
    for (int p = Y.np; p <= T_MAXP; ++p) pp.c0[p] = Y.npad;   // c0[np] = padded row count
    return pp;
  };
  auto one = [&](const int* dims, int Ln, const float* const* W, const float* const* b, TDir& F, TDir& Bw, int sbase) {
    for (int l = 0; l < Ln; ++l) {
      const int K = dims[l], N = dims[l + 1], blocks = (K * N + 255) / 256;
      const TLayer& f = F.layer[l];
      const TLayer& rv = Bw.layer[Ln - 1 - l];
      h16_absmax_kernel<<<std::min(blocks, 64), 256, 0, st>>>(W[l], K * N, S.d_absmax + sbase + l);
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 0, const_cast<uint8_t*>(F.gsrc) + f.goff, parts_of(f), f.nks,
                                               S.d_absmax + sbase + l, S.d_scale + sbase + l);
      t128_pack_kernel<<<blocks, 256, 0, st>>>(W[l], K, N, 1, const_cast<uint8_t*>(Bw.gsrc) + rv.goff, parts_of(rv),
                                               rv.nks, S.d_absmax + sbase + l, nullptr);
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
  Q.acc_comp = S.acc_comp;
  Q.x0 = P.x0; Q.U_in = P.U_in; Q.goal = P.goal; Q.mpcw = P.mpcw;
  Q.bias = S.d_bias; Q.inv_scale = S.d_scale;
  Q.U_out = P.U_out; Q.X_out = P.X_out; Q.J_out = P.J_out; Q.dU_out = P.dU_out; Q.lam_out = P.lam_out;
  Q.ws_X = S.ws_X; Q.ws_G = S.ws_G; Q.ws_U = S.ws_U; Q.ws_M = S.ws_M; Q.ws_V = S.ws_V;
  Q.ws_mask = S.ws_mask;
  Q.ovf = S.d_ovf;
  Q.dbg = S.d_dbg;
  if (Q.ntiles <= 0) return GMPC_OK;
  const int grid = std::min(Q.ntiles, S.num_sms);
  plan_t128_kernel<<<grid, T_THREADS, S.smem_bytes, st>>>(Q);
  ++*launches;
  return cudaGetLastError() == cudaSuccess ? GMPC_OK : GMPC_E_CUDA;
}

}  // namespace gmpc
