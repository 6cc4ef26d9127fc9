// critic.cuh -- JS-GAN critic (LSTM discriminator) forward / BCE / backward, flat-parameter Adam.
//
// Restates critic/nn.py:27-42 (scan of OptimizedLSTMCell from zero carry, final h, Dense stack,
// Dense(1)), gan/js_policy.py:41-58 (BCE with +-1 labels, batch mean, gradient) and
// norm/runner.py:53-56 (clip_by_global_norm(100) then adam).  One CTA per sample (looping),
// thread j owns gate column j; per-CTA partial gradients are reduced in a fixed order so the
// result is deterministic.  The step is latency-bound (21 057 parameters at n=17, F=64).
#pragma once
#include "common.cuh"

namespace gmpc {

struct CriticDims {
  int n, F, L, H, T1;
  long long P;
  long long oWi, oWh, obh, oDk[MAXL], oDb[MAXL], oWo, obo;
  int din[MAXL];  // input width of hidden Dense l
  int dlast;      // input width of the final Dense(1)
};

__device__ __forceinline__ float sigmoidf_(float z) { return 1.f / (1.f + expf(-z)); }
__device__ __forceinline__ float softplusf_(float x) {
  return fmaxf(x, 0.f) + log1pf(expf(-fabsf(x)));
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// xseq [*, T1, n]; sample s of the minibatch is row perm[s] (or s when perm == nullptr).
// losses[s] = unscaled per-sample loss; logits[s] (nullable) = raw score.
// When do_bwd: partial[blockIdx.x][P] accumulates inv_count * d loss_s / d params.
// When dx_out != nullptr (input-gradient mode, used by the generator loss gan/js_policy.py:60-68):
// the backward pass is seeded with d score = 1, no weight gradients are accumulated, and
// dx_out[s][t][i] = d score_s / d xseq[s][t][i].
__global__ void critic_kernel(const CriticDims D, const float* __restrict__ xseq,
                              const float* __restrict__ label, const int* __restrict__ perm,
                              const float* __restrict__ prm, float inv_count, long long Bc,
                              float* losses, float* logits, float* partial, int do_bwd,
                              float* dx_out = nullptr) {
  extern __shared__ __align__(16) float sm[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int n = D.n, F = D.F, G = 4 * D.F, T1 = D.T1;
  const int W = max(D.F, D.H);
  float* xs = sm;                      // [T1][n]
  float* gates = xs + T1 * n;          // [T1][4F]  (activations, then dz in the backward)
  float* cs = gates + T1 * G;          // [T1+1][F]
  float* hs = cs + (T1 + 1) * F;       // [T1+1][F]
  float* dh = hs + (T1 + 1) * F;       // [F]
  float* dc = dh + F;                  // [F]
  float* hact = dc + F;                // [L][W]   head activations (hact[0] = final h)
  float* hdz = hact + D.L * W;         // [W]      head adjoint scratch
  float* hda = hdz + W;                // [W]
  float* scal = hda + W;               // [4]
  float* part = partial + (size_t)blockIdx.x * D.P;
  const float* Wi = prm + D.oWi;
  const float* Wh = prm + D.oWh;
  const float* bh = prm + D.obh;

  // every parameter receives a contribution from every sample, so the CTA's first sample stores
  // and later ones accumulate: no zero-fill pass and no read of the partial buffer for it
  bool first = true;
  const bool xgrad = dx_out != nullptr;
  auto acc_part = [&](long long idx, float v) {
    if (!xgrad) part[idx] = first ? v : part[idx] + v;
  };

  for (long long s = blockIdx.x; s < Bc; s += gridDim.x, first = false) {
    const long long src = perm ? (long long)perm[s] : s;
    __syncthreads();
    for (int i = tid; i < T1 * n; i += blockDim.x) xs[i] = xseq[src * T1 * n + i];
    for (int i = tid; i < F; i += blockDim.x) { cs[i] = 0.f; hs[i] = 0.f; }
    __syncthreads();
    // ---------------------------------------------------------------- LSTM forward
    for (int t = 0; t < T1; ++t) {
      if (tid < G) {
        float z = bh[tid];
        const float* x = xs + t * n;
        for (int i = 0; i < n; ++i) z = fmaf(x[i], __ldg(Wi + (size_t)i * G + tid), z);
        const float* h = hs + t * F;
        for (int i = 0; i < F; ++i) z = fmaf(h[i], __ldg(Wh + (size_t)i * G + tid), z);
        const bool is_g = (tid >= 2 * F) && (tid < 3 * F);
        gates[t * G + tid] = is_g ? tanhf(z) : sigmoidf_(z);
      }
      __syncthreads();
      if (tid < F) {
        const float* g = gates + t * G;
        const float c = g[F + tid] * cs[t * F + tid] + g[tid] * g[2 * F + tid];
        cs[(t + 1) * F + tid] = c;
        hs[(t + 1) * F + tid] = g[3 * F + tid] * tanhf(c);
      }
      __syncthreads();
    }
    // ---------------------------------------------------------------- Dense head
    if (tid < F) hact[tid] = hs[T1 * F + tid];
    __syncthreads();
    for (int l = 0; l < D.L - 1; ++l) {
      const int din = D.din[l];
      if (tid < D.H) {
        float z = prm[D.oDb[l] + tid];
        for (int i = 0; i < din; ++i)
          z = fmaf(hact[l * W + i], __ldg(prm + D.oDk[l] + (size_t)i * D.H + tid), z);
        hact[(l + 1) * W + tid] = fmaxf(z, 0.f);  // relu; mask recovered as (a > 0)
      }
      __syncthreads();
    }
    if (warp == 0) {
      float acc = 0.f;
      const float* a = hact + (D.L - 1) * W;
      for (int i = lane; i < D.dlast; i += 32) acc = fmaf(a[i], prm[D.oWo + i], acc);
      acc = warp_sum(acc);
      if (lane == 0) {
        const float sc = acc + prm[D.obo];
        const float lab = label[src];
        const float lo = softplusf_(lab > 0.f ? -sc : sc);  // -log(where(label>0, p, 1-p))
        losses[s] = lo;
        if (logits) logits[s] = sc;
        const float sg = sigmoidf_(sc);
        scal[0] = (lab > 0.f ? sg - 1.f : sg) * inv_count;  // d(mean loss)/d score
      }
    }
    __syncthreads();
    if (!do_bwd) continue;  // (first is irrelevant without a backward pass)
    // ---------------------------------------------------------------- head backward
    const float ds = xgrad ? 1.f : scal[0];
    {
      const float* a = hact + (D.L - 1) * W;
      if (tid < D.dlast) {
        acc_part(D.oWo + tid, a[tid] * ds);
        hda[tid] = prm[D.oWo + tid] * ds;
      }
      if (tid == 0) acc_part(D.obo, ds);
    }
    __syncthreads();
    for (int l = D.L - 2; l >= 0; --l) {
      const int din = D.din[l];
      if (tid < D.H) {
        const float a = hact[(l + 1) * W + tid];
        const float dz = (a > 0.f) ? hda[tid] : 0.f;
        hdz[tid] = dz;
        acc_part(D.oDb[l] + tid, dz);
        for (int i = 0; i < din; ++i)
          acc_part(D.oDk[l] + (size_t)i * D.H + tid, hact[l * W + i] * dz);
      }
      __syncthreads();
      for (int i = warp; i < din; i += nwarps) {
        float acc = 0.f;
        for (int j = lane; j < D.H; j += 32)
          acc = fmaf(__ldg(prm + D.oDk[l] + (size_t)i * D.H + j), hdz[j], acc);
        acc = warp_sum(acc);
        if (lane == 0) hda[i] = acc;
      }
      __syncthreads();
    }
    if (tid < F) { dh[tid] = hda[tid]; dc[tid] = 0.f; }
    __syncthreads();
    // ---------------------------------------------------------------- LSTM backward through time
    for (int t = T1 - 1; t >= 0; --t) {
      if (tid < F) {
        float* g = gates + t * G;
        const float gi = g[tid], gf = g[F + tid], gg = g[2 * F + tid], go = g[3 * F + tid];
        const float tc = tanhf(cs[(t + 1) * F + tid]);
        const float dhj = dh[tid];
        const float dcj = dc[tid] + dhj * go * (1.f - tc * tc);
        g[tid] = dcj * gg * gi * (1.f - gi);
        g[F + tid] = dcj * cs[t * F + tid] * gf * (1.f - gf);
        g[2 * F + tid] = dcj * gi * (1.f - gg * gg);
        g[3 * F + tid] = dhj * tc * go * (1.f - go);
        dc[tid] = dcj * gf;
      }
      __syncthreads();
      if (xgrad) {  // d score / d x_t = Wi dz_t
        const float* dz = gates + t * G;
        for (int i = warp; i < n; i += nwarps) {
          float acc = 0.f;
          for (int j = lane; j < G; j += 32) acc = fmaf(__ldg(Wi + (size_t)i * G + j), dz[j], acc);
          acc = warp_sum(acc);
          if (lane == 0) dx_out[((size_t)s * T1 + t) * n + i] = acc;
        }
      }
      if (t > 0) {
        const float* dz = gates + t * G;
        for (int i = warp; i < F; i += nwarps) {
          float acc = 0.f;
          for (int j = lane; j < G; j += 32) acc = fmaf(__ldg(Wh + (size_t)i * G + j), dz[j], acc);
          acc = warp_sum(acc);
          if (lane == 0) dh[i] = acc;
        }
      }
      __syncthreads();
    }
    // ---------------------------------------------------------------- weight gradients
    if (tid < G && !xgrad) {
      float bsum = 0.f;
      for (int t = 0; t < T1; ++t) bsum += gates[t * G + tid];
      acc_part(D.obh + tid, bsum);
      for (int i = 0; i < n; ++i) {
        float acc = 0.f;
        for (int t = 0; t < T1; ++t) acc = fmaf(xs[t * n + i], gates[t * G + tid], acc);
        acc_part(D.oWi + (size_t)i * G + tid, acc);
      }
      for (int i = 0; i < F; i += 4) {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        for (int t = 0; t < T1; ++t) {
          const float dz = gates[t * G + tid];
          const float* h = hs + t * F + i;
          a0 = fmaf(h[0], dz, a0);
          if (i + 1 < F) a1 = fmaf(h[1], dz, a1);
          if (i + 2 < F) a2 = fmaf(h[2], dz, a2);
          if (i + 3 < F) a3 = fmaf(h[3], dz, a3);
        }
        acc_part(D.oWh + (size_t)i * G + tid, a0);
        if (i + 1 < F) acc_part(D.oWh + (size_t)(i + 1) * G + tid, a1);
        if (i + 2 < F) acc_part(D.oWh + (size_t)(i + 2) * G + tid, a2);
        if (i + 3 < F) acc_part(D.oWh + (size_t)(i + 3) * G + tid, a3);
      }
    }
  }
}

// grad[p] = sum over CTAs (fixed order) of partial[c][p]; loss[0] = inv_count * sum_s losses[s].
__global__ void critic_reduce_kernel(const float* partial, int nparts, long long P, float* grad,
                                     const float* losses, long long Bc, float inv_count,
                                     float* loss) {
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (grad != nullptr && p < P) {
    float acc = 0.f;
    for (int c = 0; c < nparts; ++c) acc += partial[(size_t)c * P + p];
    grad[p] = acc;
  }
  if (loss != nullptr && blockIdx.x == 0 && threadIdx.x < 32) {
    float acc = 0.f;
    for (long long s = threadIdx.x; s < Bc; s += 32) acc += losses[s];
    acc = warp_sum(acc);
    if (threadIdx.x == 0) loss[0] = acc * inv_count;
  }
}

// The three tail steps of one minibatch step in ONE launch (single-GPU scan, gmpc_critic_train_scan):
// deterministic reduction of the per-CTA partial gradients, global-norm clip, Adam.  Every block
// reduces its 256 parameters and its share of the squared norm, then the blocks meet at a grid-wide
// ticket barrier (the grid is ceil(P/256) blocks, far fewer than the SM count, so all of them are
// resident), sum the block norms in the same fixed order, and update their own 256 parameters.
// `expected` = number of tickets after this launch (the counter only ever grows).
// bc1 / bc2 = 1 - b^step are computed on the host in double precision.
__global__ void critic_reduce_clip_adam_kernel(const float* partial, int nparts, long long P,
                                               float* grad, const float* losses, long long Bc,
                                               float inv_count, float* loss, float* params, float* mom,
                                               float* vel, float lr, float max_norm, float b1, float b2,
                                               float eps, float bc1, float bc2, float* blocksq,
                                               unsigned long long* ticket, unsigned long long expected) {
  __shared__ float red[8];
  __shared__ float gn_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const long long p = (long long)blockIdx.x * blockDim.x + tid;
  float g = 0.f, mo = 0.f, ve = 0.f, pv = 0.f;
  if (p < P) {
    for (int c = 0; c < nparts; ++c) g += partial[(size_t)c * P + p];
    grad[p] = g;
    mo = mom[p]; ve = vel[p]; pv = params[p];   // fetched before the barrier, used after it
  }
  float ss = warp_sum(g * g);
  if (lane == 0) red[warp] = ss;
  if (loss != nullptr && blockIdx.x == 0 && warp == 0) {
    float acc = 0.f;
    for (long long s = lane; s < Bc; s += 32) acc += losses[s];
    acc = warp_sum(acc);
    if (lane == 0) loss[0] = acc * inv_count;
  }
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) tot += red[w];
    blocksq[blockIdx.x] = tot;
    __threadfence();
    atomicAdd(ticket, 1ULL);
    // (bounded: a launch of this scan that never ran -- its tickets were handed out on the host -- must surface as an
    // error, not as a hang)
    const long long t0 = clock64();
    while (*reinterpret_cast<volatile unsigned long long*>(ticket) < expected) {
      if (clock64() - t0 > 4000000000LL) __trap();
    }
    __threadfence();
  }
  __syncthreads();
  if (warp == 0) {
    float tot = 0.f;
    for (unsigned b = lane; b < gridDim.x; b += 32) tot += __ldcg(blocksq + b);
    tot = warp_sum(tot);
    if (lane == 0) gn_s = sqrtf(tot);
  }
  __syncthreads();
  if (p >= P) return;
  const float gn = gn_s;
  if (!(gn < max_norm)) g = g / gn * max_norm;
  mo = b1 * mo + (1.f - b1) * g;
  ve = b2 * ve + (1.f - b2) * g * g;
  mom[p] = mo;
  vel[p] = ve;
  params[p] = pv - lr * (mo / bc1) / (sqrtf(ve / bc2) + eps);
}

// optax.chain(clip_by_global_norm(max_norm), adam(lr)) + apply_updates on a flat vector.
// Single CTA: the vector is tiny and the global norm needs a grid-wide reduction otherwise.
__global__ void clip_adam_kernel(long long P, float* params, const float* grad, float* mom,
                                 float* vel, float bc1, float bc2, float lr, float max_norm, float gscale,
                                 float b1, float b2, float eps) {
  __shared__ float red[32];
  __shared__ float gn_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
  float ss = 0.f;
  for (long long i = tid; i < P; i += blockDim.x) {
    const float g = grad[i] * gscale;
    ss = fmaf(g, g, ss);
  }
  ss = warp_sum(ss);
  if (lane == 0) red[warp] = ss;
  __syncthreads();
  if (tid == 0) {
    float tot = 0.f;
    for (int w = 0; w < nw; ++w) tot += red[w];
    gn_s = sqrtf(tot);
  }
  __syncthreads();
  const float gn = gn_s;
  const bool clip = !(gn < max_norm);
  for (long long i = tid; i < P; i += blockDim.x) {
    float g = grad[i] * gscale;
    if (clip) g = g / gn * max_norm;
    const float mo = b1 * mom[i] + (1.f - b1) * g;
    const float ve = b2 * vel[i] + (1.f - b2) * g * g;
    mom[i] = mo;
    vel[i] = ve;
    params[i] = params[i] - lr * (mo / bc1) / (sqrtf(ve / bc2) + eps);
  }
}

// L2MPC.loss (norm/l2_policy.py:12-18): one warp per trajectory.
__global__ void l2_loss_kernel(long long B, int T1, int n, const float* __restrict__ X,
                               const float* __restrict__ des, float* loss) {
  const long long b = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (b >= B) return;
  const int lane = threadIdx.x & 31, per = T1 * n;
  float acc = 0.f;
  for (int e = lane; e < per; e += 32) {
    const float d = X[b * per + e] - des[b * per + e];
    acc = fmaf(d, d, acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) loss[b] = acc / (float)T1;
}

// K4: idx = argmin_k J_all[b, k] (first minimum on ties, jnp.argmin), gather the selected plan.
__global__ void select_best_kernel(long long B, int K, int T, int n, int m,
                                   const float* __restrict__ J_all, const float* __restrict__ U_all,
                                   const float* __restrict__ X_all, float* U_best, float* X_best,
                                   float* J_best, int* idx_best) {
  __shared__ float bj[32];
  __shared__ int bi[32];
  __shared__ int sel;
  for (long long b = blockIdx.x; b < B; b += gridDim.x) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nw = blockDim.x >> 5;
    float best = INFINITY;
    int bidx = 0x7fffffff;
    for (int k = tid; k < K; k += blockDim.x) {
      const float j = J_all[b * K + k];
      if (j < best || (j == best && k < bidx) || bidx == 0x7fffffff) { best = j; bidx = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float oj = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
      if (oi != 0x7fffffff && (bidx == 0x7fffffff || oj < best || (oj == best && oi < bidx))) {
        best = oj; bidx = oi;
      }
    }
    if (lane == 0) { bj[warp] = best; bi[warp] = bidx; }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < nw; ++w) {
        if (bi[w] != 0x7fffffff &&
            (bidx == 0x7fffffff || bj[w] < best || (bj[w] == best && bi[w] < bidx))) {
          best = bj[w]; bidx = bi[w];
        }
      }
      sel = bidx;
      J_best[b] = best;
      idx_best[b] = bidx;
    }
    __syncthreads();
    const long long q = b * K + sel;
    for (int e = tid; e < T * m; e += blockDim.x) U_best[b * T * m + e] = U_all[q * T * m + e];
    for (int e = tid; e < (T + 1) * n; e += blockDim.x)
      X_best[b * (T + 1) * n + e] = X_all[q * (T + 1) * n + e];
    __syncthreads();
  }
}

// Re-pack one Dense kernel W[K][N] into the forward (row-major [Kp][ldf]) and transposed
// (row-major [Np][ldb]) streams.  Destination buffers are pre-zeroed.
__global__ void pack_layer_kernel(const float* __restrict__ W, int K, int N, float* Wf, int ldf,
                                  float* Wb, int ldb) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= K * N) return;
  const int k = idx / N, o = idx - k * N;
  const float v = W[idx];
  Wf[(size_t)k * ldf + o] = v;
  Wb[(size_t)o * ldb + k] = v;
}

}  // namespace gmpc
