// gmpc_api.cu -- C ABI of libgmpc.so (see include/gmpc.h).  Host-side handle, weight packing,
// workspace management and kernel launches.  No torch types cross this boundary.
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <algorithm>
#include <string>
#include <vector>

#include "../../include/gmpc.h"
#include "common.cuh"
#include "critic.cuh"
#include "expert.cuh"
#include "diag.cuh"
#include "plan_ffma.cuh"
#include "ilqr.cuh"
#include "dynfit.cuh"
#include "plan_h16.cuh"
#include "plan_t128.cuh"
#include "smallgemm.cuh"

using namespace gmpc;

static thread_local std::string g_err;

static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CU_CHECK(expr)                                                                   \
  do {                                                                                   \
    cudaError_t e_ = (expr);                                                             \
    if (e_ != cudaSuccess)                                                               \
      return fail(GMPC_E_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e_));      \
  } while (0)

static inline int rup4(int v) { return (v + 3) & ~3; }

struct MlpPack {
  int L = 0;
  int dims[MAXL + 1] = {0};
  // device pointers into the handle's packed weight buffer
  float* Wf[MAXL] = {nullptr};
  float* Wb[MAXL] = {nullptr};
  float* bias[MAXL] = {nullptr};
  int ldf[MAXL] = {0}, ldb[MAXL] = {0};
};

struct gmpc_handle {
  gmpc_config cfg;
  int num_sms = 0;
  int path = GMPC_PATH_AUTO;
  int last_path = GMPC_PATH_FFMA;
  bool auto_retry = false;   // gmpc_plan_host was entered with path AUTO (range-check retries allowed)
  bool in_retry = false;
  int64_t launches = 0;
  bool have_weights = false;
  int maxt = 1;
  int hpad = 4;
  size_t smem_bytes = 0;
  MlpPack dyn, cost;
  float* d_wpack = nullptr;
  float* d_mpcw = nullptr;
  // FFMA per-CTA scratch
  float *ws_X = nullptr, *ws_G = nullptr, *ws_U = nullptr, *ws_M = nullptr, *ws_V = nullptr;
  uint32_t* ws_mask = nullptr;
  // growable scratch for K>1 candidate outputs and host staging
  void* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  void* d_stage = nullptr;
  size_t stage_bytes = 0;
  // critic
  CriticDims cd;
  float* d_partial = nullptr;
  float* d_losses = nullptr;
  float* d_fuse = nullptr;   // 8-byte ticket counter, then per-block squared norms of the fused tail kernel
  unsigned long long fuse_tickets = 0;  // tickets handed out so far (the device counter only grows)
  size_t losses_cap = 0;
  int critic_grid = 0;
  // tensor-core path state: the 32-trajectory fp16-split kernel (latency tile, wide layers) and the
  // 128-trajectory tile kernel (activations in tensor memory)
  H16State h16;
  T128State t128;
  size_t smem_optin = 0;        // cudaDevAttrMaxSharedMemoryPerBlockOptin, read once
  bool ilqr_full_tiles = false; // GMPC_ILQR_FULL_TILES / GMPC_ILQR_NO_PACK: experiment switches, read once at create
  bool ilqr_no_pack = false;
  bool ilqr_attr = false;
  // iLQR kernel scratch (allocated on first use)
  uint32_t* d_fit_masks = nullptr;   // dynfit kernel ReLU masks (grown on demand)
  size_t fit_masks_bytes = 0;
  bool fit_attr = false;
  float* d_ilqr_ws = nullptr;
  size_t ilqr_ws_bytes = 0;
  void* d_vjp_ws = nullptr;          // activations / tangents / masks / cotangents of gmpc_cost_mixed_vjp
  size_t vjp_ws_bytes = 0;
  long long* d_ilqr_stats = nullptr;
};

extern "C" const char* gmpc_last_error(void) { return g_err.c_str(); }

static int grow(void** p, size_t* cap, size_t need) {
  if (need <= *cap) return GMPC_OK;
  if (*p) cudaFree(*p);
  *p = nullptr;
  *cap = 0;
  CU_CHECK(cudaMalloc(p, need));
  *cap = need;
  return GMPC_OK;
}

static void fill_dims(MlpPack& mp, int in, int hidden, int out, int L) {
  mp.L = L;
  mp.dims[0] = in;
  for (int l = 1; l < L; ++l) mp.dims[l] = hidden;
  mp.dims[L] = out;
}

static size_t pack_floats(const MlpPack& mp) {
  size_t tot = 0;
  for (int l = 0; l < mp.L; ++l) {
    tot += (size_t)rup4(mp.dims[l]) * rup4(mp.dims[l + 1]) * 2;  // forward + transposed
    tot += rup4(mp.dims[l + 1]);                                  // bias
  }
  return tot;
}

static float* carve(MlpPack& mp, float* base) {
  for (int l = 0; l < mp.L; ++l) {
    const int Kp = rup4(mp.dims[l]), Np = rup4(mp.dims[l + 1]);
    mp.ldf[l] = Np;
    mp.ldb[l] = Kp;
    mp.Wf[l] = base; base += (size_t)Kp * Np;
    mp.Wb[l] = base; base += (size_t)Np * Kp;
    mp.bias[l] = base; base += Np;
  }
  return base;
}

static void make_layer(LayerDesc& d, const float* W, const float* bias, int Ki_true, int No,
                       int ld) {
  d.W = W;
  d.bias = bias;
  d.Ki = rup4(Ki_true);
  d.No = No;
  d.ld = ld;
  d.kc = std::max(4, (STAGE_FLOATS / ld) & ~3);
  d.nchunks = (d.Ki + d.kc - 1) / d.kc;
  d.pad_ = 0;
}

static void make_dirs(const MlpPack& mp, DirDesc& fwd, DirDesc& bwd) {
  fwd.L = bwd.L = mp.L;
  fwd.pad_ = bwd.pad_ = 0;
  for (int l = 0; l < mp.L; ++l) {
    make_layer(fwd.layer[l], mp.Wf[l], mp.bias[l], mp.dims[l], mp.dims[l + 1], mp.ldf[l]);
    const int lt = mp.L - 1 - l;  // backward pass visits transposed layers L-1 .. 0
    make_layer(bwd.layer[l], mp.Wb[lt], nullptr, mp.dims[lt + 1], mp.dims[lt], mp.ldb[lt]);
  }
}

static size_t ffma_smem_bytes(const gmpc_config& c, int hpad) {
  const int n4 = rup4(c.n), nm4 = rup4(c.n + c.m), f4 = rup4(c.cost_fout);
  return sizeof(float) *
         ((size_t)2 * hpad * RT + (size_t)NSTAGE * STAGE_FLOATS + (size_t)(2 * nm4 + n4 + f4 + c.n) * RT);
}

static void critic_dims(const gmpc_config& c, CriticDims& d) {
  memset(&d, 0, sizeof(d));
  d.n = c.n; d.F = c.critic_features; d.L = c.critic_layers; d.H = c.critic_hidden;
  long long o = 0;
  d.oWi = o; o += (long long)d.n * 4 * d.F;
  d.oWh = o; o += (long long)d.F * 4 * d.F;
  d.obh = o; o += 4 * d.F;
  int din = d.F;
  for (int l = 0; l < d.L - 1; ++l) {
    d.din[l] = din;
    d.oDk[l] = o; o += (long long)din * d.H;
    d.oDb[l] = o; o += d.H;
    din = d.H;
  }
  d.dlast = din;
  d.oWo = o; o += din;
  d.obo = o; o += 1;
  d.P = o;
}

extern "C" int gmpc_create(const gmpc_config* cfg, gmpc_handle** out) {
  if (!cfg || !out) return fail(GMPC_E_ARG, "gmpc_create: null argument");
  const gmpc_config& c = *cfg;
  if (c.n < 1 || c.m < 1 || c.T < 1) return fail(GMPC_E_ARG, "gmpc_create: n, m, T must be >= 1");
  if (c.dyn_layers < 1 || c.dyn_layers > MAXL || c.cost_layers < 1 || c.cost_layers > MAXL)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_create: num_layers must be in [1, 8]");
  if (c.dyn_hidden < 1 || c.cost_hidden < 1 || c.cost_fout < 1)
    return fail(GMPC_E_ARG, "gmpc_create: hidden/fout must be >= 1");
  const int hmax = std::max(c.dyn_layers > 1 ? c.dyn_hidden : 1, c.cost_layers > 1 ? c.cost_hidden : 1);
  if (hmax > 512 || c.n + c.m > 256 || c.cost_fout > 256)
    return fail(GMPC_E_UNSUPPORTED,
                "gmpc_create: supported widths are hidden <= 512, n+m <= 256, fout <= 256");
  if (c.critic_features < 0 || 4 * c.critic_features > 1024 || c.critic_hidden > 1024 ||
      (c.critic_features > 0 && (c.critic_layers < 1 || c.critic_layers > MAXL)))
    return fail(GMPC_E_UNSUPPORTED, "gmpc_create: critic needs features <= 256, hidden <= 1024, layers in [1, 8]");
  CU_CHECK(cudaSetDevice(c.device));
  cudaDeviceProp prop;
  CU_CHECK(cudaGetDeviceProperties(&prop, c.device));
  if (prop.major != 10)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_create: libgmpc is built for sm_100a (B200) only");

  gmpc_handle* h = new gmpc_handle();
  h->cfg = c;
  h->num_sms = prop.multiProcessorCount;
  h->maxt = hmax <= 256 ? 1 : 2;
  h->hpad = std::max(4, rup4(hmax));
  h->smem_bytes = ffma_smem_bytes(c, h->hpad);
  if (h->smem_bytes > (size_t)prop.sharedMemPerBlockOptin) {
    delete h;
    return fail(GMPC_E_UNSUPPORTED, "gmpc_create: shape needs more shared memory than one SM has");
  }
  fill_dims(h->dyn, c.n + c.m, c.dyn_hidden, c.n, c.dyn_layers);
  fill_dims(h->cost, c.n, c.cost_hidden, c.cost_fout, c.cost_layers);
  const size_t wfl = pack_floats(h->dyn) + pack_floats(h->cost);
  cudaError_t e = cudaMalloc(&h->d_wpack, wfl * sizeof(float));
  if (e == cudaSuccess) e = cudaMemset(h->d_wpack, 0, wfl * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->d_mpcw, 4 * sizeof(float));
  carve(h->cost, carve(h->dyn, h->d_wpack));
  // per-CTA scratch for a persistent grid of one CTA per SM
  const size_t G = h->num_sms;
  const size_t sx = (size_t)(c.T + 1) * c.n * RT, su = (size_t)c.T * c.m * RT;
  const size_t smk = ((size_t)c.T * (c.dyn_layers - 1) + (c.cost_layers - 1)) * 2 * NTHREADS + 1;
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_X, G * sx * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_G, G * sx * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_U, G * su * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_M, G * su * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_V, G * su * sizeof(float));
  if (e == cudaSuccess) e = cudaMalloc(&h->ws_mask, G * smk * sizeof(uint32_t));
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(plan_ffma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)prop.sharedMemPerBlockOptin);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(plan_ffma_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)prop.sharedMemPerBlockOptin);
  if (e == cudaSuccess && c.critic_features > 0) {
    critic_dims(c, h->cd);
    h->critic_grid = 2 * h->num_sms;
    e = cudaMalloc(&h->d_partial, (size_t)h->critic_grid * h->cd.P * sizeof(float));
    if (e == cudaSuccess)
      e = cudaFuncSetAttribute(critic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)prop.sharedMemPerBlockOptin);
  }
  if (e != cudaSuccess) {
    gmpc_destroy(h);
    return fail(GMPC_E_CUDA, std::string("gmpc_create: ") + cudaGetErrorString(e));
  }
  h->smem_optin = (size_t)prop.sharedMemPerBlockOptin;
  h->ilqr_full_tiles = getenv("GMPC_ILQR_FULL_TILES") != nullptr;
  h->ilqr_no_pack = getenv("GMPC_ILQR_NO_PACK") != nullptr;
  int rc = h16_create(h->h16, c, h->dyn.dims, h->cost.dims, prop);
  if (rc == GMPC_OK)
    rc = t128_create(h->t128, c, h->dyn.dims, h->cost.dims, prop.multiProcessorCount, (size_t)prop.sharedMemPerBlockOptin);
  if (rc != GMPC_OK) {
    gmpc_destroy(h);
    return fail(rc, "gmpc_create: tensor-core path setup failed (CUDA error)");
  }
  *out = h;
  return GMPC_OK;
}

extern "C" int gmpc_destroy(gmpc_handle* h) {
  if (!h) return GMPC_OK;
  cudaSetDevice(h->cfg.device);
  h16_destroy(h->h16);
  t128_destroy(h->t128);
  cudaFree(h->d_wpack); cudaFree(h->d_mpcw);
  cudaFree(h->ws_X); cudaFree(h->ws_G); cudaFree(h->ws_U); cudaFree(h->ws_M); cudaFree(h->ws_V);
  cudaFree(h->ws_mask); cudaFree(h->d_scratch); cudaFree(h->d_stage);
  cudaFree(h->d_partial); cudaFree(h->d_losses); cudaFree(h->d_fuse);
  cudaFree(h->d_ilqr_ws);
  cudaFree(h->d_vjp_ws); cudaFree(h->d_ilqr_stats); cudaFree(h->d_fit_masks);
  delete h;
  return GMPC_OK;
}

extern "C" int64_t gmpc_critic_param_count(const gmpc_handle* h) {
  return (h && h->cfg.critic_features > 0) ? h->cd.P : 0;
}

extern "C" int gmpc_set_path(gmpc_handle* h, int path) {
  if (!h || path < GMPC_PATH_AUTO || path > GMPC_PATH_T128) return fail(GMPC_E_ARG, "gmpc_set_path: bad argument");
  if (path == GMPC_PATH_TC)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_set_path: the 3xTF32 kernel was superseded by the fp16-split kernels and is "
                                    "no longer part of the library (unsupported)");
  if (path == GMPC_PATH_T128 && !h->t128.supported)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_set_path: tensor-core path unsupported for this shape: " + h->t128.why);
  if ((path == GMPC_PATH_TC16 || path == GMPC_PATH_TC16S) && !h->h16.supported)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_set_path: tensor-core path unsupported for this shape: " + h->h16.why);
  h->path = path;
  return GMPC_OK;
}
extern "C" int gmpc_last_path(const gmpc_handle* h) { return h ? h->last_path : GMPC_E_ARG; }
extern "C" int64_t gmpc_launch_count(const gmpc_handle* h) { return h ? h->launches : 0; }

static int pack_mlp(gmpc_handle* h, MlpPack& mp, const float* const* W, const float* const* b,
                    cudaStream_t st) {
  for (int l = 0; l < mp.L; ++l) {
    if (!W[l] || !b[l]) return fail(GMPC_E_ARG, "gmpc_set_weights: null layer pointer");
    const int K = mp.dims[l], N = mp.dims[l + 1];
    pack_layer_kernel<<<(K * N + 255) / 256, 256, 0, st>>>(W[l], K, N, mp.Wf[l], mp.ldf[l],
                                                            mp.Wb[l], mp.ldb[l]);
    ++h->launches;
    CU_CHECK(cudaMemcpyAsync(mp.bias[l], b[l], sizeof(float) * N, cudaMemcpyDeviceToDevice, st));
  }
  return GMPC_OK;
}

extern "C" int gmpc_set_weights(gmpc_handle* h, const float* const* dyn_W,
                                const float* const* dyn_b, const float* const* cost_W,
                                const float* const* cost_b, const float* mpc_weights,
                                void* stream) {
  if (!h || !dyn_W || !dyn_b || !cost_W || !cost_b || !mpc_weights)
    return fail(GMPC_E_ARG, "gmpc_set_weights: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  int rc = pack_mlp(h, h->dyn, dyn_W, dyn_b, st);
  if (rc) return rc;
  rc = pack_mlp(h, h->cost, cost_W, cost_b, st);
  if (rc) return rc;
  CU_CHECK(cudaMemcpyAsync(h->d_mpcw, mpc_weights, 3 * sizeof(float), cudaMemcpyDeviceToDevice, st));
  rc = t128_set_weights(h->t128, dyn_W, dyn_b, cost_W, cost_b, st, &h->launches);
  if (rc) return rc;
  rc = h16_set_weights(h->h16, dyn_W, dyn_b, cost_W, cost_b, st, &h->launches);
  if (rc) return rc;
  CU_CHECK(cudaGetLastError());
  h->have_weights = true;
  return GMPC_OK;
}

// Which kernel family serves a call: explicit choice, else by MEASURED time on B200 (C2 dims, 20 iterations):
//   * the 32-trajectory tensor-core kernel finishes a wave of 32 x SMs trajectories in 11.9 ms whatever the
//     batch, and a single trajectory in 2.2 ms (C1 dims) / 12.7 ms (C2 dims) against 9.2 / 63 ms of the fp32
//     CUDA-core kernel: it serves every batch of up to one wave, also the single state of an acting call
//     (the north star's "tile >= 64" rule would pick the 4 x slower kernel there);
//   * the 128-trajectory kernel finishes a wave of 128 x SMs trajectories in 23.5 ms (2 x the trajectories per
//     second): everything larger than one wave of the 32-trajectory kernel.
// Both tensor-core defaults rescale the forward operands per trajectory, so states of any magnitude stay inside
// the fp16 hi/lo range (the un-scaled GMPC_PATH_TC16 is opt-in only).  The fp32 kernel serves the shapes the
// tensor-core kernels do not cover (hidden < 64 or > 512, n + m > 32, fout > 32) and is the last resort of the
// operand range check.
static int pick_path(gmpc_handle* h, int64_t NQ) {
  if (h->path != GMPC_PATH_AUTO) return h->path;
  const bool dense = h16_worthwhile(h->h16, NQ);   // hidden width >= 64: a real contraction in the feature dimension
  if (dense && h->t128.supported && NQ > (int64_t)H_NB * h->num_sms) return GMPC_PATH_T128;
  if (dense && h->h16.supported) return GMPC_PATH_TC16S;
  return GMPC_PATH_FFMA;
}

// Fill the parts of PlanParams shared by every mode and launch the planner kernel of the
// selected path (tcgen05 or FFMA).
static int launch_ffma(gmpc_handle* h, PlanParams& P, cudaStream_t st) {
  const gmpc_config& c = h->cfg;
  make_dirs(h->dyn, P.dir[DIR_DYN_F], P.dir[DIR_DYN_B]);
  make_dirs(h->cost, P.dir[DIR_COST_F], P.dir[DIR_COST_B]);
  P.n = c.n; P.m = c.m; P.T = c.T;
  P.hpad = h->hpad;
  P.fout = c.cost_fout;
  P.mpcw = h->d_mpcw;
  P.ws_X = h->ws_X; P.ws_G = h->ws_G; P.ws_U = h->ws_U; P.ws_M = h->ws_M; P.ws_V = h->ws_V;
  P.ws_mask = h->ws_mask;
  P.ntiles = (int)((P.NQ + RT - 1) / RT);
  const int grid = std::min(P.ntiles, h->num_sms);
  if (grid <= 0) return GMPC_OK;
  const int path = pick_path(h, P.NQ);
  if (path == GMPC_PATH_T128 || path == GMPC_PATH_TC16 || path == GMPC_PATH_TC16S) {
    int rc = path == GMPC_PATH_T128 ? t128_launch(h->t128, P, st, &h->launches)
                                    : h16_launch(h->h16, P, path == GMPC_PATH_TC16S, st, &h->launches);
    if (rc) return fail(rc, "tensor-core planner launch failed");
    h->last_path = path;
    return GMPC_OK;
  }
  if (h->maxt == 1)
    plan_ffma_kernel<1><<<grid, NTHREADS, h->smem_bytes, st>>>(P);
  else
    plan_ffma_kernel<2><<<grid, NTHREADS, h->smem_bytes, st>>>(P);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  h->last_path = GMPC_PATH_FFMA;
  return GMPC_OK;
}

static int check_ready(gmpc_handle* h, const char* who, int64_t B) {
  if (!h) return fail(GMPC_E_ARG, std::string(who) + ": null handle");
  if (!h->have_weights) return fail(GMPC_E_STATE, std::string(who) + ": call gmpc_set_weights first");
  if (B < 0 || B > (int64_t)1 << 40) return fail(GMPC_E_ARG, std::string(who) + ": bad batch size");
  cudaError_t e = cudaSetDevice(h->cfg.device);
  if (e != cudaSuccess) return fail(GMPC_E_CUDA, cudaGetErrorString(e));
  return GMPC_OK;
}

extern "C" int gmpc_rollout(gmpc_handle* h, int64_t B, const float* x0, const float* U, float* X,
                            void* stream) {
  int rc = check_ready(h, "gmpc_rollout", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (!x0 || !U || !X) return fail(GMPC_E_ARG, "gmpc_rollout: null argument");
  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.mode = MODE_ROLLOUT; P.iters = 0; P.use_cost = 0; P.final_fwd = 1;
  P.K = 1; P.NQ = B;
  P.x0 = x0; P.U_in = U; P.goal = nullptr; P.X_out = X;
  return launch_ffma(h, P, (cudaStream_t)stream);
}

extern "C" int gmpc_objective_grad(gmpc_handle* h, int64_t B, const float* x0, const float* U,
                                   const float* goal, float* J, float* dU, float* X, float* lam,
                                   void* stream) {
  int rc = check_ready(h, "gmpc_objective_grad", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (!x0 || !U || !goal) return fail(GMPC_E_ARG, "gmpc_objective_grad: null argument");
  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.mode = MODE_OBJGRAD; P.use_cost = 1;
  const bool need_bwd = (dU != nullptr) || (lam != nullptr);
  P.iters = need_bwd ? 1 : 0;
  P.final_fwd = need_bwd ? 0 : 1;
  P.K = 1; P.NQ = B;
  P.x0 = x0; P.U_in = U; P.goal = goal;
  P.J_out = J; P.dU_out = dU; P.X_out = X; P.lam_out = lam;
  return launch_ffma(h, P, (cudaStream_t)stream);
}

extern "C" int gmpc_l2_loss_grad(gmpc_handle* h, int64_t B, const float* x0, const float* U,
                                 const float* desired, float* loss, float* dU, float* X,
                                 void* stream) {
  int rc = check_ready(h, "gmpc_l2_loss_grad", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (!x0 || !U || !desired) return fail(GMPC_E_ARG, "gmpc_l2_loss_grad: null argument");
  PlanParams P;
  memset(&P, 0, sizeof(P));
  P.mode = MODE_L2GRAD; P.use_cost = 0; P.iters = 1; P.final_fwd = 0;
  P.K = 1; P.NQ = B;
  P.x0 = x0; P.U_in = U; P.goal = desired;
  P.J_out = loss; P.dU_out = dU; P.X_out = X;
  return launch_ffma(h, P, (cudaStream_t)stream);
}

extern "C" int gmpc_plan(gmpc_handle* h, int64_t B, int32_t K, const float* x0, const float* U0,
                         const float* goal, int32_t method, int32_t N, float lr, float b1,
                         float b2, float eps, float* U_best, float* X_best, float* J_best,
                         int32_t* idx_best, float* J_all, void* stream) {
  int rc = check_ready(h, "gmpc_plan", B);
  if (rc) return rc;
  if (K < 1 || N < 0 || (method != GMPC_METHOD_GRAD && method != GMPC_METHOD_ADAM))
    return fail(GMPC_E_ARG, "gmpc_plan: need K >= 1, N >= 0, method in {GRAD, ADAM}");
  if (B == 0) return GMPC_OK;
  if (!x0 || !U0 || !goal || !U_best || !X_best || !J_best || !idx_best)
    return fail(GMPC_E_ARG, "gmpc_plan: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gmpc_config& c = h->cfg;
  const int64_t NQ = B * K;
  const size_t ub = (size_t)c.T * c.m, xb = (size_t)(c.T + 1) * c.n;
  float *U_out = U_best, *X_out = X_best, *J_out = J_best;
  if (K > 1) {
    const size_t need = sizeof(float) * ((size_t)NQ * (ub + xb + 1));
    rc = grow(&h->d_scratch, &h->scratch_bytes, need);
    if (rc) return rc;
    U_out = (float*)h->d_scratch;
    X_out = U_out + (size_t)NQ * ub;
    J_out = J_all ? J_all : X_out + (size_t)NQ * xb;
  }
  {
    PlanParams P;
    memset(&P, 0, sizeof(P));
    P.mode = MODE_PLAN; P.method = method; P.iters = N; P.use_cost = 1; P.final_fwd = 1;
    P.K = K; P.NQ = NQ;
    P.lr = lr; P.b1 = b1; P.b2 = b2; P.eps = eps;
    P.x0 = x0; P.U_in = U0; P.goal = goal;
    P.U_out = U_out; P.X_out = X_out; P.J_out = J_out;
    rc = launch_ffma(h, P, st);
    if (rc) return rc;
  }
  if (K > 1) {
    const int grid = (int)std::min<int64_t>(B, 8 * h->num_sms);
    select_best_kernel<<<grid, 128, 0, st>>>(B, K, c.T, c.n, c.m, J_out, U_out, X_out, U_best,
                                             X_best, J_best, idx_best);
    ++h->launches;
  } else {
    CU_CHECK(cudaMemsetAsync(idx_best, 0, sizeof(int32_t) * B, st));
    if (J_all) CU_CHECK(cudaMemcpyAsync(J_all, J_best, sizeof(float) * B, cudaMemcpyDeviceToDevice, st));
  }
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

// ilqr_solve (policy/optimizers.py:10-21): the whole trajax iLQR loop as one kernel (csrc/ilqr.cuh).
struct BilevelArgs {
  const float* desired;   // L2 mode: desired states; generic mode: d loss / d X (is_dLdX)
  float *loss, *Bvec, *hess, *H, *dxT, *gw;
  const float* V;
  int is_dLdX;
};

static int ilqr_launch(gmpc_handle* h, int64_t B, const float* x0, const float* U0,
                       const float* goal, const gmpc_ilqr_options* opt, float* X, float* U,
                       float* obj, float* gradient, float* adjoints, int32_t* iteration,
                       float* lqr_A, float* lqr_B, const BilevelArgs* bl, void* stream) {
  int rc = check_ready(h, "gmpc_ilqr", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (!x0 || !U0 || !goal || !opt || !X || !U || !obj) return fail(GMPC_E_ARG, "gmpc_ilqr: null argument");
  if (opt->maxiter < 0) return fail(GMPC_E_ARG, "gmpc_ilqr: maxiter must be >= 0");
  const gmpc_config& c = h->cfg;
  if (c.m > IL_MAXM) return fail(GMPC_E_UNSUPPORTED, "gmpc_ilqr: action size above 16");
  cudaStream_t st = (cudaStream_t)stream;
  const IlqrSmem SL = ilqr_smem_layout(c.n, c.m, c.cost_fout, h->hpad);
  if (SL.bytes > h->smem_optin)
    return fail(GMPC_E_UNSUPPORTED, "gmpc_ilqr: the Riccati matrices of a 32-trajectory tile do not fit shared memory (state size too large)");
  const IlqrWs WL = ilqr_ws_layout(c.n, c.m, c.T, c.cost_fout, bl != nullptr);
  const size_t need = (size_t)h->num_sms * WL.total * sizeof(float);
  if (need > h->ilqr_ws_bytes) {
    CU_CHECK(cudaStreamSynchronize(st));
    rc = grow((void**)&h->d_ilqr_ws, &h->ilqr_ws_bytes, need);
    if (rc) return rc;
  }
  if (!h->d_ilqr_stats) {
    CU_CHECK(cudaMalloc(&h->d_ilqr_stats, 2 * sizeof(long long)));
    CU_CHECK(cudaMemset(h->d_ilqr_stats, 0, 2 * sizeof(long long)));
  }
  if (!h->ilqr_attr) {
    CU_CHECK(cudaFuncSetAttribute(ilqr_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
    CU_CHECK(cudaFuncSetAttribute(ilqr_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)h->smem_optin));
    h->ilqr_attr = true;
  }
  IlqrParams Q;
  memset(&Q, 0, sizeof(Q));
  PlanParams& P = Q.pp;
  make_dirs(h->dyn, P.dir[DIR_DYN_F], P.dir[DIR_DYN_B]);
  make_dirs(h->cost, P.dir[DIR_COST_F], P.dir[DIR_COST_B]);
  P.n = c.n; P.m = c.m; P.T = c.T; P.K = 1;
  P.hpad = h->hpad;
  P.fout = c.cost_fout;
  P.use_cost = 1;
  P.mpcw = h->d_mpcw;
  P.ws_mask = h->ws_mask;
  P.NQ = B;
  // a batch smaller than 32 x SMs is cut into smaller tiles so that it still covers the machine: fewer
  // trajectories share a CTA's lock step (a tile runs until its slowest lane stops) and tiny tiles pack
  // Jacobian rows / line-search trials into the idle lanes
  int tile_traj = (int)std::max<int64_t>(1, (B + h->num_sms - 1) / h->num_sms);
  // measured (C1 dims, one B200): 128 states 68 ms vs 366 ms with full tiles, 1024 states 337 vs 478 ms;
  // from about half-full tiles on there is nothing left to gain (4096 states: 28-lane tiles are no faster)
  if (tile_traj > RT / 2 || h->ilqr_full_tiles) tile_traj = RT;
  Q.tile_traj = tile_traj;
  P.ntiles = (int)((B + tile_traj - 1) / tile_traj);
  P.x0 = x0; P.U_in = U0; P.goal = goal;
  P.X_out = X; P.U_out = U; P.J_out = obj; P.dU_out = gradient; P.lam_out = adjoints;
  Q.maxiter = opt->maxiter;
  Q.gthr = opt->grad_norm_threshold;
  Q.alpha0 = opt->alpha_0;
  Q.alpha_min = opt->alpha_min;
  Q.grad_lag = opt->gradient_lag ? 1 : 0;
  Q.it_out = iteration;
  Q.A_out = lqr_A;
  Q.B_out = lqr_B;
  Q.ws = h->d_ilqr_ws;
  Q.ws_stride = (long long)WL.total;
  Q.stats = h->d_ilqr_stats;
  Q.pack_small = h->ilqr_no_pack ? 0 : 1;   // experiment switch: lane = trajectory everywhere
  if (bl) {
    Q.desired = bl->desired; Q.bl_loss = bl->loss; Q.bl_B = bl->Bvec; Q.bl_hess = bl->hess;
    Q.bl_H = bl->H; Q.bl_dxT = bl->dxT; Q.bl_gw = bl->gw; Q.bl_V = bl->V;
    Q.bl_generic = bl->is_dLdX;
  }
  const int grid = std::min(P.ntiles, h->num_sms);
  if (h->maxt == 1)
    ilqr_kernel<1><<<grid, NTHREADS, SL.bytes, st>>>(Q);
  else
    ilqr_kernel<2><<<grid, NTHREADS, SL.bytes, st>>>(Q);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

extern "C" int gmpc_ilqr(gmpc_handle* h, int64_t B, const float* x0, const float* U0,
                         const float* goal, const gmpc_ilqr_options* opt, float* X, float* U,
                         float* obj, float* gradient, float* adjoints, int32_t* iteration,
                         float* lqr_A, float* lqr_B, void* stream) {
  return ilqr_launch(h, B, x0, U0, goal, opt, X, U, obj, gradient, adjoints, iteration, lqr_A, lqr_B,
                     nullptr, stream);
}

// gmpc_ilqr with HOST buffers: the end-to-end call of a caller that is not GPU aware (acting loop of
// utils.run_dm_policy: one state in, one plan out).  Synchronises `stream` before returning.
extern "C" int gmpc_ilqr_host(gmpc_handle* h, int64_t B, const float* x0_host, const float* U0_host,
                              const float* goal_host, const gmpc_ilqr_options* opt, float* X_host,
                              float* U_host, float* obj_host, float* gradient_host,
                              float* adjoints_host, int32_t* iteration_host, void* stream) {
  int rc = check_ready(h, "gmpc_ilqr_host", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (!x0_host || !U0_host || !goal_host || !opt || !X_host || !U_host || !obj_host)
    return fail(GMPC_E_ARG, "gmpc_ilqr_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gmpc_config& c = h->cfg;
  const size_t ub = (size_t)B * c.T * c.m, xb = (size_t)B * (c.T + 1) * c.n, x0b = (size_t)B * c.n;
  const size_t tot = x0b + 2 * ub + 2 * xb + B /*obj*/ + ub /*grad*/ + xb /*adjoints*/ + B /*iteration*/;
  rc = grow(&h->d_stage, &h->stage_bytes, tot * sizeof(float));
  if (rc) return rc;
  float* d_x0 = (float*)h->d_stage;
  float* d_U0 = d_x0 + x0b;
  float* d_goal = d_U0 + ub;
  float* d_X = d_goal + xb;
  float* d_U = d_X + xb;
  float* d_obj = d_U + ub;
  float* d_g = d_obj + B;
  float* d_lam = d_g + ub;
  int32_t* d_it = (int32_t*)(d_lam + xb);
  CU_CHECK(cudaMemcpyAsync(d_x0, x0_host, x0b * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_CHECK(cudaMemcpyAsync(d_U0, U0_host, ub * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_CHECK(cudaMemcpyAsync(d_goal, goal_host, xb * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = gmpc_ilqr(h, B, d_x0, d_U0, d_goal, opt, d_X, d_U, d_obj, gradient_host ? d_g : nullptr,
                 adjoints_host ? d_lam : nullptr, iteration_host ? d_it : nullptr, nullptr, nullptr, st);
  if (rc) return rc;
  CU_CHECK(cudaMemcpyAsync(X_host, d_X, xb * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemcpyAsync(U_host, d_U, ub * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemcpyAsync(obj_host, d_obj, B * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (gradient_host) CU_CHECK(cudaMemcpyAsync(gradient_host, d_g, ub * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (adjoints_host) CU_CHECK(cudaMemcpyAsync(adjoints_host, d_lam, xb * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (iteration_host) CU_CHECK(cudaMemcpyAsync(iteration_host, d_it, B * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaStreamSynchronize(st));
  return GMPC_OK;
}

// bilevel_optimization (policy/optimizers.py:34-75) for loss = L2MPC.loss: iLQR + the bilevel tail in
// the same kernel launch.
extern "C" int gmpc_bilevel_l2(gmpc_handle* h, int64_t B, const float* x0, const float* U0,
                               const float* goal, const float* desired, const gmpc_ilqr_options* opt,
                               float* X, float* U, float* obj, float* low_level_grad,
                               int32_t* iteration, float* loss, float* loss_grad_U, float* hessian,
                               float* H, float* dxT, float* grad_mpc_weights, const float* V,
                               void* stream) {
  if (h && B == 0) return GMPC_OK;
  if (!desired || !loss || !H || !dxT || !grad_mpc_weights)
    return fail(GMPC_E_ARG, "gmpc_bilevel_l2: null argument");
  if (V && hessian) return fail(GMPC_E_ARG, "gmpc_bilevel_l2: a given direction V skips the Hessian");
  BilevelArgs bl{desired, loss, loss_grad_U, hessian, H, dxT, grad_mpc_weights, V, 0};
  return ilqr_launch(h, B, x0, U0, goal, opt, X, U, obj, low_level_grad, nullptr, iteration, nullptr,
                     nullptr, &bl, stream);
}

// The same tail for ANY loss of the planned states, given its gradient dL/dX at the plan (e.g. the
// generator loss of gan/js_policy.py:60-68 through gmpc_critic_input_grad): evaluated at U itself.
extern "C" int gmpc_bilevel_tail(gmpc_handle* h, int64_t B, const float* x0, const float* U,
                                 const float* goal, const float* dLdX, float* loss_grad_U,
                                 float* hessian, float* H, float* dxT, float* grad_mpc_weights,
                                 void* stream) {
  if (h && B == 0) return GMPC_OK;
  if (!dLdX || !H || !dxT || !grad_mpc_weights) return fail(GMPC_E_ARG, "gmpc_bilevel_tail: null argument");
  int rc = check_ready(h, "gmpc_bilevel_tail", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  const gmpc_config& c = h->cfg;
  // rollout / objective / plan outputs of the embedded iLQR (0 iterations) are not wanted: scratch
  const size_t fl = (size_t)B * ((size_t)(c.T + 1) * c.n + (size_t)c.T * c.m + 1);
  rc = grow(&h->d_scratch, &h->scratch_bytes, fl * sizeof(float));
  if (rc) return rc;
  float* sX = (float*)h->d_scratch;
  float* sU = sX + (size_t)B * (c.T + 1) * c.n;
  float* sJ = sU + (size_t)B * c.T * c.m;
  const gmpc_ilqr_options opt{0, 0.f, 1.f, 0.f, 0};
  BilevelArgs bl{dLdX, nullptr, loss_grad_U, hessian, H, dxT, grad_mpc_weights, nullptr, 1};
  return ilqr_launch(h, B, x0, U, goal, &opt, sX, sU, sJ, nullptr, nullptr, nullptr, nullptr, nullptr,
                     &bl, stream);
}

extern "C" int gmpc_ilqr_stats(gmpc_handle* h, int64_t* outer_iterations, int64_t* rollouts, void* stream) {
  if (!h || !outer_iterations || !rollouts) return fail(GMPC_E_ARG, "gmpc_ilqr_stats: null argument");
  *outer_iterations = 0;
  *rollouts = 0;
  if (!h->d_ilqr_stats) return GMPC_OK;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  cudaStream_t st = (cudaStream_t)stream;
  long long v[2] = {0, 0};
  CU_CHECK(cudaMemcpyAsync(v, h->d_ilqr_stats, sizeof(v), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemsetAsync(h->d_ilqr_stats, 0, sizeof(v), st));
  CU_CHECK(cudaStreamSynchronize(st));
  *outer_iterations = v[0];
  *rollouts = v[1];
  return GMPC_OK;
}

extern "C" int gmpc_range_overflow(gmpc_handle* h, int32_t* count, void* stream) {
  if (!h || !count) return fail(GMPC_E_ARG, "gmpc_range_overflow: null argument");
  *count = 0;
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  uint32_t v[2] = {0, 0};
  uint32_t* src[2] = {h->h16.supported ? h->h16.d_ovf : nullptr, h->t128.supported ? h->t128.d_ovf : nullptr};
  for (int i = 0; i < 2; ++i)
    if (src[i]) CU_CHECK(cudaMemcpyAsync(&v[i], src[i], sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaStreamSynchronize(st));
  for (int i = 0; i < 2; ++i)
    if (src[i] && v[i] != 0) CU_CHECK(cudaMemsetAsync(src[i], 0, sizeof(uint32_t), st));
  *count = (int32_t)std::min<uint64_t>((uint64_t)v[0] + v[1], 0x7fffffffu);
  return GMPC_OK;
}

extern "C" int gmpc_plan_host(gmpc_handle* h, int64_t B, int32_t K, const float* x0_host,
                              const float* U0_host, const float* goal_host, int32_t method,
                              int32_t N, float lr, float b1, float b2, float eps,
                              float* U_best_host, float* X_best_host, float* J_best_host,
                              int32_t* idx_best_host, float* J_all_host, void* stream) {
  int rc = check_ready(h, "gmpc_plan_host", B);
  if (rc) return rc;
  if (!h->in_retry) h->auto_retry = (h->path == GMPC_PATH_AUTO);  // decided by the outermost call
  if (B == 0) return GMPC_OK;
  if (K < 1) return fail(GMPC_E_ARG, "gmpc_plan_host: K must be >= 1");
  if (!x0_host || !U0_host || !goal_host || !U_best_host || !X_best_host || !J_best_host ||
      !idx_best_host)
    return fail(GMPC_E_ARG, "gmpc_plan_host: null argument");
  cudaStream_t st = (cudaStream_t)stream;
  const gmpc_config& c = h->cfg;
  const size_t ub = (size_t)c.T * c.m, xb = (size_t)(c.T + 1) * c.n;
  const size_t f_x0 = (size_t)B * c.n, f_U0 = (size_t)B * K * ub, f_goal = (size_t)B * xb;
  const size_t f_Ub = (size_t)B * ub, f_Xb = (size_t)B * xb, f_Jb = B, f_idx = B, f_Ja = (size_t)B * K;
  const size_t tot = f_x0 + f_U0 + f_goal + f_Ub + f_Xb + f_Jb + f_idx + f_Ja;
  rc = grow(&h->d_stage, &h->stage_bytes, tot * sizeof(float));
  if (rc) return rc;
  float* d_x0 = (float*)h->d_stage;
  float* d_U0 = d_x0 + f_x0;
  float* d_goal = d_U0 + f_U0;
  float* d_Ub = d_goal + f_goal;
  float* d_Xb = d_Ub + f_Ub;
  float* d_Jb = d_Xb + f_Xb;
  int32_t* d_idx = (int32_t*)(d_Jb + f_Jb);
  float* d_Ja = (float*)(d_idx + f_idx);
  CU_CHECK(cudaMemcpyAsync(d_x0, x0_host, f_x0 * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_CHECK(cudaMemcpyAsync(d_U0, U0_host, f_U0 * sizeof(float), cudaMemcpyHostToDevice, st));
  CU_CHECK(cudaMemcpyAsync(d_goal, goal_host, f_goal * sizeof(float), cudaMemcpyHostToDevice, st));
  rc = gmpc_plan(h, B, K, d_x0, d_U0, d_goal, method, N, lr, b1, b2, eps, d_Ub, d_Xb, d_Jb, d_idx,
                 J_all_host ? d_Ja : nullptr, st);
  if (rc) return rc;
  CU_CHECK(cudaMemcpyAsync(U_best_host, d_Ub, f_Ub * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemcpyAsync(X_best_host, d_Xb, f_Xb * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemcpyAsync(J_best_host, d_Jb, f_Jb * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaMemcpyAsync(idx_best_host, d_idx, f_idx * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  if (J_all_host)
    CU_CHECK(cudaMemcpyAsync(J_all_host, d_Ja, f_Ja * sizeof(float), cudaMemcpyDeviceToHost, st));
  CU_CHECK(cudaStreamSynchronize(st));
  if (h->last_path == GMPC_PATH_TC16 || h->last_path == GMPC_PATH_TC16S || h->last_path == GMPC_PATH_T128) {
    int32_t clamped = 0;
    rc = gmpc_range_overflow(h, &clamped, stream);
    if (rc) return rc;
    if (clamped > 0) {
      // an operand left the fp16-split kernel's range: when the caller left the choice to us the plan
      // is re-done with forward rescaling and, should even that clamp, on the fp32 CUDA-core kernel
      // (still the GPU; there is no CPU path)
      if (!h->auto_retry)
        return fail(GMPC_E_UNSUPPORTED, "gmpc_plan_host: operand magnitude above 65000 on the fp16-split path; "
                                        "use GMPC_PATH_AUTO, GMPC_PATH_T128, GMPC_PATH_TC16S or GMPC_PATH_FFMA");
      const int retry = h->last_path == GMPC_PATH_TC16 ? GMPC_PATH_TC16S : GMPC_PATH_FFMA;
      h->path = retry;
      h->in_retry = true;
      rc = gmpc_plan_host(h, B, K, x0_host, U0_host, goal_host, method, N, lr, b1, b2, eps, U_best_host,
                          X_best_host, J_best_host, idx_best_host, J_all_host, stream);
      h->in_retry = false;
      h->path = GMPC_PATH_AUTO;
      return rc;
    }
  }
  return GMPC_OK;
}

// ----------------------------------------------------------------------------------- small contractions
// C[M,N] = alpha A B^T (+ C), A[M,R], B[N,R] row-major; rowsum_B[N] = alpha sum_r B[n,r] (nullable): the weight and
// bias gradients of the dynamics fit from gmpc_dynamics_fit's factors (norm/dynamics_trainer.py:64-79).
extern "C" int gmpc_gemm_nt(gmpc_handle* h, int32_t M, int32_t N, int64_t R, const float* A, const float* B, float alpha,
                            int32_t accumulate, float* C, float* rowsum_B, void* stream) {
  if (!h) return fail(GMPC_E_ARG, "gmpc_gemm_nt: null handle");
  if (M < 0 || N < 0 || R < 0 || !A || !B || !C) return fail(GMPC_E_ARG, "gmpc_gemm_nt: bad argument");
  if (M == 0 || N == 0) return GMPC_OK;
  cudaStream_t st = (cudaStream_t)stream;
  SGemm G;
  memset(&G, 0, sizeof(G));
  G.A = A; G.sam = R; G.sar = 1;
  G.B = B; G.sbr = 1; G.sbn = R;
  G.C = C; G.ldc = N; G.M = M; G.N = N; G.R = R; G.alpha = alpha; G.accumulate = accumulate;
  float* part = nullptr;
  if (const int sl = small_gemm_slices(M, N, R)) {   // long reduction, few output tiles: split it over the machine
    int rc = grow(&h->d_vjp_ws, &h->vjp_ws_bytes, sizeof(float) * (size_t)sl * M * N);
    if (rc) return rc;
    part = (float*)h->d_vjp_ws;
    ++h->launches;
  }
  CU_CHECK(small_gemm(G, SG_EPI_NONE, st, part));
  ++h->launches;
  if (rowsum_B) {
    strided_sum_kernel<<<(N + 7) / 8, 256, 0, st>>>(B, R, 1, N, R, alpha, rowsum_B);
    CU_CHECK(cudaGetLastError());
    ++h->launches;
  }
  return GMPC_OK;
}

// The cost-MLP part of cost_vjp (policy/optimizers.py:93-105), reduced over the batch:
//   gW[l], gb[l] = scale * sum_b grad_theta [ w2 d/de |f(x_T[b] + e dx_T[b]; theta)|^2 ],  w2 = sigmoid(mpc_weights[2]),
// f = the staged cost MLP (cost/nn.py:23-29).  phi = 2 f(x) . (Jf dx): the cotangent 2 w2 (Jf dx) goes back through
// the primal network (activations a_l), the cotangent 2 w2 f(x) through the tangent network (da_l, linear, no bias):
//   gW_l = a_l^T c_l + da_l^T d_l,  gb_l = colsum(c_l).   gW[l] is [in_l, out_l] (flax kernel layout).
extern "C" int gmpc_cost_mixed_vjp(gmpc_handle* h, int64_t B, const float* xT, const float* dxT, float scale,
                                   float* const* gW, float* const* gb, void* stream) {
  int rc = check_ready(h, "gmpc_cost_mixed_vjp", B);
  if (rc) return rc;
  if (!gW || !gb || (B > 0 && (!xT || !dxT))) return fail(GMPC_E_ARG, "gmpc_cost_mixed_vjp: null argument");
  if (B > (int64_t)1 << 24) return fail(GMPC_E_ARG, "gmpc_cost_mixed_vjp: batch too large");
  const MlpPack& mp = h->cost;
  const int L = mp.L;
  cudaStream_t st = (cudaStream_t)stream;
  if (B == 0) {
    for (int l = 0; l < L; ++l) {
      CU_CHECK(cudaMemsetAsync(gW[l], 0, sizeof(float) * mp.dims[l] * mp.dims[l + 1], st));
      CU_CHECK(cudaMemsetAsync(gb[l], 0, sizeof(float) * mp.dims[l + 1], st));
    }
    return GMPC_OK;
  }
  int dmax = 0;
  size_t hid = 0;
  for (int l = 0; l <= L; ++l) dmax = std::max(dmax, mp.dims[l]);
  for (int l = 1; l < L; ++l) hid += mp.dims[l];
  // workspace: a_l, da_l, mask_l for the hidden layers; y, dy; two pairs of cotangent buffers
  const size_t need = sizeof(float) * (size_t)B * (3 * hid + 2 * (size_t)mp.dims[L] + 4 * (size_t)dmax);
  rc = grow(&h->d_vjp_ws, &h->vjp_ws_bytes, need);
  if (rc) return rc;
  float* w = (float*)h->d_vjp_ws;
  const float* a[MAXL + 1];
  const float* da[MAXL + 1];
  float* mk[MAXL + 1];
  a[0] = xT; da[0] = dxT; mk[0] = nullptr;
  for (int l = 1; l < L; ++l) {
    float* al = w; w += (size_t)B * mp.dims[l];
    float* dl = w; w += (size_t)B * mp.dims[l];
    mk[l] = w; w += (size_t)B * mp.dims[l];
    SGemm G;
    memset(&G, 0, sizeof(G));
    G.A = a[l - 1]; G.sam = mp.dims[l - 1]; G.sar = 1;
    G.B = mp.Wf[l - 1]; G.sbr = mp.ldf[l - 1]; G.sbn = 1;
    G.C = al; G.ldc = mp.dims[l]; G.M = (int)B; G.N = mp.dims[l]; G.R = mp.dims[l - 1]; G.alpha = 1.f;
    G.bias = mp.bias[l - 1]; G.mask = mk[l];
    CU_CHECK(small_gemm(G, SG_EPI_BIAS_RELU_MASK, st));
    G.A = da[l - 1]; G.C = dl; G.bias = nullptr;
    CU_CHECK(small_gemm(G, SG_EPI_MASK, st));
    a[l] = al; da[l] = dl;
    h->launches += 2;
  }
  float* y = w; w += (size_t)B * mp.dims[L];
  float* dy = w; w += (size_t)B * mp.dims[L];
  float* cbuf[2] = {w, w + (size_t)B * dmax};
  float* dbuf[2] = {w + 2 * (size_t)B * dmax, w + 3 * (size_t)B * dmax};
  {
    SGemm G;
    memset(&G, 0, sizeof(G));
    G.A = a[L - 1]; G.sam = mp.dims[L - 1]; G.sar = 1;
    G.B = mp.Wf[L - 1]; G.sbr = mp.ldf[L - 1]; G.sbn = 1;
    G.C = y; G.ldc = mp.dims[L]; G.M = (int)B; G.N = mp.dims[L]; G.R = mp.dims[L - 1]; G.alpha = 1.f;
    G.bias = mp.bias[L - 1];
    CU_CHECK(small_gemm(G, SG_EPI_BIAS, st));
    G.A = da[L - 1]; G.C = dy; G.bias = nullptr;
    CU_CHECK(small_gemm(G, SG_EPI_NONE, st));
    const long long cnt = (long long)B * mp.dims[L];
    mixed_vjp_seed_kernel<<<(int)std::min<long long>((cnt + 255) / 256, 1024), 256, 0, st>>>(y, dy, h->d_mpcw, scale, cnt,
                                                                                              cbuf[0], dbuf[0]);
    CU_CHECK(cudaGetLastError());
    h->launches += 3;
  }
  int cur = 0;
  for (int l = L - 1; l >= 0; --l) {
    const int Ki = mp.dims[l], No = mp.dims[l + 1];
    SGemm G;
    memset(&G, 0, sizeof(G));
    // gW_l [Ki, No] = a_l^T c + da_l^T d   (reduction over the batch)
    G.A = a[l]; G.sam = 1; G.sar = Ki;
    G.B = cbuf[cur]; G.sbr = No; G.sbn = 1;
    G.C = gW[l]; G.ldc = No; G.M = Ki; G.N = No; G.R = B; G.alpha = 1.f;
    CU_CHECK(small_gemm(G, SG_EPI_NONE, st));
    G.A = da[l]; G.B = dbuf[cur]; G.accumulate = 1;
    CU_CHECK(small_gemm(G, SG_EPI_NONE, st));
    strided_sum_kernel<<<(No + 7) / 8, 256, 0, st>>>(cbuf[cur], 1, No, No, B, 1.f, gb[l]);
    CU_CHECK(cudaGetLastError());
    h->launches += 3;
    if (l > 0) {
      // c' = (c W_l^T) * mask_l, the same for d: W_l^T(r = output, n = input) = Wb[l][r * ldb + n]
      memset(&G, 0, sizeof(G));
      G.A = cbuf[cur]; G.sam = No; G.sar = 1;
      G.B = mp.Wb[l]; G.sbr = mp.ldb[l]; G.sbn = 1;
      G.C = cbuf[cur ^ 1]; G.ldc = Ki; G.M = (int)B; G.N = Ki; G.R = No; G.alpha = 1.f;
      G.mask = mk[l];
      CU_CHECK(small_gemm(G, SG_EPI_MASK, st));
      G.A = dbuf[cur]; G.C = dbuf[cur ^ 1];
      CU_CHECK(small_gemm(G, SG_EPI_MASK, st));
      h->launches += 2;
      cur ^= 1;
    }
  }
  return GMPC_OK;
}

// ----------------------------------------------------------------------------------- dynamics fit
extern "C" int64_t gmpc_dynamics_fit_columns(const gmpc_handle* h, int64_t B, int32_t S) {
  if (!h || B < 0 || S < 1) return -1;
  return (int64_t)RT * ((B + RT - 1) / RT) * S;
}

// predict_loss (norm/dynamics_trainer.py:13-44) for a batch of windows and the factors of its weight
// gradient (csrc/dynfit.cuh).
extern "C" int gmpc_dynamics_fit(gmpc_handle* h, int64_t B, int32_t S, const float* xseq,
                                 const float* useq, const float* next_xseq, float discount_factor,
                                 int32_t teacher_forcing, float* loss, float* const* act,
                                 float* const* cot, void* stream) {
  int rc = check_ready(h, "gmpc_dynamics_fit", B);
  if (rc) return rc;
  if (B == 0) return GMPC_OK;
  if (S < 1 || !xseq || !useq || !next_xseq || !loss || !act || !cot)
    return fail(GMPC_E_ARG, "gmpc_dynamics_fit: null argument or S < 1");
  const gmpc_config& c = h->cfg;
  cudaStream_t st = (cudaStream_t)stream;
  DynFitParams Q;
  memset(&Q, 0, sizeof(Q));
  PlanParams& P = Q.pp;
  make_dirs(h->dyn, P.dir[DIR_DYN_F], P.dir[DIR_DYN_B]);
  P.n = c.n; P.m = c.m; P.T = c.T; P.K = 1;
  P.hpad = h->hpad;
  P.iters = S;     // SCHED_FIT: S forward passes then S adjoint passes
  P.NQ = B;
  P.ntiles = (int)((B + RT - 1) / RT);
  const int L = h->dyn.L;
  for (int l = 0; l < L; ++l) {
    if (!act[l] || !cot[l]) return fail(GMPC_E_ARG, "gmpc_dynamics_fit: null act/cot pointer");
    Q.act[l] = act[l];
    Q.cot[l] = cot[l];
  }
  Q.S = S;
  Q.teacher_forcing = teacher_forcing ? 1 : 0;
  Q.gamma = discount_factor;
  Q.xseq = xseq; Q.useq = useq; Q.yseq = next_xseq;
  Q.loss = loss;
  Q.R = (long long)RT * P.ntiles * S;
  const int grid = std::min(P.ntiles, h->num_sms);
  const size_t need = (size_t)grid * S * std::max(1, L - 1) * h->maxt * NTHREADS * sizeof(uint32_t);
  if (need > h->fit_masks_bytes) {
    CU_CHECK(cudaStreamSynchronize(st));
    rc = grow((void**)&h->d_fit_masks, &h->fit_masks_bytes, need);
    if (rc) return rc;
  }
  Q.masks = h->d_fit_masks;
  const int n4 = rup4(c.n), nm4 = rup4(c.n + c.m);
  const size_t smem = sizeof(float) * ((size_t)2 * h->hpad * RT + (size_t)NSTAGE * STAGE_FLOATS +
                                       (size_t)(2 * nm4 + 2 * n4) * RT);
  if (!h->fit_attr) {
    CU_CHECK(cudaFuncSetAttribute(dynfit_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    CU_CHECK(cudaFuncSetAttribute(dynfit_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    h->fit_attr = true;
  }
  if (smem > 227 * 1024) return fail(GMPC_E_UNSUPPORTED, "gmpc_dynamics_fit: shape needs more shared memory than one SM has");
  if (h->maxt == 1)
    dynfit_kernel<1><<<grid, NTHREADS, smem, st>>>(Q);
  else
    dynfit_kernel<2><<<grid, NTHREADS, smem, st>>>(Q);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

// ----------------------------------------------------------------------------------- expert
// Flat parameter layout of the expert proposal network (expert/nn.py): trunk, then the next_x head,
// then the action head; every kernel is flax [in, out] row-major.
static int expert_dims(int n, int m, int T, int hist, int F, int num_layers, int H, ExpertDims& d) {
  memset(&d, 0, sizeof(d));
  d.n = n; d.m = m; d.T = T; d.hist = hist; d.F = F; d.H = H;
  d.Lh = F > 0 ? num_layers : num_layers - 1;
  d.Y = F > 0 ? F : H;
  if (d.Lh < 1 || d.Lh > MAXL) return -1;
  long long o = 0;
  if (F > 0) {
    d.oWi = o; o += (long long)n * 4 * F;
    d.oWh = o; o += (long long)F * 4 * F;
    d.obh = o; o += 4 * F;
  } else {
    d.oD0 = o; o += (long long)n * H;
    d.ob0 = o; o += H;
  }
  for (int hd = 0; hd < 2; ++hd) {
    int din = d.Y;
    for (int l = 0; l < d.Lh; ++l) {
      const int dout = l == d.Lh - 1 ? (hd == 0 ? n : m) : H;
      d.oHk[hd][l] = o; o += (long long)din * dout;
      d.oHb[hd][l] = o; o += dout;
      din = H;
    }
  }
  d.P = o;
  return 0;
}

extern "C" int64_t gmpc_expert_param_count(int32_t n, int32_t m, int32_t lstm_features,
                                           int32_t num_layers, int32_t num_hidden_units) {
  ExpertDims d;
  if (n < 1 || m < 1 || lstm_features < 0 || num_hidden_units < 1 ||
      expert_dims(n, m, 1, 0, lstm_features, num_layers, num_hidden_units, d))
    return -1;
  return d.P;
}

extern "C" int gmpc_expert_propose(gmpc_handle* h, int64_t B, int32_t hist, const float* history_x,
                                   const float* params_flat, int32_t lstm_features,
                                   int32_t num_layers, int32_t num_hidden_units, float* goal_xseq,
                                   float* init_useq, void* stream) {
  if (!h) return fail(GMPC_E_ARG, "gmpc_expert_propose: null handle");
  if (B < 0 || hist < 0) return fail(GMPC_E_ARG, "gmpc_expert_propose: need B >= 0, hist >= 0");
  if (B == 0) return GMPC_OK;
  if (!history_x || !params_flat || !goal_xseq || !init_useq)
    return fail(GMPC_E_ARG, "gmpc_expert_propose: null argument");
  const gmpc_config& c = h->cfg;
  ExpertDims d;
  if (lstm_features < 0 || num_hidden_units < 1 ||
      expert_dims(c.n, c.m, c.T, hist, lstm_features, num_layers, num_hidden_units, d))
    return fail(GMPC_E_ARG, "gmpc_expert_propose: bad network shape");
  const int width = std::max(std::max(4 * d.F, d.H), std::max(c.n, c.m));
  if (width > 1024) return fail(GMPC_E_UNSUPPORTED, "gmpc_expert_propose: needs 4*lstm_features, hidden, n, m <= 1024");
  CU_CHECK(cudaSetDevice(c.device));
  const int threads = std::max(64, ((width + 31) / 32) * 32);
  const size_t smem = sizeof(float) * ((size_t)2 * c.n + (size_t)6 * d.F + d.Y + 2 * d.H);
  const int grid = (int)std::min<int64_t>(B, (int64_t)4 * h->num_sms);
  expert_kernel<<<grid, threads, smem, (cudaStream_t)stream>>>(d, history_x, params_flat, B, goal_xseq, init_useq);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

// ----------------------------------------------------------------------------------- critic
static size_t critic_smem(const CriticDims& d, int T1) {
  const int G = 4 * d.F, W = std::max(d.F, d.H);
  return sizeof(float) * ((size_t)T1 * d.n + (size_t)T1 * G + (size_t)2 * (T1 + 1) * d.F + 2 * d.F +
                          (size_t)d.L * W + 2 * W + 4);
}

static int critic_launch(gmpc_handle* h, const char* who, int64_t Bc, int32_t T1,
                         const float* xseq, const float* label, const int32_t* perm,
                         const float* params, float inv_count, float* loss, float* grad,
                         float* logits, cudaStream_t st, float* dx_out = nullptr) {
  if (!h) return fail(GMPC_E_ARG, std::string(who) + ": null handle");
  if (h->cfg.critic_features <= 0) return fail(GMPC_E_STATE, std::string(who) + ": handle was created without a critic");
  if (Bc < 0 || T1 < 1) return fail(GMPC_E_ARG, std::string(who) + ": need Bc >= 0, T1 >= 1");
  if (Bc == 0) return GMPC_OK;
  if (!xseq || !params || (!label && (loss || grad)))
    return fail(GMPC_E_ARG, std::string(who) + ": null argument");
  CU_CHECK(cudaSetDevice(h->cfg.device));
  CriticDims d = h->cd;
  d.T1 = T1;
  const size_t smem = critic_smem(d, T1);
  if (smem > 227 * 1024) return fail(GMPC_E_UNSUPPORTED, std::string(who) + ": sequence too long for one SM's shared memory");
  if ((size_t)Bc > h->losses_cap) {
    if (h->d_losses) cudaFree(h->d_losses);
    h->d_losses = nullptr; h->losses_cap = 0;
    CU_CHECK(cudaMalloc(&h->d_losses, sizeof(float) * (size_t)Bc * 2));
    h->losses_cap = (size_t)Bc;
  }
  const int threads = std::max(64, ((std::max(4 * d.F, d.H) + 31) / 32) * 32);
  const int grid = (int)std::min<int64_t>(Bc, h->critic_grid);
  // labels are unused when only logits are requested: feed the losses buffer as a dummy
  const float* lab = label ? label : h->d_losses + h->losses_cap;
  if (!label) CU_CHECK(cudaMemsetAsync(h->d_losses + h->losses_cap, 0, sizeof(float) * (size_t)Bc, st));
  critic_kernel<<<grid, threads, smem, st>>>(d, xseq, lab, perm, params, inv_count, Bc,
                                             h->d_losses, logits, h->d_partial, (grad || dx_out) ? 1 : 0,
                                             dx_out);
  ++h->launches;
  if (loss || grad) {
    const int rb = grad ? (int)((d.P + 255) / 256) : 1;
    critic_reduce_kernel<<<rb, 256, 0, st>>>(h->d_partial, grid, d.P, grad, h->d_losses, Bc,
                                             inv_count, loss);
    ++h->launches;
  }
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

extern "C" int gmpc_critic_forward(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                                   const float* params_flat, float* logit, void* stream) {
  if (!logit) return fail(GMPC_E_ARG, "gmpc_critic_forward: null argument");
  return critic_launch(h, "gmpc_critic_forward", Bc, T1, xseq, nullptr, nullptr, params_flat, 0.f,
                       nullptr, nullptr, logit, (cudaStream_t)stream);
}

extern "C" int gmpc_critic_input_grad(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                                      const float* params_flat, float* logit, float* dxseq,
                                      void* stream) {
  if (!logit || !dxseq) return fail(GMPC_E_ARG, "gmpc_critic_input_grad: null argument");
  return critic_launch(h, "gmpc_critic_input_grad", Bc, T1, xseq, nullptr, nullptr, params_flat, 0.f,
                       nullptr, nullptr, logit, (cudaStream_t)stream, dxseq);
}

extern "C" int gmpc_critic_loss_grad(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                                     const float* label, const float* params_flat,
                                     float inv_count, float* loss, float* grad_flat, void* stream) {
  if (!loss) return fail(GMPC_E_ARG, "gmpc_critic_loss_grad: null loss pointer");
  return critic_launch(h, "gmpc_critic_loss_grad", Bc, T1, xseq, label, nullptr, params_flat,
                       inv_count, loss, grad_flat, nullptr, (cudaStream_t)stream);
}

extern "C" int gmpc_critic_loss_grad_gather(gmpc_handle* h, int64_t Bc, int32_t T1,
                                            const float* data_xseq, const float* data_label,
                                            const int32_t* perm, const float* params_flat,
                                            float inv_count, float* loss, float* grad_flat,
                                            void* stream) {
  if (!loss || !perm) return fail(GMPC_E_ARG, "gmpc_critic_loss_grad_gather: null argument");
  return critic_launch(h, "gmpc_critic_loss_grad_gather", Bc, T1, data_xseq, data_label, perm,
                       params_flat, inv_count, loss, grad_flat, nullptr, (cudaStream_t)stream);
}

extern "C" int gmpc_clip_adam_step(gmpc_handle* h, int64_t P, float* params_flat,
                                   const float* grad_flat, float* mom, float* vel, int32_t step,
                                   float lr, float max_norm, float grad_scale, float b1, float b2,
                                   float eps, void* stream) {
  if (!h || !params_flat || !grad_flat || !mom || !vel || P < 0 || step < 1)
    return fail(GMPC_E_ARG, "gmpc_clip_adam_step: bad argument");
  if (P == 0) return GMPC_OK;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
  const float bc2 = (float)(1.0 - pow((double)b2, (double)step));
  clip_adam_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(P, params_flat, grad_flat, mom, vel, bc1, bc2,
                                                         lr, max_norm, grad_scale, b1, b2, eps);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

extern "C" int gmpc_critic_train_scan(gmpc_handle* h, int32_t steps, int64_t Bc, int32_t T1,
                                      const float* data_xseq, const float* data_label,
                                      const int32_t* perm, float* params_flat, float* mom, float* vel,
                                      int32_t step0, float lr, float max_norm, float b1, float b2,
                                      float eps, float* losses, float* grad_scratch, void* stream) {
  if (!h || steps < 0 || Bc < 1 || step0 < 0 || !perm || !params_flat || !mom || !vel || !losses ||
      !grad_scratch || !data_xseq || !data_label)
    return fail(GMPC_E_ARG, "gmpc_critic_train_scan: bad argument");
  if (h->cfg.critic_features <= 0) return fail(GMPC_E_STATE, "gmpc_critic_train_scan: handle was created without a critic");
  cudaStream_t st = (cudaStream_t)stream;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  CriticDims d = h->cd;
  d.T1 = T1;
  const size_t smem = critic_smem(d, T1);
  if (smem > 227 * 1024) return fail(GMPC_E_UNSUPPORTED, "gmpc_critic_train_scan: sequence too long for one SM's shared memory");
  if ((size_t)Bc > h->losses_cap) {
    if (h->d_losses) cudaFree(h->d_losses);
    h->d_losses = nullptr; h->losses_cap = 0;
    CU_CHECK(cudaMalloc(&h->d_losses, sizeof(float) * (size_t)Bc * 2));
    h->losses_cap = (size_t)Bc;
  }
  const int rb = (int)((d.P + 255) / 256);
  if (rb > h->num_sms) return fail(GMPC_E_UNSUPPORTED, "gmpc_critic_train_scan: critic too large for the fused tail kernel");
  if (!h->d_fuse) {
    CU_CHECK(cudaMalloc(&h->d_fuse, sizeof(float) * (size_t)(rb + 2)));
    CU_CHECK(cudaMemsetAsync(h->d_fuse, 0, sizeof(float) * (size_t)(rb + 2), st));
    h->fuse_tickets = 0;
  }
  const int threads = std::max(64, ((std::max(4 * d.F, d.H) + 31) / 32) * 32);
  const int grid = (int)std::min<int64_t>(Bc, h->critic_grid);
  const float inv_count = 1.f / (float)Bc;
  for (int32_t s = 0; s < steps; ++s) {
    critic_kernel<<<grid, threads, smem, st>>>(d, data_xseq, data_label, perm + (size_t)s * Bc, params_flat,
                                               inv_count, Bc, h->d_losses, nullptr, h->d_partial, 1);
    const int step = step0 + s + 1;
    const float bc1 = (float)(1.0 - pow((double)b1, (double)step));
    const float bc2 = (float)(1.0 - pow((double)b2, (double)step));
    critic_reduce_clip_adam_kernel<<<rb, 256, 0, st>>>(
        h->d_partial, grid, d.P, grad_scratch, h->d_losses, Bc, inv_count, losses + s, params_flat, mom, vel,
        lr, max_norm, b1, b2, eps, bc1, bc2, h->d_fuse + 2, reinterpret_cast<unsigned long long*>(h->d_fuse),
        h->fuse_tickets += (unsigned long long)rb);
    if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
      // the tail kernel of this step was not enqueued: take its tickets back, or the next scan would wait for them
      h->fuse_tickets -= (unsigned long long)rb;
      return fail(GMPC_E_CUDA, std::string("gmpc_critic_train_scan: launch failed: ") + cudaGetErrorString(e));
    }
    h->launches += 2;
  }
  return GMPC_OK;
}

extern "C" int gmpc_l2_loss(gmpc_handle* h, int64_t B, const float* X, const float* desired,
                            float* loss, void* stream) {
  if (!h || !X || !desired || !loss || B < 0) return fail(GMPC_E_ARG, "gmpc_l2_loss: bad argument");
  if (B == 0) return GMPC_OK;
  CU_CHECK(cudaSetDevice(h->cfg.device));
  const int wpb = 4;
  l2_loss_kernel<<<(unsigned)((B + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
      B, h->cfg.T + 1, h->cfg.n, X, desired, loss);
  ++h->launches;
  CU_CHECK(cudaGetLastError());
  return GMPC_OK;
}

// ----------------------------------------------------------------------------------- diagnostics
extern "C" int gmpc_measure_fp32_peak(int device, float* tflops_out) {
  if (!tflops_out) return fail(GMPC_E_ARG, "gmpc_measure_fp32_peak: null argument");
  CU_CHECK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU_CHECK(cudaGetDeviceProperties(&prop, device));
  float* d = nullptr;
  const int grid = prop.multiProcessorCount * 8, iters = 4096;
  CU_CHECK(cudaMalloc(&d, sizeof(float) * grid * 256));
  cudaEvent_t e0, e1;
  CU_CHECK(cudaEventCreate(&e0));
  CU_CHECK(cudaEventCreate(&e1));
  float best = 0.f;
  for (int rep = 0; rep < 5; ++rep) {
    CU_CHECK(cudaEventRecord(e0));
    ffma_peak_kernel<<<grid, 256>>>(d, iters, 1.0f + rep);
    CU_CHECK(cudaEventRecord(e1));
    CU_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 16 * 8 * (double)iters * 256.0 * grid;
    if (rep > 0) best = fmaxf(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  *tflops_out = best;
  return GMPC_OK;
}

// Dense kind::f16 tcgen05 peak (TFLOP/s) of `device`: the tensor pipe the planner kernels run on, measured
// with operands resident in shared memory (roofline context beside the cuBLAS figure of MEASURED_PEAKS.json).
extern "C" int gmpc_measure_f16_mma_peak(int device, float* tflops_out) {
  if (!tflops_out) return fail(GMPC_E_ARG, "gmpc_measure_f16_mma_peak: null argument");
  CU_CHECK(cudaSetDevice(device));
  int sms = 0;
  CU_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
  CU_CHECK(cudaFuncSetAttribute(f16_mma_peak_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 48 * 1024));
  cudaEvent_t e0, e1;
  CU_CHECK(cudaEventCreate(&e0));
  CU_CHECK(cudaEventCreate(&e1));
  const int mmas = 200000;
  float best = 0.f;
  for (int rep = 0; rep < 4; ++rep) {
    CU_CHECK(cudaEventRecord(e0));
    f16_mma_peak_kernel<<<sms, 128, 48 * 1024>>>(mmas);
    CU_CHECK(cudaEventRecord(e1));
    CU_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CU_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double flops = 2.0 * 128 * 256 * 16 * (double)mmas * sms;
    if (rep > 0) best = fmaxf(best, (float)(flops / (ms * 1e-3) / 1e12));
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  *tflops_out = best;
  return GMPC_OK;
}
