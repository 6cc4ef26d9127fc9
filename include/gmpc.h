/* gmpc.h -- C ABI of libgmpc.so: the B200-native gan_mpc planning hot path.
 *
 * This is the drop-in boundary.  The reference (returaj/gan_mpc) has no FFI; its boundary for
 * this path is the Python API of gan_mpc.policy.optimizers + the policy classes.  Each entry
 * point below cites the reference interface it replaces (paths relative to the reference root).
 * The Python mirror of that API (gan_mpc_b200/policy/..., gan/..., norm/...) binds these with
 * ctypes; see INTEGRATION.md for the stub a maintainer would add on the reference side.
 *
 * Conventions
 *  - every call returns int: 0 ok, <0 error (GMPC_E_*); message via gmpc_last_error().
 *  - all array arguments are DEVICE pointers to fp32 row-major data unless the name ends in
 *    _host or the comment says otherwise.  The caller owns every buffer.  The library keeps no
 *    pointer past the call (weights are copied/packed into the handle by gmpc_set_weights).
 *  - work is enqueued on the caller's stream (a cudaStream_t passed as void*) and is
 *    asynchronous; no hidden host sync except in the *_host convenience entry points.
 *  - a handle is bound to one device and is not thread-safe (one handle per stream/GPU).
 *  - weights use the flax layout kernel[in, out] (y = x @ kernel + bias), as stored in the
 *    reference's params pytree ({"params": {"Dense_i": {"kernel", "bias"}}}).
 */
#ifndef GMPC_H_
#define GMPC_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GMPC_OK 0
#define GMPC_E_ARG (-1)         /* bad argument / null pointer */
#define GMPC_E_UNSUPPORTED (-2) /* shape outside what the kernels support */
#define GMPC_E_CUDA (-3)        /* CUDA runtime error */
#define GMPC_E_STATE (-4)       /* call order (e.g. plan before set_weights) */

#define GMPC_METHOD_GRAD 0 /* U <- U - lr g */
#define GMPC_METHOD_ADAM 1 /* optax adam(lr, b1, b2, eps), per-trajectory state */

#define GMPC_MAX_LAYERS 8

/* kernel family used for the MLP contractions */
#define GMPC_PATH_AUTO 0 /* tcgen05 (T128, else TC16S) when the tile is a real dense contraction (>= 64 trajectories,
                            hidden >= 64), else FFMA */
#define GMPC_PATH_FFMA 1 /* fp32 CUDA-core path (exact fp32 FMA arithmetic) */
#define GMPC_PATH_TC 2   /* (retired) the first, 3xTF32 tensor-core kernel: gmpc_set_path answers GMPC_E_UNSUPPORTED */
#define GMPC_PATH_TC16 3 /* tcgen05 fp16-split (hi/lo, 3 products, fp32 accumulate) path, pipelined */
#define GMPC_PATH_TC16S 4 /* same, with per-trajectory power-of-two scaling of the forward operands too
                            (states of any magnitude; ~6 % slower).  Hidden widths above 256 always use it. */
#define GMPC_PATH_T128 5  /* tcgen05 fp16-split, 128-trajectory tiles: trajectories are the M dimension, the
                            activations live in tensor memory as the A operand, weights stream through shared
                            memory as B; forward and adjoint operands rescaled per trajectory (hidden <= 256) */

typedef struct gmpc_config {
  int32_t n;             /* state size x_size                       (dynamics/nn.py:13 x_out) */
  int32_t m;             /* action size u_size */
  int32_t T;             /* MPC horizon            (config mpc.horizon, cost/cost_model.py:35) */
  int32_t dyn_layers;    /* dynamics MLP num_layers L: L-1 hidden Dense+ReLU, then Dense(n) */
  int32_t dyn_hidden;    /* dynamics MLP num_hidden_units */
  int32_t cost_layers;   /* cost MLP num_layers                          (cost/nn.py:11-13) */
  int32_t cost_hidden;   /* cost MLP num_hidden_units */
  int32_t cost_fout;     /* cost MLP fout */
  int32_t critic_features; /* critic LSTM features F (0 = no critic)    (critic/nn.py:11-14) */
  int32_t critic_layers;   /* critic num_layers (num_layers-1 hidden Dense, then Dense(1)) */
  int32_t critic_hidden;   /* critic num_hidden_units */
  int32_t device;        /* CUDA device ordinal */
} gmpc_config;

typedef struct gmpc_handle gmpc_handle;

/* Create/destroy.  Replaces constructing EvalMPC/BaseMPC (policy/eval.py:26-39) for the
 * structured DynamicsModel(MLP) + MujocoBasedModel(cost MLP) (+ CriticModel(LSTM)) case. */
int gmpc_create(const gmpc_config* cfg, gmpc_handle** out);
int gmpc_destroy(gmpc_handle* h);

/* Thread-local message for the last non-zero return. */
const char* gmpc_last_error(void);

/* Number of floats of the flat critic parameter vector for this handle's config
 * (layout: Wi[n,4F] | Wh[F,4F] | bh[4F] | {Dk[in,H], Db[H]}* | Wo[in,1] | bo[1]; gates i,f,g,o). */
int64_t gmpc_critic_param_count(const gmpc_handle* h);

/* Select the contraction path (GMPC_PATH_*).  Default AUTO. */
int gmpc_set_path(gmpc_handle* h, int path);
/* Which path the last plan/objective call actually used (GMPC_PATH_FFMA, _TC16, _TC16S or _T128). */
int gmpc_last_path(const gmpc_handle* h);

/* Stage model weights (copied and re-packed inside the handle; safe to free after return of
 * the stream work).  dyn_W[l] is kernel[in_l, out_l], dyn_b[l] is bias[out_l]; l < dyn_layers.
 * mpc_weights is the raw (pre-sigmoid) [action, state, terminal] vector, DEVICE pointer.
 * Replaces reading params["dynamics_params"|"cost_params"|"mpc_weights"]
 * (policy/eval.py:64-73). */
int gmpc_set_weights(gmpc_handle* h, const float* const* dyn_W, const float* const* dyn_b,
                     const float* const* cost_W, const float* const* cost_b,
                     const float* mpc_weights, void* stream);

/* trajax rollout as used at policy/optimizers.py:28,80:  X[b,0]=x0[b]; X[b,t+1]=dyn(X[b,t],U[b,t]).
 * x0[B,n], U[B,T,m] -> X[B,T+1,n]. */
int gmpc_rollout(gmpc_handle* h, int64_t B, const float* x0, const float* U, float* X,
                 void* stream);

/* objective (policy/optimizers.py:24-31) and its action gradient (jax.grad at :83,:103; trajax
 * ilqr's `gradient`/`adjoints` outputs consumed at :55).
 * x0[B,n], U[B,T,m], goal[B,T+1,n] -> J[B], dU[B,T,m] (nullable), X[B,T+1,n] (nullable),
 * lam[B,T+1,n] (nullable adjoints). */
int gmpc_objective_grad(gmpc_handle* h, int64_t B, const float* x0, const float* U,
                        const float* goal, float* J, float* dU, float* X, float* lam,
                        void* stream);

/* loss_grad_wrt_control (policy/optimizers.py:78-83) for loss = L2MPC.loss
 * (norm/l2_policy.py:12-18): x0[B,n], U[B,T,m], desired[B,T+1,n] -> loss[B], dU[B,T,m] (nullable),
 * X[B,T+1,n] (nullable). */
int gmpc_l2_loss_grad(gmpc_handle* h, int64_t B, const float* x0, const float* U,
                      const float* desired, float* loss, float* dU, float* X, void* stream);

/* The fused planner: replaces ilqr_solve (policy/optimizers.py:10-21) /
 * EvalMPC.get_optimal_values (policy/eval.py:109-124) with the north-star first-order planner
 * on the reference's exact objective.  For every start state b and candidate k: N iterations of
 * {rollout, cost, adjoint, update}, final evaluation, then idx = argmin_k J (first minimum).
 * x0[B,n], U0[B,K,T,m], goal[B,T+1,n] ->
 * U_best[B,T,m], X_best[B,T+1,n], J_best[B], idx_best[B] (int32), J_all[B,K] (nullable). */
int gmpc_plan(gmpc_handle* h, int64_t B, int32_t K, const float* x0, const float* U0,
              const float* goal, int32_t method, int32_t N, float lr, float b1, float b2,
              float eps, float* U_best, float* X_best, float* J_best, int32_t* idx_best,
              float* J_all, void* stream);

/* Same as gmpc_plan but every array argument is a HOST pointer (pinned or pageable): copies
 * inputs host->device, plans, copies results back and synchronises `stream` before returning.
 * This is the end-to-end call a non-GPU-aware reference caller would bind. */
int gmpc_plan_host(gmpc_handle* h, int64_t B, int32_t K, const float* x0_host,
                   const float* U0_host, const float* goal_host, int32_t method, int32_t N,
                   float lr, float b1, float b2, float eps, float* U_best_host,
                   float* X_best_host, float* J_best_host, int32_t* idx_best_host,
                   float* J_all_host, void* stream);

/* trajax iLQR options as the reference passes them (policy/eval.py:10-20, TRAJAX_iLQR_KWARGS).
 * The thresholds the reference leaves at 0.0 and make_psd = False are not configurable. */
typedef struct gmpc_ilqr_options {
  int32_t maxiter;             /* 100 */
  float grad_norm_threshold;   /* 1e-4 */
  float alpha_0;               /* 1.0 */
  float alpha_min;             /* 0.00005 */
  int32_t gradient_lag;        /* 0 (default): `gradient` / `adjoints` and the continuation test belong to the RETURNED
                                * trajectory.  1: trajax's loop body as two independent recollections of the source have it
                                * (`gradient, adjoints = adjoint(A, B, q, r)` with the lqr tuple unpacked BEFORE the step,
                                * then `lqr = get_lqr_params(X, U)`): the returned gradient / adjoints and the
                                * `grad_norm_threshold` test lag one iterate, so a converging solve runs one more
                                * iteration.  trajax@c94a637 is not on disk, hence an option and not the default. */
} gmpc_ilqr_options;

/* The reference's OWN planner step: ilqr_solve (policy/optimizers.py:10-21) = trajax.optimizers.ilqr
 * on cost = EvalMPC.cost (policy/eval.py:64-69) and dynamics = EvalMPC.dynamics (:71-73), batched
 * with vmap semantics (every trajectory gets what its unbatched call would return).  One fused
 * kernel per call: rollout, linearisation (dynamics Jacobians by n adjoint passes through the ReLU
 * masks, closed-form cost Hessians), Riccati sweep with the 1e-8 eigenvalue floor, backtracking
 * line search of feedback rollouts, trajax's continuation test.  fp32 CUDA cores.
 * x0[B,n], U0[B,T,m], goal[B,T+1,n] ->
 * X[B,T+1,n], U[B,T,m], obj[B], gradient[B,T,m] (nullable), adjoints[B,T+1,n] (nullable),
 * iteration[B] (int32, nullable), lqr_A[B,T,n,n] and lqr_B[B,T,n,m] (nullable: the dynamics
 * Jacobians at the returned trajectory, elements 5 and 6 of trajax's `lqr` tuple) -- the 7-tuple
 * unpacked at policy/optimizers.py:55.  Limits: m <= 16 and n small enough for the Riccati
 * matrices of 32 trajectories to fit one SM's shared memory (n <= 24 at m = 6). */
int gmpc_ilqr(gmpc_handle* h, int64_t B, const float* x0, const float* U0, const float* goal,
              const gmpc_ilqr_options* opt, float* X, float* U, float* obj, float* gradient,
              float* adjoints, int32_t* iteration, float* lqr_A, float* lqr_B, void* stream);

/* Same as gmpc_ilqr with HOST pointers everywhere (pinned or pageable): copies in, solves, copies out
 * and synchronises `stream` -- the call a non-GPU-aware acting loop binds (EvalMPC.get_optimal_action,
 * policy/eval.py:126-128, once per environment step).  gradient / adjoints / iteration are nullable. */
int gmpc_ilqr_host(gmpc_handle* h, int64_t B, const float* x0_host, const float* U0_host,
                   const float* goal_host, const gmpc_ilqr_options* opt, float* X_host, float* U_host,
                   float* obj_host, float* gradient_host, float* adjoints_host,
                   int32_t* iteration_host, void* stream);

/* bilevel_optimization (policy/optimizers.py:34-75) for loss = L2MPC.loss (norm/l2_policy.py:12-18),
 * the per-sample body of BaseMPC.loss_and_grad (policy/base.py:87-128), in the same kernel launch
 * as the iLQR solve it starts with (maxiter = 0 evaluates the tail at U0 itself):
 *   X, U, obj, low_level_grad, iteration  = ilqr(...)                         (:55)
 *   loss        = L2MPC.loss(X, desired)                                       (:73)
 *   loss_grad_U = loss_grad_wrt_control          [B,T,m]   (nullable)          (:61-63, :78-83)
 *   hessian     = cost_hessian_wrt_control       [B,Tm,Tm] (nullable)          (:64-66, :86-90)
 *   H           = solve(hessian, loss_grad_U)    [B,T,m]   LU, partial pivoting (:67)
 * and the pieces of cost_vjp (:69-71, :93-105) that do not involve the cost-MLP weights:
 *   dxT              [B,n]  tangent of the terminal state along H (d x_T / dU . H)
 *   grad_mpc_weights [B,3]  d (H . grad_U J) / d mpc_weights (raw, pre-sigmoid values)
 * The cost-MLP part of cost_vjp is  w2 * grad_theta d/de ||f(x_T + e dxT; theta)||^2  (one forward-
 * over-reverse pass of the cost MLP per sample), assembled by the host mirror from x_T and dxT.
 * The dynamics MLP is piecewise linear, so jax.hessian of the objective equals
 * sum_t S_t^T Q_t S_t + blockdiag(R_t) with S_t = d x_t / dU; that is what is computed.
 * V (nullable, [B,T,m]): cost_vjp's direction given by the caller (policy/optimizers.py:93) -- used
 * in place of H, skipping the Hessian and the solve (hessian must then be null; H returns V). */
int gmpc_bilevel_l2(gmpc_handle* h, int64_t B, const float* x0, const float* U0, const float* goal,
                    const float* desired, const gmpc_ilqr_options* opt, float* X, float* U,
                    float* obj, float* low_level_grad, int32_t* iteration, float* loss,
                    float* loss_grad_U, float* hessian, float* H, float* dxT,
                    float* grad_mpc_weights, const float* V, void* stream);

/* The same tail for an arbitrary loss of the planned states, given dL/dX [B,T+1,n] at the plan U
 * (policy/optimizers.py:59-71 with loss = JS_MPC.generator_loss, gan/js_policy.py:60-74: dL/dX comes
 * from gmpc_critic_input_grad): loss_grad_U (nullable) = back-propagation of dL/dX through the
 * rollout, hessian (nullable), H, dxT, grad_mpc_weights as in gmpc_bilevel_l2. */
int gmpc_bilevel_tail(gmpc_handle* h, int64_t B, const float* x0, const float* U, const float* goal,
                      const float* dLdX, float* loss_grad_U, float* hessian, float* H, float* dxT,
                      float* grad_mpc_weights, void* stream);

/* d score / d xseq of CriticModel.predict (critic/critic_model.py:15-16): BPTT through the LSTM to
 * its inputs.  xseq[Bc,T1,n] -> logit[Bc], dxseq[Bc,T1,n].  The generator loss of
 * gan/js_policy.py:60-68, mean(-log p + log(1 - p)) with p = sigmoid(score), equals -score, so its
 * state gradient is -dxseq. */
int gmpc_critic_input_grad(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                           const float* params_flat, float* logit, float* dxseq, void* stream);

/* Work counters of the gmpc_ilqr calls since the last query (synchronises `stream`): tile-level
 * outer iterations and rollouts (1 + line-search trials) summed over the 32-trajectory tiles. */
int gmpc_ilqr_stats(gmpc_handle* h, int64_t* outer_iterations, int64_t* rollouts, void* stream);

/* The fp16-split tensor-core kernels represent operands as fp16 hi + lo parts.  GMPC_PATH_T128 and
 * GMPC_PATH_TC16S (what AUTO picks) rescale every forward and adjoint operand per trajectory by an exact
 * power of two, so states of any magnitude are in range; the opt-in GMPC_PATH_TC16 does not rescale the
 * forward operands.  A scaled operand above 65000 (only hidden activations ~4000 x larger than their layer's
 * input can get there) is clamped and counted.  This call synchronises `stream`, returns the number of CTAs
 * that clamped since the last call in *count and resets the counter.  A non-zero count means the results of
 * those calls are outside the 1e-4 parity contract: re-run them on GMPC_PATH_FFMA.  gmpc_plan_host does this by
 * itself when the path is AUTO; the device-pointer gmpc_plan never synchronises, so its callers check this
 * counter themselves (the Python mirror does: _lib.Handle.plan(check_range=True), EvalMPC._plan). */
int gmpc_range_overflow(gmpc_handle* h, int32_t* count, void* stream);

/* Kernel launches issued by this handle since creation (for bench.py's gpu_launches). */
int64_t gmpc_launch_count(const gmpc_handle* h);

/* The dynamics trainer's loss (norm/dynamics_trainer.py:13-44 predict_loss, vmapped at :66-77) and the
 * factors of its weight gradient.  For B windows of S recorded steps: roll the dynamics MLP from the
 * recorded state at every step (teacher_forcing != 0) or from its own prediction, and
 *   loss[b] = sum_t discount^t |x'_t - next_xseq[b,t]|^2                 (utils.discounted_sum).
 * xseq[B,S,n], useq[B,S,m], next_xseq[B,S,n] -> loss[B], and for every Dense layer l of the dynamics
 * MLP two GEMM-ready matrices with R = gmpc_dynamics_fit_columns(h, B, S) columns (one column per
 * (window, step); padded windows are zero in cot):
 *   act[l] [K_l, R]  the layer's inputs,      cot[l] [N_l, R]  d (sum_b loss[b]) / d (layer output),
 * including the back-propagation through time of the state adjoint when free running.
 * The gradient of the batch-mean loss is then  dW_l = act[l] cot[l]^T / B,  db_l = rowsum(cot[l]) / B:
 * one plain GEMM per layer, left to the caller's BLAS (cuBLAS; the host mirror uses torch.matmul). */
int gmpc_dynamics_fit(gmpc_handle* h, int64_t B, int32_t S, const float* xseq, const float* useq,
                      const float* next_xseq, float discount_factor, int32_t teacher_forcing,
                      float* loss, float* const* act, float* const* cot, void* stream);
int64_t gmpc_dynamics_fit_columns(const gmpc_handle* h, int64_t B, int32_t S);

/* The contraction gmpc_dynamics_fit leaves:  C[M,N] = alpha A B^T (+ C when accumulate != 0),  A[M,R], B[N,R] row-major,
 * and optionally  rowsum_B[N] = alpha sum_r B[n,r]  (nullable).  With A = act[l], B = cot[l], alpha = 1/B this is
 * value_and_grad of the batch-mean predict_loss w.r.t. Dense_l's kernel [in,out] and bias
 * (norm/dynamics_trainer.py:64-79).  Deterministic (no split over the reduction). */
int gmpc_gemm_nt(gmpc_handle* h, int32_t M, int32_t N, int64_t R, const float* A, const float* B, float alpha,
                 int32_t accumulate, float* C, float* rowsum_B, void* stream);

/* The cost-MLP part of cost_vjp (policy/optimizers.py:93-105, the last factor of bilevel_optimization :69-71),
 * reduced over the batch:  gW[l] [in_l,out_l], gb[l] [out_l]  =  scale * sum_b grad_theta of
 *   sigmoid(mpc_weights[2]) * d/de |f(x_T[b] + e dx_T[b]; theta)|^2 ,   f = the staged cost MLP (cost/nn.py:23-29),
 * x_T[B,n] the terminal states of the plans, dx_T[B,n] their tangents along H (gmpc_bilevel_l2 / _tail outputs).
 * scale = 1/B gives policy/base.py:126-127's batch mean, scale = 1 the sum a data-parallel caller all-reduces. */
int gmpc_cost_mixed_vjp(gmpc_handle* h, int64_t B, const float* xT, const float* dxT, float scale,
                        float* const* gW, float* const* gb, void* stream);

/* The expert proposal network in front of the planner: EvalMPC.get_goal_states_init_actions
 * (policy/eval.py:87-107, policy/base.py:40-61) = ExpertModel.get_history_carry (LSTM carry warmed up
 * on the history rows with teacher forcing, expert/expert_model.py:60-71) followed by
 * get_carry_next_state_and_action_seq with teacher_forcing=False (:73-91) on the scanned cell of
 * expert/nn.py:22-60.  history_x[B,hist+1,n] (last row = current state) ->
 * goal_xseq[B,T+1,n] (row 0 = current state), init_useq[B,T,m]: exactly the planner's inputs, so
 * acting is expert -> plan with no host round trip.
 * lstm_features > 0: ScanLSTM (OptimizedLSTMCell(F), heads MLPCell(num_layers));
 * lstm_features = 0: ScanMLP (Dense(H)+relu trunk, heads MLPCell(num_layers - 1)).
 * params_flat layout (flax kernels [in,out], gates i,f,g,o):
 *   LSTM trunk  Wi[n,4F] | Wh[F,4F] | bh[4F]        or   MLP trunk  D0[n,H] | b0[H]
 *   then the next_x head {Dk, Db}* | Wo[.,n] | bo[n], then the action head {Dk, Db}* | Wo[.,m] | bo[m]. */
int gmpc_expert_propose(gmpc_handle* h, int64_t B, int32_t hist, const float* history_x,
                        const float* params_flat, int32_t lstm_features, int32_t num_layers,
                        int32_t num_hidden_units, float* goal_xseq, float* init_useq, void* stream);
/* Number of floats of that layout (-1 for an invalid shape). */
int64_t gmpc_expert_param_count(int32_t n, int32_t m, int32_t lstm_features, int32_t num_layers,
                                int32_t num_hidden_units);

/* CriticModel.predict (critic/critic_model.py:15-16, critic/nn.py:27-42):
 * xseq[Bc,T1,n], params_flat -> logit[Bc].  T1 = number of rows (T+1 in the reference). */
int gmpc_critic_forward(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                        const float* params_flat, float* logit, void* stream);

/* JS_MPC.critic_loss_and_grad (gan/js_policy.py:41-58): BCE with +-1 labels (label > 0 test),
 * batch mean, gradient w.r.t. the flat critic params.  `inv_count` scales loss and gradient
 * (1/Bc on one GPU; 1/global_batch when the gradient is all-reduced across ranks afterwards).
 * xseq[Bc,T1,n], label[Bc] -> loss[1], grad_flat[P]. */
int gmpc_critic_loss_grad(gmpc_handle* h, int64_t Bc, int32_t T1, const float* xseq,
                          const float* label, const float* params_flat, float inv_count,
                          float* loss, float* grad_flat, void* stream);

/* Same with a gather fused in front (gan/critic_trainer.py:55-56, X[p], Y[p]):
 * sample i of the minibatch is data_xseq[perm[i]], data_label[perm[i]] (perm int32, device). */
int gmpc_critic_loss_grad_gather(gmpc_handle* h, int64_t Bc, int32_t T1, const float* data_xseq,
                                 const float* data_label, const int32_t* perm,
                                 const float* params_flat, float inv_count, float* loss,
                                 float* grad_flat, void* stream);

/* optax.chain(clip_by_global_norm(max_norm), adam(lr)) + apply_updates on a flat vector
 * (norm/runner.py:46-58, gan/critic_trainer.py:58-59).  `step` is the 1-based update count.
 * grad_flat is scaled by grad_scale first (1.0, or 1/world after a sum all-reduce).
 * In place on params_flat[P], mom[P], vel[P]. */
int gmpc_clip_adam_step(gmpc_handle* h, int64_t P, float* params_flat, const float* grad_flat,
                        float* mom, float* vel, int32_t step, float lr, float max_norm,
                        float grad_scale, float b1, float b2, float eps, void* stream);

/* The whole minibatch scan of train_critic_parameters (gan/critic_trainer.py:48-65, a lax.scan in the
 * reference) enqueued by one call: for s < steps { gather perm[s, :], BCE loss + flat gradient
 * (batch mean), clip_by_global_norm(max_norm) + Adam step number step0 + s + 1 } on one GPU.
 * perm is int32 [steps, Bc] on the device (sampled with replacement by the caller, :88-90);
 * losses[steps] receives the minibatch losses; grad_scratch[P] is caller-owned scratch.
 * Multi-GPU data-parallel training keeps the per-step calls with the all-reduce in between. */
int gmpc_critic_train_scan(gmpc_handle* h, int32_t steps, int64_t Bc, int32_t T1,
                           const float* data_xseq, const float* data_label, const int32_t* perm,
                           float* params_flat, float* mom, float* vel, int32_t step0, float lr,
                           float max_norm, float b1, float b2, float eps, float* losses,
                           float* grad_scratch, void* stream);

/* L2MPC.loss (norm/l2_policy.py:12-18): X[B,T+1,n], desired[B,T+1,n] -> loss[B]. */
int gmpc_l2_loss(gmpc_handle* h, int64_t B, const float* X, const float* desired, float* loss,
                 void* stream);

/* Diagnostics (not on the hot path): measured FP32 FFMA throughput of `device` in TFLOP/s, the
 * second roofline denominator bench.py reports for the CUDA-core path. */
int gmpc_measure_fp32_peak(int device, float* tflops_out);

/* Diagnostics: dense kind::f16 tcgen05 throughput of `device` in TFLOP/s (every SM issuing M=128, N=256,
 * K=16 MMAs from shared memory): the peak of the tensor pipe the planner kernels run on.  With the
 * three-product fp16 split (ah Wh + al Wh + ah Wl) the algorithmic ceiling is a third of it. */
int gmpc_measure_f16_mma_peak(int device, float* tflops_out);

#ifdef __cplusplus
}
#endif
#endif /* GMPC_H_ */
