"""CPU restatement of the gan_mpc planner arithmetic (torch, any float dtype).

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (no
reference tests/golden vectors exist and the JAX stack is not installable).

Every function cites the reference file:line it restates (paths relative to
/root/reference).  Everything is batched over a leading axis so the fp32
variant doubles as the timed, vectorised CPU baseline (whole-batch matmul per
layer per step, all host threads).

Weights use the flax layout: kernel[in, out], y = x @ kernel + bias.
`params` here is a plain dict:
    dyn_W / dyn_b   : lists of L_d tensors   (dynamics/nn.py:27-34)
    cost_W / cost_b : lists of L_c tensors   (cost/nn.py:23-29)
    mpc_weights     : tensor[3]  (action, state, terminal)  (policy/eval.py:58)
"""

import torch

ALPHA = 1e-2  # cost/cost_model.py:22


# --------------------------------------------------------------------------- models
def dynamics_mlp(x, u, dyn_W, dyn_b):
    """dynamics/nn.py:27-34 -- x' = Dense(n)(relu-MLP([x,u])) + x (carry empty for MLP)."""
    q = torch.cat([x, u], dim=-1)
    for W, b in zip(dyn_W[:-1], dyn_b[:-1]):
        q = torch.relu(q @ W + b)
    return q @ dyn_W[-1] + dyn_b[-1] + x


def cost_mlp(x, cost_W, cost_b):
    """cost/nn.py:23-29 -- ||Dense(fout)(relu-MLP(x))||^2."""
    z = x
    for W, b in zip(cost_W[:-1], cost_b[:-1]):
        z = torch.relu(z @ W + b)
    y = z @ cost_W[-1] + cost_b[-1]
    return (y * y).sum(-1)


def pseudo_huber(v):
    """cost/cost_model.py:22-27 -- sqrt(v.v + alpha^2) - alpha over the last axis."""
    return torch.sqrt((v * v).sum(-1) + ALPHA ** 2) - ALPHA


def staging_cost(x, u, w2, goal):
    """cost/cost_model.py:20-28 -- weights @ [u_cost, x_cost]."""
    return w2[0] * pseudo_huber(u) + w2[1] * pseudo_huber(x - goal)


def step_cost(x, u, t, T, params, goal_X):
    """cost/cost_model.py:33-42 -- where(t == horizon, terminal, staging), sigmoid weights.

    goal_X: [..., T+1, n]; x: [..., n]; u: [..., m]."""
    w = torch.sigmoid(params["mpc_weights"])
    if t == T:
        return w[2] * cost_mlp(x, params["cost_W"], params["cost_b"])
    return staging_cost(x, u, w[:2], goal_X[..., t, :])


# --------------------------------------------------------------------------- objective
def rollout(x0, U, params):
    """trajax rollout as used at policy/optimizers.py:28,80 -- X[0]=x0, X[t+1]=dyn(X[t],U[t]).

    x0 [..., n], U [..., T, m] -> X [..., T+1, n]."""
    T = U.shape[-2]
    xs = [x0]
    x = x0
    for t in range(T):
        x = dynamics_mlp(x, U[..., t, :], params["dyn_W"], params["dyn_b"])
        xs.append(x)
    return torch.stack(xs, dim=-2)


def objective(x0, U, goal_X, params):
    """policy/optimizers.py:24-31 -- J = sum_t cost(X[t], pad(U)[t], t), t = 0..T."""
    T = U.shape[-2]
    X = rollout(x0, U, params)
    J = torch.zeros(X.shape[:-2], dtype=X.dtype)
    zero_u = torch.zeros_like(U[..., 0, :])
    for t in range(T + 1):
        u = U[..., t, :] if t < T else zero_u  # trajax pad(U): one zero row appended
        J = J + step_cost(X[..., t, :], u, t, T, params, goal_X)
    return X, J


def _mlp_forward_keep(q, Ws, bs, margin=None):
    """forward through relu layers keeping the (pre-activation > 0) masks.
    `margin` (optional 1-element list holding a tensor): running per-row min |pre-activation|,
    i.e. the distance of the row to the nearest ReLU kink (test diagnostics)."""
    masks = []
    for W, b in zip(Ws[:-1], bs[:-1]):
        z = q @ W + b
        masks.append(z > 0)  # jax relu grad is 1 only for x > 0
        if margin is not None:
            mz = z.abs().amin(-1)
            margin[0] = mz if margin[0] is None else torch.minimum(margin[0], mz)
        q = torch.relu(z)
    return q @ Ws[-1] + bs[-1], masks


def _mlp_input_vjp(dy, Ws, masks):
    """input-adjoint of the relu MLP (no weight grads): back through W_{L-1}, masks, ... W_0."""
    d = dy @ Ws[-1].T
    for W, mk in zip(reversed(Ws[:-1]), reversed(masks)):
        d = (d * mk.to(d.dtype)) @ W.T
    return d


def objective_grad(x0, U, goal_X, params, margin=None):
    """Hand-written adjoint of `objective` w.r.t. U -- what jax.grad computes at
    policy/optimizers.py:83,103 and what trajax ilqr returns as (gradient, adjoints).

    Returns X [..,T+1,n], J [..], dU [..,T,m], lam [..,T+1,n].
    margin: optional [None] list, filled with the per-trajectory min |hidden pre-activation|
    (the gradient is discontinuous where it is 0: a row with a tiny margin has no well-defined
    1e-4-accurate gradient in ANY fp32 implementation)."""
    T, m = U.shape[-2], U.shape[-1]
    n = x0.shape[-1]
    w = torch.sigmoid(params["mpc_weights"])
    dyn_W, dyn_b = params["dyn_W"], params["dyn_b"]
    a2 = ALPHA ** 2

    # forward, keeping relu masks
    xs, masks_t = [x0], []
    x = x0
    J = torch.zeros(x0.shape[:-1], dtype=x0.dtype)
    for t in range(T):
        u = U[..., t, :]
        J = J + staging_cost(x, u, w[:2], goal_X[..., t, :])
        out, masks = _mlp_forward_keep(torch.cat([x, u], -1), dyn_W, dyn_b, margin)
        masks_t.append(masks)
        x = out + x
        xs.append(x)
    y, cmasks = _mlp_forward_keep(x, params["cost_W"], params["cost_b"], margin)
    J = J + w[2] * (y * y).sum(-1)

    # adjoint sweep
    lam = _mlp_input_vjp(2.0 * w[2] * y, params["cost_W"], cmasks)
    lams = [lam]
    dUs = []
    for t in range(T - 1, -1, -1):
        u = U[..., t, :]
        dq = _mlp_input_vjp(lam, dyn_W, masks_t[t])
        g = w[0] * u / torch.sqrt((u * u).sum(-1, keepdim=True) + a2) + dq[..., n:]
        d = xs[t] - goal_X[..., t, :]
        lam = w[1] * d / torch.sqrt((d * d).sum(-1, keepdim=True) + a2) + lam + dq[..., :n]
        dUs.append(g)
        lams.append(lam)
    dU = torch.stack(dUs[::-1], dim=-2)
    lam_all = torch.stack(lams[::-1], dim=-2)
    return torch.stack(xs, dim=-2), J, dU, lam_all


# --------------------------------------------------------------------------- planner
def plan(x0, U0, goal_X, params, method="adam", iters=20, lr=1e-2,
         b1=0.9, b2=0.999, eps=1e-8):
    """First-order planner of BASELINE.json:north_star on the reference's exact objective.

    x0 [B,n], U0 [B,K,T,m], goal_X [B,T+1,n].
    grad : U <- U - lr g.
    adam : optax 0.1.7 scale_by_adam + scale(-lr) (norm/runner.py:53 defaults):
           m<-b1 m+(1-b1)g ; v<-b2 v+(1-b2)g^2 ; U <- U - lr (m/(1-b1^k)) / (sqrt(v/(1-b2^k))+eps).
    Then J_k = objective(U_k); idx = argmin_k (first minimum on ties, jnp.argmin).
    Returns U_best [B,T,m], X_best [B,T+1,n], J_best [B], idx [B] int32, J_all [B,K]."""
    B, K, T, m = U0.shape
    x0k = x0[:, None, :].expand(B, K, x0.shape[-1])
    gk = goal_X[:, None].expand(B, K, T + 1, goal_X.shape[-1])
    U = U0.clone()
    mom = torch.zeros_like(U)
    vel = torch.zeros_like(U)
    for k in range(1, iters + 1):
        _, _, g, _ = objective_grad(x0k, U, gk, params)
        if method == "grad":
            U = U - lr * g
        elif method == "adam":
            mom = b1 * mom + (1.0 - b1) * g
            vel = b2 * vel + (1.0 - b2) * g * g
            mhat = mom / (1.0 - b1 ** k)
            vhat = vel / (1.0 - b2 ** k)
            U = U - lr * mhat / (torch.sqrt(vhat) + eps)
        else:
            raise ValueError(method)
    X, J = objective(x0k, U, gk, params)
    idx = torch.argmin(J, dim=1)
    ar = torch.arange(B)
    return U[ar, idx], X[ar, idx], J[ar, idx], idx.to(torch.int32), J


# --------------------------------------------------------------------------- outer losses
def l2_loss(xcseq, desired_xseq):
    """norm/l2_policy.py:12-18 -- sum_j mean_t (X[t,j]-desired[t,j])^2 (carry columns sliced off)."""
    n = desired_xseq.shape[-1]
    diff = (xcseq[..., :n] - desired_xseq) ** 2
    return diff.mean(-2).sum(-1)


def loss_grad_wrt_control_l2(x0, U, desired_xseq, params):
    """policy/optimizers.py:78-83 with loss = L2MPC.loss: d loss(rollout(U)) / dU by BPTT."""
    T = U.shape[-2]
    n = x0.shape[-1]
    dyn_W, dyn_b = params["dyn_W"], params["dyn_b"]
    xs, masks_t = [x0], []
    x = x0
    for t in range(T):
        out, masks = _mlp_forward_keep(torch.cat([x, U[..., t, :]], -1), dyn_W, dyn_b)
        masks_t.append(masks)
        x = out + x
        xs.append(x)
    scale = 2.0 / (T + 1)
    lam = scale * (xs[T] - desired_xseq[..., T, :])
    dUs = []
    for t in range(T - 1, -1, -1):
        dq = _mlp_input_vjp(lam, dyn_W, masks_t[t])
        dUs.append(dq[..., n:])
        lam = scale * (xs[t] - desired_xseq[..., t, :]) + lam + dq[..., :n]
    return torch.stack(dUs[::-1], dim=-2)
