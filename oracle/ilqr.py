"""CPU restatement of the reference's ACTUAL planner step: trajax iLQR on the gan_mpc objective.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED, twice over: the reference
ships no tests or golden vectors, and the algorithm lives in a third-party dependency that is not
vendored: `trajax` pinned at git commit c94a637c5a397b3d4100153f25b4b165507b5b20
(requirements.txt:51).  Its source is not on disk and cannot be fetched, so `ilqr`, `tvlqr`,
`lqr_step`, `line_search_ddp`, `ddp_rollout` and `adjoint` are restated here from the published
algorithm as recalled; the call sites that anchor it are
policy/optimizers.py:10-21 (ilqr_solve), :55 (the 7-tuple that is unpacked) and policy/eval.py:10-20
(the options: maxiter 100, grad_norm_threshold 1e-4, alpha_0 1.0, alpha_min 5e-5, make_psd False).

What trajax computes with jax.jacobian / jax.hessian is written in closed form here (the models are
ReLU MLPs and pseudo-Huber norms); tests/test_oracle_ilqr.py checks every closed form against
torch.func autodiff of the plain cost/dynamics functions of oracle/planner.py.

Everything is batched over a leading axis with vmap semantics: a lane whose own loop condition is
false is frozen while other lanes continue (jax.vmap of lax.while_loop), so each lane's result is
what the unbatched call would return.
"""

import torch

from oracle import planner as pl

DELTA = 1e-8  # trajax tvlqr.lqr_step: smallest eigenvalue floor of G

ILQR_KWARGS = {  # policy/eval.py:10-20
    "maxiter": 100, "grad_norm_threshold": 1e-4, "relative_grad_norm_threshold": 0.0,
    "obj_step_threshold": 0.0, "inputs_step_threshold": 0.0, "make_psd": False, "psd_delta": 0.0,
    "alpha_0": 1.0, "alpha_min": 0.00005,
}


def _sym(M):
    return 0.5 * (M + M.transpose(-1, -2))


# --------------------------------------------------------------------------- derivatives
def mlp_jacobian(q, Ws, bs):
    """d MLP(q) / d q for the relu MLP of dynamics/nn.py:27-34 / cost/nn.py:23-29 (no residual).
    q [..., in] -> (out [..., out], Jac [..., out, in]); relu'(0) = 0 as in jax."""
    out, masks = pl._mlp_forward_keep(q, Ws, bs)
    nout = Ws[-1].shape[1]
    eye = torch.eye(nout, dtype=q.dtype)
    seeds = eye.expand(*q.shape[:-1], nout, nout)
    mk = [m.unsqueeze(-2) for m in masks]
    return out, pl._mlp_input_vjp(seeds, Ws, mk)


def huber_grad_hess(v, w):
    """gradient and Hessian of w * (sqrt(v.v + a^2) - a) (cost/cost_model.py:22-27) w.r.t. v."""
    s = torch.sqrt((v * v).sum(-1, keepdim=True) + pl.ALPHA ** 2)
    g = w * v / s
    eye = torch.eye(v.shape[-1], dtype=v.dtype)
    H = w * (eye / s.unsqueeze(-1) - v.unsqueeze(-1) * v.unsqueeze(-2) / (s ** 3).unsqueeze(-1))
    return g, H


def lqr_params(X, U, goal_X, params):
    """trajax ilqr.get_lqr_params for the gan_mpc cost/dynamics: (Q, q, R, r, M, A, B).
    Q [B,T+1,n,n], q [B,T+1,n], R [B,T+1,m,m], r [B,T+1,m], M [B,T+1,n,m] (identically 0: the cost has
    no state-action cross term), A [B,T,n,n], B [B,T,n,m] (the unused Jacobians at t = T are not
    computed).  Rows t = T of R, r are 0 (pad(U) row, staging branch not selected)."""
    B_, T, m = U.shape
    n = X.shape[-1]
    w = torch.sigmoid(params["mpc_weights"])
    qd = torch.cat([X[:, :T], U], -1)
    _, Jd = mlp_jacobian(qd, params["dyn_W"], params["dyn_b"])
    A = Jd[..., :n] + torch.eye(n, dtype=X.dtype)
    Bm = Jd[..., n:]
    gq, Hq = huber_grad_hess(X[:, :T] - goal_X[:, :T], w[1])
    gr, Hr = huber_grad_hess(U, w[0])
    y, Jc = mlp_jacobian(X[:, T], params["cost_W"], params["cost_b"])
    qT = 2.0 * w[2] * (Jc.transpose(-1, -2) @ y.unsqueeze(-1)).squeeze(-1)
    QT = 2.0 * w[2] * (Jc.transpose(-1, -2) @ Jc)
    Q = torch.cat([Hq, QT[:, None]], 1)
    q = torch.cat([gq, qT[:, None]], 1)
    R = torch.cat([Hr, torch.zeros(B_, 1, m, m, dtype=X.dtype)], 1)
    r = torch.cat([gr, torch.zeros(B_, 1, m, dtype=X.dtype)], 1)
    M = torch.zeros(B_, T + 1, n, m, dtype=X.dtype)
    return Q, q, R, r, M, A, Bm


def adjoint(A, Bm, q, r):
    """trajax `adjoint`: g_t = r_t + B_t^T p_{t+1}; p_t = A_t^T p_{t+1} + q_t from p_T = q_T.
    Returns gradient [B,T,m], adjoints [B,T+1,n]."""
    T = A.shape[1]
    p = q[:, T]
    ps, gs = [p], []
    for t in range(T - 1, -1, -1):
        gs.append(r[:, t] + (Bm[:, t].transpose(-1, -2) @ p.unsqueeze(-1)).squeeze(-1))
        p = (A[:, t].transpose(-1, -2) @ p.unsqueeze(-1)).squeeze(-1) + q[:, t]
        ps.append(p)
    return torch.stack(gs[::-1], 1), torch.stack(ps[::-1], 1)


def lqr_step(P, p, Q, q, R, r, M, A, Bm):
    """trajax tvlqr.lqr_step with c = 0 (the trajectory is dynamically feasible)."""
    At, Bt = A.transpose(-1, -2), Bm.transpose(-1, -2)
    AtP = At @ P
    AtPA = _sym(AtP @ A)
    BtP = Bt @ P
    BtPA = BtP @ A
    G = _sym(R + BtP @ Bm)
    S = torch.linalg.eigvalsh(G)
    shift = torch.clamp(DELTA - S[..., 0], min=0.0)
    G_ = G + shift[..., None, None] * torch.eye(G.shape[-1], dtype=G.dtype)
    H = BtPA + M.transpose(-1, -2)
    h = (Bt @ p.unsqueeze(-1)).squeeze(-1) + r
    K = -torch.linalg.solve(G_, H)
    k = -torch.linalg.solve(G_, h.unsqueeze(-1)).squeeze(-1)
    H_GK = H + G @ K
    Pn = _sym(Q + AtPA + H_GK.transpose(-1, -2) @ K + K.transpose(-1, -2) @ H)
    pn = (q + (At @ p.unsqueeze(-1)).squeeze(-1)
          + (H_GK.transpose(-1, -2) @ k.unsqueeze(-1)).squeeze(-1)
          + (K.transpose(-1, -2) @ h.unsqueeze(-1)).squeeze(-1))
    return Pn, pn, K, k


def tvlqr(Q, q, R, r, M, A, Bm):
    """trajax tvlqr backward pass: gains K [B,T,m,n], k [B,T,m]."""
    T = A.shape[1]
    P, p = _sym(Q[:, T]), q[:, T]
    Ks, ks = [], []
    for t in range(T - 1, -1, -1):
        P, p, K, k = lqr_step(P, p, Q[:, t], q[:, t], R[:, t], r[:, t], M[:, t], A[:, t], Bm[:, t])
        Ks.append(K)
        ks.append(k)
    return torch.stack(Ks[::-1], 1), torch.stack(ks[::-1], 1)


def ddp_rollout(X, U, K, k, alpha, goal_X, params):
    """trajax ddp_rollout + total cost: u = U[t] + alpha k[t] + K[t] (x_new[t] - X[t])."""
    T = U.shape[1]
    x = X[:, 0]
    xs, us = [x], []
    for t in range(T):
        u = U[:, t] + alpha * k[:, t] + (K[:, t] @ (x - X[:, t]).unsqueeze(-1)).squeeze(-1)
        x = pl.dynamics_mlp(x, u, params["dyn_W"], params["dyn_b"])
        xs.append(x)
        us.append(u)
    Xn, Un = torch.stack(xs, 1), torch.stack(us, 1)
    J = torch.zeros(X.shape[0], dtype=X.dtype)
    zero_u = torch.zeros_like(Un[:, 0])
    for t in range(T + 1):
        J = J + pl.step_cost(Xn[:, t], Un[:, t] if t < T else zero_u, t, T, params, goal_X)
    return Xn, Un, J


def line_search_ddp(X, U, K, k, obj, goal_X, params, alpha_0, alpha_min, lanes):
    """trajax line_search_ddp under vmap: halve alpha until the objective strictly decreases.
    `lanes`: lanes of the outer loop that are still iterating (the others are frozen).
    Returns X, U, obj, alpha (alpha = half of the last alpha tried)."""
    obj = torch.where(torch.isnan(obj), torch.full_like(obj, float("inf")), obj)
    Xr, Ur, objr = X.clone(), U.clone(), obj.clone()
    alpha_r = torch.full_like(obj, alpha_0)
    searching = lanes.clone()  # while cond at entry: obj >= obj and alpha_0 > alpha_min
    searching &= alpha_r > alpha_min
    alpha = alpha_0
    while bool(searching.any()):
        Xn, Un, objn = ddp_rollout(X, U, K, k, alpha, goal_X, params)
        objn = torch.where(torch.isnan(objn), obj, objn)
        better = searching & (objn < obj)
        Xr[better], Ur[better] = Xn[better], Un[better]
        objr = torch.where(better, torch.minimum(objn, obj), objr)
        alpha = 0.5 * alpha
        alpha_r = torch.where(searching, torch.full_like(obj, alpha), alpha_r)
        searching = searching & ~better & (alpha_r > alpha_min)
    return Xr, Ur, objr, alpha_r


def ilqr(x0, U0, goal_X, params, maxiter=100, grad_norm_threshold=1e-4, alpha_0=1.0,
         alpha_min=0.00005, gradient_lag=False, **unused):
    """trajax ilqr as called by policy/optimizers.py:19-21, batched (x0 [B,n], U0 [B,T,m],
    goal_X [B,T+1,n]).  Returns (X, U, obj, gradient, adjoints, lqr, iteration) like the 7-tuple
    unpacked at policy/optimizers.py:55; lqr = (Q, q, R, r, M, A, B) at the returned trajectory.
    The thresholds relative_grad_norm / obj_step / inputs_step are 0.0 in the reference
    (policy/eval.py:13-15): with strict comparisons they only stop a lane that made no progress,
    which the alpha > alpha_min test already does.

    gradient_lag: trajax's loop body as recalled (source not on disk) unpacks the lqr
    tuple BEFORE the step and calls `adjoint(A, B, q, r)` on it after the line search, before
    `lqr = get_lqr_params(X, U)`: the returned gradient / adjoints and the grad_norm test then belong to the
    iterate before the last step.  False (default) keeps them at the returned trajectory."""
    X, obj = pl.objective(x0, U0, goal_X, params)
    U = U0.clone()
    lqr = lqr_params(X, U, goal_X, params)
    gradient, adjoints = adjoint(lqr[5], lqr[6], lqr[1], lqr[3])   # at the current iterate
    g_ret, a_ret = gradient, adjoints                                # what the loop carries when gradient_lag
    B = x0.shape[0]
    alpha = torch.full((B,), alpha_0, dtype=x0.dtype)
    iteration = torch.zeros(B, dtype=torch.int32)

    def cont():
        gn = (g_ret if gradient_lag else gradient).flatten(1).norm(dim=1)
        gn = torch.where(torch.isnan(gn), torch.full_like(gn, float("inf")), gn)   # trajax: a NaN gradient norm keeps iterating
        return (iteration < maxiter) & (gn > grad_norm_threshold) & (alpha > alpha_min)

    lanes = cont()
    while bool(lanes.any()):
        K, k = tvlqr(*lqr)
        if gradient_lag:   # adjoint of the lqr tuple unpacked before the step = the gradient at the old iterate
            g_ret = torch.where(lanes[:, None, None], gradient, g_ret)
            a_ret = torch.where(lanes[:, None, None], adjoints, a_ret)
        X, U, obj_n, alpha_n = line_search_ddp(X, U, K, k, obj, goal_X, params, alpha_0, alpha_min, lanes)
        obj = torch.where(lanes, obj_n, obj)
        alpha = torch.where(lanes, alpha_n, alpha)
        lqr = lqr_params(X, U, goal_X, params)
        gradient, adjoints = adjoint(lqr[5], lqr[6], lqr[1], lqr[3])
        iteration = iteration + lanes.to(torch.int32)
        lanes = lanes & cont()
    if gradient_lag:
        return X, U, obj, g_ret, a_ret, lqr, iteration
    return X, U, obj, gradient, adjoints, lqr, iteration
