"""CPU restatement of the reference's bilevel (cost-training) gradient, by literal autodiff.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (no reference tests exist and
the JAX stack is not installable).

Follows policy/optimizers.py:34-105 line by line with torch autograd standing in for jax.grad /
jax.hessian (ReLU has zero second derivative in both), and policy/base.py:87-128 for the batch
mean.  The kernels compute the same quantities in structured (Gauss-Newton) form from the explicit
Jacobians, so agreement with this file checks that algebra, not just the arithmetic.

Reference quirks reproduced as-is: the sign of the high-level gradient is
+H^T d_theta grad_U J (the implicit-function theorem has a minus); only the COST side of `params`
(cost_params, mpc_weights) receives a gradient, the dynamics is closed over the outer params
(policy/optimizers.py:50-51,96-101); the Hessian is solved without regularisation (:67).
"""

import torch

from oracle import ilqr as oilqr
from oracle import planner as pl


def _objective_of(U, x0, goal, params):
    """objective (policy/optimizers.py:24-31) as a differentiable function of U and params."""
    return pl.objective(x0[None], U[None], goal[None], params)[1][0]


def _loss_fn(desired, loss_fn):
    """loss(X) -> scalar: L2MPC.loss against `desired` (norm/l2_policy.py:12-18) unless a callable
    (e.g. the generator loss of gan/js_policy.py:60-68 on the oracle critic) is given."""
    return loss_fn if loss_fn is not None else (lambda X: pl.l2_loss(X, desired))


def loss_grad_wrt_control(x0, U, desired, params, loss_fn=None):
    """policy/optimizers.py:78-83."""
    U = U.detach().clone().requires_grad_(True)
    X = pl.rollout(x0[None], U[None], params)[0]
    (g,) = torch.autograd.grad(_loss_fn(desired, loss_fn)(X), U)
    return g


def cost_hessian_wrt_control(x0, U, goal, params):
    """policy/optimizers.py:86-90 -- jax.hessian of the objective w.r.t. U, [T,m,T,m]."""
    return torch.autograd.functional.hessian(lambda u: _objective_of(u, x0, goal, params), U.detach())


def cost_vjp(V, x0, U, goal, params):
    """policy/optimizers.py:93-105 -- grad_theta ( V . grad_U J(U; theta) ) over the cost side of
    params.  Returns dict(cost_W=[...], cost_b=[...], mpc_weights=...)."""
    leaves = [w.detach().clone().requires_grad_(True) for w in params["cost_W"]]
    leaves += [b.detach().clone().requires_grad_(True) for b in params["cost_b"]]
    leaves += [params["mpc_weights"].detach().clone().requires_grad_(True)]
    L = len(params["cost_W"])
    p2 = dict(params, cost_W=leaves[:L], cost_b=leaves[L:2 * L], mpc_weights=leaves[2 * L])
    Uv = U.detach().clone().requires_grad_(True)
    (gU,) = torch.autograd.grad(_objective_of(Uv, x0, goal, p2), Uv, create_graph=True)
    outer = (V.reshape(-1) * gU.reshape(-1)).sum()
    gs = torch.autograd.grad(outer, leaves, allow_unused=True)
    gs = [torch.zeros_like(l) if g is None else g for g, l in zip(gs, leaves)]
    return dict(cost_W=gs[:L], cost_b=gs[L:2 * L], mpc_weights=gs[2 * L])


def bilevel_tail(x0, U, goal, desired, params, loss_fn=None):
    """policy/optimizers.py:59-73 at a given planned U (unbatched):
    returns (high_level_loss, B [T,m], A [Tm,Tm], H [T,m], high_level_grad dict)."""
    T, m = U.shape
    X = pl.rollout(x0[None], U[None], params)[0]
    Bv = loss_grad_wrt_control(x0, U, desired, params, loss_fn).reshape(T * m)
    A = cost_hessian_wrt_control(x0, U, goal, params).reshape(T * m, T * m)
    H = torch.linalg.solve(A, Bv)
    grad = cost_vjp(H, x0, U, goal, params)
    return _loss_fn(desired, loss_fn)(X), Bv.reshape(T, m), A, H.reshape(T, m), grad


def bilevel_optimization(x0, init_U, goal, desired, params, **ilqr_kwargs):
    """policy/optimizers.py:34-75 (unbatched): (high_level_loss, low_level_grad, high_level_grad, itr)."""
    X, U, _, low, _, _, itr = oilqr.ilqr(x0[None], init_U[None], goal[None], params, **ilqr_kwargs)
    loss, _, _, _, grad = bilevel_tail(x0, U[0], goal, desired, params)
    return loss, low[0], grad, itr[0]


def loss_and_grad(x0, init_U, goal, desired, params, **ilqr_kwargs):
    """BaseMPC.loss_and_grad (policy/base.py:87-128): vmap of bilevel_optimization, mean loss and
    leaf-wise mean of the per-sample gradients.  Batched inputs."""
    B = x0.shape[0]
    losses, acc = [], None
    for b in range(B):
        loss, _, g, _ = bilevel_optimization(x0[b], init_U[b], goal[b], desired[b], params, **ilqr_kwargs)
        losses.append(loss)
        flat = g["cost_W"] + g["cost_b"] + [g["mpc_weights"]]
        acc = flat if acc is None else [a + f for a, f in zip(acc, flat)]
    L = len(params["cost_W"])
    acc = [a / B for a in acc]
    return torch.stack(losses).mean(), dict(cost_W=acc[:L], cost_b=acc[L:2 * L], mpc_weights=acc[2 * L])
