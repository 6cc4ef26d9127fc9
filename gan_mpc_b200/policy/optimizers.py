"""Planner entry points with the reference's names (policy/optimizers.py).

The reference passes arbitrary Python closures cost(x,u,t,...) / dynamics(x,u,t,...) to trajax.
A CUDA kernel cannot call closures, so this mirror recognises the STRUCTURED case -- the bound
methods `policy.cost` / `policy.dynamics` of an EvalMPC built from DynamicsModel(MLP) +
MujocoBasedModel(cost MLP) -- extracts the weights and runs libgmpc.  Anything else raises
(there is no CPU fallback).  Every function accepts the reference's unbatched shapes and the
same shapes with leading batch axes (the reference vmaps; the kernels batch natively).

Two planners sit behind ilqr_solve: the north-star first-order planner on the reference's exact
objective (planner method "adam" / "grad", the default) and the reference's own step, trajax iLQR
(planner method "ilqr": gmpc_ilqr), which honours trajax_ilqr_kwargs."""

import torch


def _owner(fn, what):
    pol = getattr(fn, "__self__", None)
    if pol is None or not hasattr(pol, "_plan"):
        raise TypeError(
            f"{what} must be the bound method of a gan_mpc_b200 policy (policy.cost / "
            "policy.dynamics): arbitrary closures cannot run inside the CUDA planner and there "
            "is no CPU fallback")
    return pol


class Bound:
    """`wrapped_cost` / `wrapped_dynamics` of the reference (policy/optimizers.py:13-17,47-51):
    a structured closure = (policy method, params, extra args)."""

    def __init__(self, fn, params, args=()):
        self.policy = _owner(fn, "cost/dynamics")
        self.params = params
        self.args = tuple(args)


def bind(fn, params, args=()):
    return Bound(fn, params, args)


def ilqr_solve(cost, dynamics, x0, U, params, cost_args, dynamics_args, trajax_ilqr_kwargs=None):
    """policy/optimizers.py:10-21.  Returns the trajax 7-tuple
    (X, U, obj, gradient, adjoints, lqr, iteration).  With the policy's planner method "ilqr" this
    is trajax iLQR under `trajax_ilqr_kwargs` (lqr = (A, B) Jacobians when planner_kwargs has
    return_lqr, else None); with the first-order planner `lqr` is None and `iteration` is the number
    of planning iterations run."""
    pol = _owner(cost, "cost")
    if _owner(dynamics, "dynamics") is not pol:
        raise ValueError("cost and dynamics must belong to the same policy")
    if len(dynamics_args) != 0:
        raise ValueError("MLP dynamics take no extra arguments (policy/eval.py:121)")
    if trajax_ilqr_kwargs is not None and pol.planner_kwargs["method"] == "ilqr":
        saved, pol.trajax_ilqr_kwargs = pol.trajax_ilqr_kwargs, trajax_ilqr_kwargs
        try:
            return pol._plan(x0, U, params, cost_args[0])
        finally:
            pol.trajax_ilqr_kwargs = saved
    return pol._plan(x0, U, params, cost_args[0])


def objective(cost, dynamics, U, x0):
    """policy/optimizers.py:24-31 -- cost/dynamics are `Bound` closures."""
    if not isinstance(cost, Bound) or not isinstance(dynamics, Bound):
        raise TypeError("objective() needs optimizers.bind(policy.cost, params, (goal_X,)) closures")
    J, _, _, _ = cost.policy._objective(x0, U, cost.params, cost.args[0], grad=False)
    return J


def objective_and_grad(cost, dynamics, U, x0):
    """value and jax.grad of `objective` w.r.t. U (what policy/optimizers.py:103 differentiates);
    also returns the rollout and the adjoints."""
    if not isinstance(cost, Bound) or not isinstance(dynamics, Bound):
        raise TypeError("objective_and_grad() needs optimizers.bind(...) closures")
    return cost.policy._objective(x0, U, cost.params, cost.args[0], grad=True)


def rollout(dynamics, U, x0):
    """trajax_opt.rollout as called at policy/optimizers.py:28,80."""
    if not isinstance(dynamics, Bound):
        raise TypeError("rollout() needs an optimizers.bind(policy.dynamics, params) closure")
    return dynamics.policy._rollout(x0, U, dynamics.params)


def loss_grad_wrt_control(loss, dynamics, x0, U, loss_args):
    """policy/optimizers.py:78-83 for loss = L2MPC.loss: d loss(rollout(U), U, params, desired)/dU.
    loss_args = (params, desired_xseq) as assembled at :59."""
    pol = _owner(loss, "loss")
    if getattr(loss, "__func__", None) is not getattr(type(pol), "loss", None) or not pol._loss_is_l2:
        raise NotImplementedError("only L2MPC.loss has a fused BPTT kernel in this round")
    if not isinstance(dynamics, Bound):
        raise TypeError("dynamics must be an optimizers.bind(policy.dynamics, params) closure")
    params, desired = loss_args
    return pol._l2_loss_grad(x0, U, params, desired)[1]


def bilevel_optimization(cost, dynamics, loss, x0, init_U, params, cost_args, dynamics_args,
                         loss_args, trajax_ilqr_kwargs=None):
    """policy/optimizers.py:34-75 for loss = L2MPC.loss: (high_level_loss, low_level_grad,
    high_level_grad, itr) from ONE kernel launch (gmpc_bilevel_l2: iLQR, loss gradient, (T m)^2 Hessian,
    LU solve, tangent rollout) plus the cost-MLP mixed VJP (policy/bilevel.py).  Unbatched shapes as in
    the reference, or batched (then every output, the gradient leaves included, has a leading batch
    axis -- what jax.vmap of this function returns at policy/base.py:122-125)."""
    from gan_mpc_b200.policy import bilevel
    pol = _owner(cost, "cost")
    if _owner(dynamics, "dynamics") is not pol or _owner(loss, "loss") is not pol:
        raise ValueError("cost, dynamics and loss must belong to the same policy")
    if len(dynamics_args) != 0:
        raise ValueError("MLP dynamics take no extra arguments (policy/eval.py:121)")
    (desired,) = loss_args
    batched, out = pol._bilevel(x0, init_U, params, cost_args[0], desired, trajax_ilqr_kwargs)
    g = bilevel.high_level_grad_tree(params, out, reduce_mean=False)
    res = (out["loss"], out["low_level_grad"], g, out["iteration"])
    if not batched:
        first = lambda t: ({k: first(v) for k, v in t.items()} if isinstance(t, dict)
                           else (t[0] if isinstance(t, torch.Tensor) else t))
        res = tuple(first(r) for r in res)
    return res


def cost_hessian_wrt_control(cost, dynamics, x0, U):
    """policy/optimizers.py:86-90 -- jax.hessian of `objective` w.r.t. U: [T,m,T,m] (or [B,T,m,T,m]).
    cost / dynamics are optimizers.bind(...) closures."""
    if not isinstance(cost, Bound) or not isinstance(dynamics, Bound):
        raise TypeError("cost_hessian_wrt_control() needs optimizers.bind(...) closures")
    pol, goal = cost.policy, cost.args[0]
    batched, out = pol._bilevel(x0, U, cost.params, goal, torch.zeros_like(goal),
                                dict(pol.trajax_ilqr_kwargs, maxiter=0), want_hessian=True)
    T, m = out["U"].shape[1:]
    Hs = out["hessian"].reshape(-1, T, m, T, m)
    return Hs if batched else Hs[0]


def cost_vjp(cost, dynamics, V, x0, U, params, cost_args):
    """policy/optimizers.py:93-105 -- grad_params ( V . grad_U objective(U; params) ), non-zero only
    on the cost side of params.  cost is the policy's bound method, dynamics a bind(...) closure,
    V is [T*m] (or [B,T*m])."""
    from gan_mpc_b200.policy import bilevel
    pol = _owner(cost, "cost")
    goal = cost_args[0]
    batched = x0.dim() == 2
    T, m = U.shape[-2:]
    Vb = (V if batched else V[None]).reshape(-1, T, m).to(pol.device, torch.float32).contiguous()
    _, out = pol._bilevel(x0, U, params, goal, torch.zeros_like(goal),
                          dict(pol.trajax_ilqr_kwargs, maxiter=0), V=Vb)
    g = bilevel.high_level_grad_tree(params, out, reduce_mean=False)
    if not batched:
        first = lambda t: ({k: first(v) for k, v in t.items()} if isinstance(t, dict)
                           else (t[0] if isinstance(t, torch.Tensor) else t))
        g = first(g)
    return g
