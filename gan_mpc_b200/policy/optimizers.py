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


def _next_row(name):
    raise NotImplementedError(
        f"{name}: the bilevel (implicit-function) gradient -- (T*m)^2 Hessian, dense solve, mixed "
        "VJP -- is the next scope row (SURVEY.md 8f-2); it is not part of the fused planner path")


def bilevel_optimization(*a, **k):
    """policy/optimizers.py:34-75."""
    _next_row("bilevel_optimization")


def cost_hessian_wrt_control(*a, **k):
    """policy/optimizers.py:86-90."""
    _next_row("cost_hessian_wrt_control")


def cost_vjp(*a, **k):
    """policy/optimizers.py:93-105."""
    _next_row("cost_vjp")
