"""Host side of the bilevel (cost-training) gradient -- policy/optimizers.py:34-105 and
policy/base.py:87-128 of the reference -- over libgmpc's gmpc_bilevel_l2.

The kernel does the iLQR solve, loss_grad_wrt_control, cost_hessian_wrt_control, the (T m)^2 solve,
the tangent rollout along H and the mpc_weights part of cost_vjp.  What is left of cost_vjp is the
derivative of  w2 * d/de ||f(x_T + e dx_T; theta)||^2  w.r.t. the cost-MLP weights theta: one
forward-over-reverse pass of a 3-layer MLP per sample, assembled here from x_T and dx_T with
torch.func on the device (plumbing-sized: 2 x the cost MLP's weight count in FLOPs per sample).

Reference quirks kept: the + sign of the high-level gradient, gradients only
on the cost side of `params` (dynamics / expert / critic leaves are exactly zero)."""

import torch
from torch.func import grad, jvp, vmap

from gan_mpc_b200.dynamics.nn import dense_stack_lists


def _mlp(x, Ws, bs):
    z = x
    for W, b in zip(Ws[:-1], bs[:-1]):
        z = torch.relu(z @ W + b)
    return z @ Ws[-1] + bs[-1]


def _phi(Ws, bs, x, dx):
    """d/de ||f(x + e dx)||^2 = 2 f(x) . (Jf dx)   (cost/nn.py:23-29)."""
    y, yd = jvp(lambda xx: _mlp(xx, Ws, bs), (x,), (dx,))
    return 2.0 * (y * yd).sum()


def cost_mlp_mixed_vjp(cost_params, w2, xT, dxT):
    """per-sample gradients of w2 * _phi w.r.t. the cost MLP leaves: lists of [B, ...] tensors."""
    Ws, bs = dense_stack_lists(cost_params)
    gW, gb = vmap(grad(_phi, argnums=(0, 1)), in_dims=(None, None, 0, 0))(tuple(Ws), tuple(bs), xT, dxT)
    return [w2 * g for g in gW], [w2 * g for g in gb]


def zeros_like_tree(tree, lead=()):
    if isinstance(tree, dict):
        return {k: zeros_like_tree(v, lead) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return torch.zeros(*lead, *tree.shape, dtype=tree.dtype, device=tree.device)
    return tree


def high_level_grad_tree(params, out, reduce_mean):
    """params-shaped pytree of d (H . grad_U J) / d params from one gmpc_bilevel_l2 result `out`
    (policy/optimizers.py:69-71).  reduce_mean: True = leaf-wise batch mean (policy/base.py:126-127),
    "sum" = batch sum (data-parallel callers divide by the global batch after the all-reduce), False =
    every leaf keeps a leading batch axis (what vmap of bilevel_optimization returns)."""
    B = out["X"].shape[0]
    w2 = torch.sigmoid(params["mpc_weights"][2])
    gW, gb = cost_mlp_mixed_vjp(params["cost_params"], w2, out["X"][:, -1].contiguous(), out["dxT"])
    red = ((lambda t: t.sum(0)) if reduce_mean == "sum" else (lambda t: t.mean(0))) if reduce_mean else (lambda t: t)
    lead = () if reduce_mean else (B,)
    tree = {k: zeros_like_tree(v, lead) for k, v in params.items()}
    tree["mpc_weights"] = red(out["grad_mpc_weights"])
    tree["cost_params"] = {"params": {f"Dense_{i}": {"kernel": red(gW[i]), "bias": red(gb[i])}
                                      for i in range(len(gW))}}
    return tree


def tree_leaves(tree):
    if isinstance(tree, dict):
        for k in sorted(tree):
            yield from tree_leaves(tree[k])
    elif isinstance(tree, torch.Tensor):
        yield tree


def allreduce_mean_tree_(loss_sum, tree, global_batch):
    """Data-parallel reduction of (sum of losses, sum of per-sample gradients) over the ranks: ONE
    sum all-reduce of the flattened pytree (NCCL on the box), then the division by the global batch
    -- the batch mean of policy/base.py:126-127 with the batch sharded over GPUs."""
    from gan_mpc_b200 import parallel
    leaves = list(tree_leaves(tree))
    flat = torch.cat([loss_sum.reshape(1)] + [t.reshape(-1) for t in leaves])
    parallel.allreduce_sum_(flat)
    flat /= float(global_batch)
    o = 1
    for t in leaves:
        t.copy_(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    return flat[0].clone(), tree
