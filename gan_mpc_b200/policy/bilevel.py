"""Host side of the bilevel (cost-training) gradient -- policy/optimizers.py:34-105 and
policy/base.py:87-128 of the reference -- over libgmpc's gmpc_bilevel_l2.

The kernel does the iLQR solve, loss_grad_wrt_control, cost_hessian_wrt_control, the (T m)^2 solve,
the tangent rollout along H and the mpc_weights part of cost_vjp.  What is left of cost_vjp is the
derivative of  w2 * d/de ||f(x_T + e dx_T; theta)||^2  w.r.t. the cost-MLP weights theta: one
forward-over-reverse pass of a 3-layer MLP per sample, assembled here from x_T and dx_T as a handful of
batched matmuls on the device (plumbing-sized: a few times the cost MLP's weight count in FLOPs per sample).

Reference quirks kept: the + sign of the high-level gradient, gradients only
on the cost side of `params` (dynamics / expert / critic leaves are exactly zero)."""

import torch

from gan_mpc_b200.dynamics.nn import dense_stack_lists


def cost_mlp_mixed_vjp(cost_params, w2, xT, dxT, reduce=None):
    """Gradients of  w2 * phi,  phi = d/de ||f(x_T + e dx_T)||^2 = 2 f(x_T) . (Jf dx_T)  (cost/nn.py:23-29),
    w.r.t. the cost-MLP kernels and biases, by two explicit back-propagations through the ReLU masks:
    the cotangent 2 (Jf dx) through the primal network (activations a_l) and the cotangent 2 f(x) through
    the tangent network (tangent activations da_l; linear, no bias):
        dW_l = a_l (x) c_l + da_l (x) d_l,   db_l = c_l.
    reduce None: per-sample gradients [B, ...]; "sum" / "mean": reduced over the batch (plain matmuls)."""
    Ws, bs = dense_stack_lists(cost_params)
    L = len(Ws)
    a, da, masks = [xT], [dxT], []
    for l in range(L - 1):
        z = a[-1] @ Ws[l] + bs[l]
        mk = (z > 0).to(z.dtype)
        masks.append(mk)
        a.append(z * mk)
        da.append((da[-1] @ Ws[l]) * mk)
    y = a[-1] @ Ws[-1] + bs[-1]
    dy = da[-1] @ Ws[-1]
    c, d = 2.0 * w2 * dy, 2.0 * w2 * y        # cotangents of the last layer's output / tangent output
    gW, gb = [None] * L, [None] * L
    B = xT.shape[0]
    for l in range(L - 1, -1, -1):
        if reduce is None:
            gW[l] = a[l][:, :, None] * c[:, None, :] + da[l][:, :, None] * d[:, None, :]
            gb[l] = c
        else:
            gW[l] = a[l].t() @ c + da[l].t() @ d
            gb[l] = c.sum(0)
            if reduce == "mean":
                gW[l], gb[l] = gW[l] / B, gb[l] / B
        if l > 0:
            c = (c @ Ws[l].t()) * masks[l - 1]
            d = (d @ Ws[l].t()) * masks[l - 1]
    return gW, gb


def zeros_like_tree(tree, lead=()):
    if isinstance(tree, dict):
        return {k: zeros_like_tree(v, lead) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return torch.zeros(*lead, *tree.shape, dtype=tree.dtype, device=tree.device)
    return tree


def high_level_grad_tree(params, out, reduce_mean, handle=None):
    """params-shaped pytree of d (H . grad_U J) / d params from one gmpc_bilevel_l2 result `out`
    (policy/optimizers.py:69-71).  reduce_mean: True = leaf-wise batch mean (policy/base.py:126-127),
    "sum" = batch sum (data-parallel callers divide by the global batch after the all-reduce), False =
    every leaf keeps a leading batch axis (what vmap of bilevel_optimization returns).
    With `handle` (a _lib.Handle whose staged weights are `params`) the batch-reduced cost-MLP part runs in
    libgmpc (gmpc_cost_mixed_vjp): the trainers' path.  The per-sample variant (reduce_mean False, the
    functional API of policy/optimizers.py) stays on the torch formulation above."""
    B = out["X"].shape[0]
    w2 = torch.sigmoid(params["mpc_weights"][2])
    mode = None if not reduce_mean else ("sum" if reduce_mean == "sum" else "mean")
    xT = out["X"][:, -1].contiguous()
    if handle is not None and mode is not None:
        Ws, _ = dense_stack_lists(params["cost_params"])
        dims = [Ws[0].shape[0]] + [W.shape[1] for W in Ws]
        gW, gb = handle.cost_mixed_vjp(xT, out["dxT"].contiguous(), 1.0 if mode == "sum" else 1.0 / B, dims)
    else:
        gW, gb = cost_mlp_mixed_vjp(params["cost_params"], w2, xT, out["dxT"], mode)
    red = ((lambda t: t.sum(0)) if reduce_mean == "sum" else (lambda t: t.mean(0))) if reduce_mean else (lambda t: t)
    lead = () if reduce_mean else (B,)
    tree = {k: zeros_like_tree(v, lead) for k, v in params.items()}
    tree["mpc_weights"] = red(out["grad_mpc_weights"])
    tree["cost_params"] = {"params": {f"Dense_{i}": {"kernel": gW[i], "bias": gb[i]} for i in range(len(gW))}}
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
