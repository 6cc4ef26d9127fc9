"""Dynamics trainer entry points (reference norm/dynamics_trainer.py:13-120).

predict_loss / train_per_update / train_params run on libgmpc: the rollout over the window, the
discounted loss and the back-propagation through time are gmpc_dynamics_fit (csrc/dynfit.cuh); the
contraction that is left for the weight gradient, dW_l = act_l cot_l^T (and the bias gradient, the row sums of
cot_l), is gmpc_gemm_nt (csrc/smallgemm.cuh); the update is optax chain(clip_by_global_norm(100), adam) on the
leaves labelled "tx" through gmpc_clip_adam_step.  `train` (:123-194) interleaves simulator
episodes (dm_control) with train_params and is outside the B200 hot path."""

import torch

from gan_mpc_b200 import utils
from gan_mpc_b200.dynamics.nn import dense_stack_lists
from gan_mpc_b200.norm import cost_trainer


def _fit(policy, params, xseq, useq, next_xseq, discount_factor, teacher_forcing):
    batched = xseq.dim() == 3
    f = lambda t: (t if batched else t[None]).to(policy.device, torch.float32).contiguous()
    xs, us, ys = f(xseq), f(useq), f(next_xseq)
    h = policy._handle(xs.shape[-1], us.shape[-1])
    policy._stage(h, params)
    Ws, _ = dense_stack_lists(params["dynamics_params"])
    dims = [Ws[0].shape[0]] + [W.shape[1] for W in Ws]
    return batched, h.dynamics_fit(xs, us, ys, discount_factor, teacher_forcing, dims)


def predict_loss(policy, params, xseq, useq, next_xseq, discount_factor, teacher_forcing):
    """dynamics_trainer.py:13-44: sum_t discount^t |x'_t - next_xseq[t]|^2 over a window [S,n]
    (or a batch of windows [B,S,n] -> [B])."""
    batched, (loss, _, _) = _fit(policy, params, xseq, useq, next_xseq, discount_factor, teacher_forcing)
    return loss if batched else loss[0]


def loss_and_grad(policy, params, batch_x, batch_u, batch_y, discount_factor, teacher_forcing):
    """value_and_grad of the batch-mean predict_loss w.r.t. the whole params pytree
    (dynamics_trainer.py:64-79): only the dynamics leaves are non-zero."""
    from gan_mpc_b200.policy import bilevel
    _, (loss, act, cot) = _fit(policy, params, batch_x, batch_u, batch_y, discount_factor, teacher_forcing)
    B = loss.shape[0]
    h = policy._handle(batch_x.shape[-1], batch_u.shape[-1])
    grads = bilevel.zeros_like_tree(params)
    dp = grads["dynamics_params"]["params"]
    for l in range(len(act)):
        dW, db = h.gemm_nt(act[l], cot[l], alpha=1.0 / B, want_rowsum=True)
        dp[f"Dense_{l}"] = {"kernel": dW, "bias": db}
    return loss.mean(), grads


def train_per_update(train_args, opt_state, params, perm, dataset, discount_factor, teacher_forcing):
    """dynamics_trainer.py:47-84: scan over the rows of perm [steps, batch]."""
    policy, opt = train_args
    X, U, Y = dataset
    params = utils.tree_clone(params)
    leaves = cost_trainer._trained_leaves(opt, params)
    flat = torch.cat([t.reshape(-1) for t in leaves]) if leaves else None
    losses = []
    for s in range(perm.shape[0]):
        p = perm[s].long()
        loss, grads = loss_and_grad(policy, params, X[p], U[p], Y[p], discount_factor, teacher_forcing)
        losses.append(loss)
        if flat is None:
            continue
        g = torch.cat([t.reshape(-1) for t in cost_trainer._trained_leaves(opt, grads)])
        opt_state["count"] += 1
        h = next(iter(policy._handles.values()))
        opt.step_flat(h, opt_state, "dynamics_trainer", flat, g)
        o = 0
        for t in leaves:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return params, opt_state, torch.stack(losses).mean()


def train_params(train_args, opt_state, params, dataset, num_updates, batch_size, discount_factor,
                 teacher_forcing_factor, key, id):
    """dynamics_trainer.py:87-120: minibatch indices WITH replacement; teacher forcing for the first
    num_updates * teacher_forcing_factor updates ((id + up) <= ...).  `key` is an int seed."""
    datasize = dataset[0].shape[0]
    steps_per_update = datasize // batch_size
    g = torch.Generator(device=dataset[0].device)
    g.manual_seed(int(key))
    train_losses = []
    for up in range(1, num_updates + 1):
        perm = torch.randint(0, datasize, (steps_per_update, batch_size), generator=g, device=dataset[0].device)
        teacher_forcing = (id + up) <= (num_updates * teacher_forcing_factor)
        params, opt_state, train_loss = train_per_update(
            train_args=train_args, opt_state=opt_state, params=params, perm=perm, dataset=dataset,
            discount_factor=discount_factor, teacher_forcing=teacher_forcing)
        train_losses.append(float(train_loss))
    return params, opt_state, train_losses


def train(*args, **kwargs):
    raise NotImplementedError("dynamics_trainer.train interleaves dm_control episodes with train_params: "
                              "the simulator is outside the B200 hot path (call train_params on recorded windows)")
