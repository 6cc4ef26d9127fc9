"""Cost-parameter trainer entry points (reference norm/cost_trainer.py:12-93).

`calculate_loss` plans every test sample and evaluates the policy loss; `train_cost_parameters` is
the minibatch scan over policy.loss_and_grad (the bilevel gradient, gmpc_bilevel_l2) with the
clipped-Adam update of the trained leaves (gmpc_clip_adam_step); `train` adds the Polyak blend."""

import torch

from gan_mpc_b200 import utils


def calculate_loss(policy, params, dataset):
    """cost_trainer.py:12-21 -- mean over the batch of policy.loss(plan(x))."""
    batch_x, batch_y = dataset
    pred_y, pred_u, *_ = policy.get_optimal_values(params, batch_x)
    return policy.loss(pred_y, pred_u, params, batch_y).mean()


def _trained_leaves(opt, params):
    """tensors of the entries labelled "tx" (norm/runner.py:46-58), in a fixed order."""
    def leaves(tree):
        if isinstance(tree, dict):
            for k in sorted(tree):
                yield from leaves(tree[k])
        elif isinstance(tree, torch.Tensor):
            yield tree
    return [t for k in opt.trained for t in leaves(params[k])]


def train_cost_parameters(train_args, opt_state, params, perm, dataset):
    """cost_trainer.py:24-48 -- scan over the rows of `perm` [steps, batch]: gather the minibatch,
    policy.loss_and_grad (the bilevel gradient), optax chain(clip_by_global_norm(100), adam) on the
    leaves labelled "tx" (one flat vector through gmpc_clip_adam_step; the global norm is over those
    leaves only), apply_updates.  Inputs are not mutated."""
    policy, opt = train_args
    X, Y = dataset
    params = utils.tree_clone(params)
    leaves = _trained_leaves(opt, params)
    flat = torch.cat([t.reshape(-1) for t in leaves]) if leaves else None
    losses = []
    for s in range(perm.shape[0]):
        p = perm[s].long()
        loss, grads = policy.loss_and_grad(X[p], params, (Y[p],))
        losses.append(loss)
        if flat is None:
            continue
        g = torch.cat([t.reshape(-1) for t in _trained_leaves(opt, grads)])
        opt_state["count"] += 1
        h = next(iter(policy._handles.values()))
        opt.step_flat(h, opt_state, "cost_trainer", flat, g)
        o = 0
        for t in leaves:  # write the updated flat vector back into the (cloned) pytree
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return params, opt_state, torch.stack(losses).mean()


@utils.timeit
def train(train_args, opt_state, params, dataset, num_updates, batch_size, polyak_factor, key, id):
    """cost_trainer.py:51-93: returns (params, opt_state, train_losses, test_losses) + minutes
    appended by timeit.  `key` is an int seed (JAX threefry streams are not reproducible here);
    minibatch indices are sampled WITH replacement (jax.random.choice default, :72-74); all leaves
    are blended prev * polyak + new * (1 - polyak) after the updates (:88-92)."""
    del id
    policy, opt = train_args
    train_data, test_data = dataset
    prev_params = params
    datasize = train_data[0].shape[0]
    steps_per_update = datasize // batch_size
    g = torch.Generator(device=train_data[0].device)
    g.manual_seed(int(key))
    train_losses, test_losses = [], []
    for _ in range(1, num_updates + 1):
        perm = torch.randint(0, datasize, (steps_per_update, batch_size), generator=g,
                             device=train_data[0].device)
        params, opt_state, train_loss = train_cost_parameters(
            train_args=(policy, opt), opt_state=opt_state, params=params, perm=perm, dataset=train_data)
        test_loss = calculate_loss(policy=policy, params=params, dataset=test_data)
        train_losses.append(float(train_loss))
        test_losses.append(float(test_loss))
    params = utils.tree_map2(lambda x, y: polyak_factor * x + (1 - polyak_factor) * y, prev_params, params)
    return params, opt_state, train_losses, test_losses
