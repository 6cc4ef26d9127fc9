"""Critic (discriminator) trainer with the reference's entry points (gan/critic_trainer.py:12-104).

get_dataset runs the fused planner on every sample; train_critic_parameters is the sequential
minibatch scan: gather (fused into the kernel) -> BCE loss + flat gradient -> [sum all-reduce
across ranks] -> clip_by_global_norm(100) + Adam.  Parameters are kept as one flat device vector
for the whole scan and unflattened into the flax pytree at the end."""

import torch

from gan_mpc_b200 import parallel, utils


def _gen(key, device):
    g = torch.Generator(device=device)
    g.manual_seed(int(key))
    return g


def get_dataset(policy, params, true_dataset, key):
    """critic_trainer.py:12-38: expert windows labelled +1, planner rollouts labelled -1,
    train part permuted.  `key` is an int seed (JAX threefry streams are not reproducible here)."""
    def func(X, true_Y):
        D = true_Y.shape[0]
        xsize = X.shape[-1]
        xc, *_ = policy.get_optimal_values(params, X)
        pred_Y = xc[..., :xsize]
        ones = torch.ones(D, dtype=torch.float32, device=pred_Y.device)
        return (torch.cat([true_Y.to(pred_Y.device, torch.float32), pred_Y], dim=0),
                torch.cat([ones, -ones], dim=0))

    true_train, true_test = true_dataset
    train_X, train_label = func(*true_train)
    test_X, test_label = func(*true_test)
    perm = torch.randperm(train_X.shape[0], generator=_gen(key, train_X.device), device=train_X.device)
    return (train_X[perm].contiguous(), train_label[perm].contiguous()), (test_X, test_label)


def calculate_loss(policy, params, dataset):
    """critic_trainer.py:41-45."""
    X, Y = dataset
    return policy.critic_loss(X, Y, params)


def train_critic_parameters(train_args, opt_state, params, perm, dataset):
    """critic_trainer.py:48-65: scan over the rows of `perm` [steps, batch] (int32, device).
    With torch.distributed initialised each rank takes its slice of every minibatch and the
    flat gradient is sum all-reduced before the identical update on every rank."""
    policy, opt = train_args
    X, Y = dataset
    n = X.shape[-1]
    h = policy.critic_handle(n)
    flat = policy.critic_flat(params).clone()
    rank, world = parallel.rank_world()
    Bc = perm.shape[1]
    lo, hi = parallel.shard_range(Bc, rank, world)
    perm = perm.to(torch.int32)
    if world == 1 and perm.shape[0] > 0:
        # one C-ABI call enqueues the whole scan (the reference's lax.scan is one executable too)
        mu, nu = opt.moments(opt_state, "critic_params", flat)
        losses = h.critic_train_scan(X, Y, perm.contiguous(), flat, mu, nu, step0=opt_state["count"],
                                     lr=opt.lr, max_norm=opt.max_norm, b1=opt.b1, b2=opt.b2, eps=opt.eps)
        opt_state["count"] += perm.shape[0]
        params = dict(params)
        params["critic_params"] = policy.critic_model.model.unflatten(flat, n)
        return params, opt_state, losses.mean()
    losses = []
    for s in range(perm.shape[0]):
        loss, g = h.critic_loss_grad(X, Y, flat, inv_count=1.0 / Bc, perm=perm[s, lo:hi].contiguous())
        if world > 1:
            parallel.allreduce_sum_(g)
            parallel.allreduce_sum_(loss)
        opt_state["count"] += 1
        opt.step_flat(h, opt_state, "critic_params", flat, g)
        losses.append(loss)
    params = dict(params)
    params["critic_params"] = policy.critic_model.model.unflatten(flat, n)
    return params, opt_state, torch.stack(losses).mean()


@utils.timeit
def train(train_args, opt_state, params, true_dataset, num_updates, batch_size, key, id):
    """critic_trainer.py:68-104.  Returns (params, opt_state, train_losses, test_losses) and,
    through timeit, the wall time in minutes."""
    del id
    policy, opt = train_args
    train_data, test_data = get_dataset(policy, params, true_dataset, key)
    datasize = train_data[0].shape[0]
    steps_per_update = datasize // batch_size
    g = _gen(int(key) + 1, train_data[0].device)
    train_losses, test_losses = [], []
    for _ in range(1, num_updates + 1):
        # jax.random.choice default: sampled WITH replacement (critic_trainer.py:88-90)
        perm = torch.randint(0, datasize, (steps_per_update, batch_size), generator=g,
                             device=train_data[0].device, dtype=torch.int32)
        params, opt_state, train_loss = train_critic_parameters(
            train_args=(policy, opt), opt_state=opt_state, params=params, perm=perm,
            dataset=train_data)
        test_loss = calculate_loss(policy=policy, params=params, dataset=test_data)
        train_losses.append(float(train_loss))
        test_losses.append(float(test_loss))
    return params, opt_state, train_losses, test_losses
