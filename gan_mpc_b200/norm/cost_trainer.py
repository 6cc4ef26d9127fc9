"""Cost-parameter trainer entry points (reference norm/cost_trainer.py:12-93).

`calculate_loss` (plan every test sample, evaluate the policy loss) runs on the fused planner.
`train_cost_parameters` / `train` need the bilevel gradient (policy.loss_and_grad), the next scope
row: they keep the reference's signatures and raise until it lands."""

import torch

from gan_mpc_b200 import utils


def calculate_loss(policy, params, dataset):
    """cost_trainer.py:12-21 -- mean over the batch of policy.loss(plan(x))."""
    batch_x, batch_y = dataset
    pred_y, pred_u, *_ = policy.get_optimal_values(params, batch_x)
    return policy.loss(pred_y, pred_u, params, batch_y).mean()


def train_cost_parameters(train_args, opt_state, params, perm, dataset):
    """cost_trainer.py:24-48."""
    policy, opt = train_args
    policy.loss_and_grad(None, params, None)  # raises: bilevel gradient is the next scope row


@utils.timeit
def train(train_args, opt_state, params, dataset, num_updates, batch_size, polyak_factor, key, id):
    """cost_trainer.py:51-93 (signature kept; returns (params, opt_state, train_losses,
    test_losses) + minutes appended by timeit)."""
    del id
    policy, opt = train_args
    policy.loss_and_grad(None, params, None)
