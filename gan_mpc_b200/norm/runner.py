"""Runner entry points with the reference's names (norm/runner.py:13-177).

get_policy / get_params / get_optimizer are drop-in.  train / run interleave dm_control episodes with
the trainers and need a pre-trained expert checkpoint and trajectories.json -- none of which the
reference ships: they keep their signatures and raise.  Everything they call between two simulator
episodes is built: cost_trainer.train, dynamics_trainer.train_params, (gan) critic_trainer.train."""

from gan_mpc_b200 import expert, optim, utils
from gan_mpc_b200.norm import l2_policy
from gan_mpc_b200.policy import eval


def get_policy(config, x_size, u_size, expert_model=None):
    cost, _ = utils.get_cost_model(config)
    dynamics, _ = utils.get_dynamics_model(config, x_size)
    if expert_model == "network":   # the reference's expert proposal network (utils.get_expert_model)
        expert_model = utils.get_expert_model(config, x_size, u_size)
    elif expert_model is None:      # no checkpoint is shipped: seeded synthetic proposals by default
        expert_model = expert.SyntheticExpert(config, x_size, u_size, seed=config.seed)
    train_policy = l2_policy.L2MPC(config=config, cost_model=cost, dynamics_model=dynamics,
                                   expert_model=expert_model)
    eval_policy = eval.EvalMPC(config=config, cost_model=cost, dynamics_model=dynamics,
                               expert_model=expert_model)
    return train_policy, eval_policy, config.mpc


def get_params(policy, config, x_size, u_size, load_expert=True):
    """expert params: load_expert=True reads the trained checkpoint as the reference does
    (norm/runner.py get_params passes (True,)); False draws flax-default random weights."""
    seed = config.seed
    mpc_weights = tuple(config.mpc.model.cost.weights.to_dict().values())
    return policy.init(mpc_weights, (seed, x_size), (seed, u_size),
                       (True,) if load_expert else (False, seed, 1, 1, x_size))


def get_optimizer(params, masked_vars, lr):
    labels = utils.get_masked_labels(all_vars=params.keys(), masked_vars=masked_vars,
                                     tx_key="tx", zero_key="zero")
    opt = optim.MaskedClipAdam(lr, labels, max_norm=100.0)
    return opt, opt.init(params)


def train(*args, **kwargs):
    raise NotImplementedError("norm.runner.train drives the dm_control simulator: outside the B200 hot path "
                              "(SURVEY.md section 2, #13); call cost_trainer.train / dynamics_trainer.train_params "
                              "on recorded data")


def run(config_path, dataset_path=None):
    raise NotImplementedError("norm.runner.run needs dm_control, an expert checkpoint and "
                              "trajectories.json, none of which the reference ships")
