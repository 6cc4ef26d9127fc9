"""Runner entry points with the reference's names (gan/runner.py:13-212); see norm/runner.py."""

from gan_mpc_b200 import expert, utils
from gan_mpc_b200.gan import js_policy
from gan_mpc_b200.norm import runner as norm_runner
from gan_mpc_b200.policy import eval

get_optimizer = norm_runner.get_optimizer


def get_policy(config, x_size, u_size, expert_model=None):
    cost, _ = utils.get_cost_model(config)
    dynamics, _ = utils.get_dynamics_model(config, x_size)
    critic, _ = utils.get_critic_model(config)
    if expert_model == "network":   # the reference's expert proposal network (utils.get_expert_model)
        expert_model = utils.get_expert_model(config, x_size, u_size)
    elif expert_model is None:      # no checkpoint is shipped: seeded synthetic proposals by default
        expert_model = expert.SyntheticExpert(config, x_size, u_size, seed=config.seed)
    train_policy = js_policy.JS_MPC(config=config, cost_model=cost, dynamics_model=dynamics,
                                    expert_model=expert_model, critic_model=critic)
    eval_policy = eval.EvalMPC(config=config, cost_model=cost, dynamics_model=dynamics,
                               expert_model=expert_model)
    return train_policy, eval_policy, config.mpc


def get_params(policy, config, x_size, u_size, load_expert=True):
    """expert params: load_expert=True reads the trained checkpoint as the reference does
    (norm/runner.py get_params passes (True,)); False draws flax-default random weights."""
    seed = config.seed
    mpc_weights = tuple(config.mpc.model.cost.weights.to_dict().values())
    return policy.init(mpc_weights, (seed, x_size), (seed, u_size),
                       (True,) if load_expert else (False, seed, 1, 1, x_size), (seed, x_size))


def train(*args, **kwargs):
    raise NotImplementedError("gan.runner.train drives the dm_control simulator: outside the B200 hot path")


def run(config_path):
    raise NotImplementedError("gan.runner.run needs dm_control, an expert checkpoint and "
                              "trajectories.json, none of which the reference ships")
