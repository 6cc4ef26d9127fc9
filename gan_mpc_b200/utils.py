"""Host-side helpers with the reference's names (utils.py:26-35,159-213): timeit, model factories,
mask labels.  Simulator / video / file IO helpers of the reference are out of scope."""

import time

from gan_mpc_b200.config import load_config
from gan_mpc_b200.cost import cost_model
from gan_mpc_b200.cost import nn as cost_nn
from gan_mpc_b200.critic import critic_model
from gan_mpc_b200.critic import nn as critic_nn
from gan_mpc_b200.dynamics import dynamics_model
from gan_mpc_b200.dynamics import nn as dynamics_nn


def timeit(fn):
    """utils.py:26-35 -- appends the wall time in minutes to the return tuple."""
    def wrapper_fn(*args, **kwargs):
        start_time = time.time()
        ret = fn(*args, **kwargs)
        exe_time = (time.time() - start_time) / 60
        if isinstance(ret, tuple):
            return *ret, exe_time
        return ret, exe_time
    return wrapper_fn


def get_config(config_path):
    return load_config.Config.from_yaml(config_path)


def get_masked_labels(all_vars, masked_vars, tx_key, zero_key):
    return {v: (zero_key if v in masked_vars else tx_key) for v in all_vars}


def get_cost_model(config):
    model_config = config.mpc.model.cost
    mlp = model_config.mlp
    nn_model = cost_nn.MLP(num_layers=mlp.num_layers, num_hidden_units=mlp.num_hidden_units,
                           fout=mlp.fout)
    return cost_model.MujocoBasedModel(config, nn_model), model_config


def get_dynamics_model(config, x_size):
    model_config = config.mpc.model.dynamics
    if model_config.use == "mlp":
        mlp = model_config.mlp
        nn_model = dynamics_nn.MLP(num_layers=mlp.num_layers,
                                   num_hidden_units=mlp.num_hidden_units, x_out=x_size)
    elif model_config.use == "lstm":
        raise NotImplementedError("LSTM dynamics are out of scope of the B200 hot path")
    else:
        raise ValueError("Choose either mlp or lstm model.")
    return dynamics_model.DynamicsModel(config, nn_model), model_config


def get_critic_model(config):
    model_config = config.mpc.model.critic
    if model_config.use == "lstm":
        lstm = model_config.lstm
        nn_model = critic_nn.LSTM(lstm_features=lstm.lstm_features, num_layers=lstm.num_layers,
                                  num_hidden_units=lstm.num_hidden_units)
    else:
        raise ValueError("Choose lstm model.")
    return critic_model.CriticModel(config, nn_model), model_config


def tree_clone(tree):
    """deep copy of a params pytree (dict of dicts of tensors); non-tensor leaves are shared."""
    import torch
    if isinstance(tree, dict):
        return {k: tree_clone(v) for k, v in tree.items()}
    return tree.clone() if isinstance(tree, torch.Tensor) else tree


def tree_map2(fn, a, b):
    """jax.tree_map over two pytrees of the same structure (tensor leaves only)."""
    import torch
    if isinstance(a, dict):
        return {k: tree_map2(fn, a[k], b[k]) for k in a}
    return fn(a, b) if isinstance(a, torch.Tensor) else a
