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


def get_expert_model(config, x_size, u_size, model_config=None):
    """utils.py:216-227: the reference reads the model block from the saved expert run
    (trained_models/expert/<type>/<name>/<id>/config.json); when that file is absent (the reference
    ships none) the `expert_prediction.model` block of the YAML -- or `model_config` -- is used."""
    import os
    from gan_mpc_b200.expert import expert_model
    if model_config is None:
        saved = os.path.join("trained_models", "expert", str(config.env.type), str(config.env.expert.name),
                             str(config.mpc.model.expert.load_id), "config.json")
        if os.path.exists(saved):
            model_config = load_config.Config.from_dict(load_json(saved)["model"])
        else:
            model_config = config.expert_prediction.model
    nn_model = expert_model.ExpertModel.get_model(model_config=model_config, x_size=x_size, u_size=u_size)
    return expert_model.ExpertModel(config, nn_model)


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


# ------------------------------------------------------------------------------------------------
# On-disk format either side of the path (reference utils.py:117-156): <dir>/<run id>/config.json +
# params.npy, the params pytree pickled by np.save(..., allow_pickle=True).  Leaves are written as
# numpy arrays (a pickle of jax.Arrays needs jax to load; numpy leaves load everywhere, and
# jax.numpy accepts them unchanged), in the flax layout the kernels consume (Dense_i/{kernel[in,out],
# bias}), so weights trained by the reference flow in without conversion.
def check_or_create_dir(path):
    import os
    if not os.path.exists(path):
        os.makedirs(path, exist_ok=True)


def save_json(data, dir_path, basename):
    import json
    import os
    check_or_create_dir(dir_path)
    with open(os.path.join(dir_path, basename), "w") as fp:
        json.dump(data, fp, indent=4, sort_keys=True)


def load_json(path):
    import json
    with open(path, "r") as fp:
        return json.load(fp)


def _to_numpy_tree(tree):
    import numpy as np
    import torch
    if isinstance(tree, dict):
        return {k: _to_numpy_tree(v) for k, v in tree.items()}
    if isinstance(tree, torch.Tensor):
        return tree.detach().cpu().numpy()
    return np.asarray(tree) if hasattr(tree, "__array__") else tree


def _to_torch_tree(tree, device):
    import numpy as np
    import torch
    if hasattr(tree, "items"):  # dict / FrozenDict-like
        return {k: _to_torch_tree(v, device) for k, v in tree.items()}
    if hasattr(tree, "__array__"):
        return torch.from_numpy(np.ascontiguousarray(np.asarray(tree, dtype=np.float32))).to(device)
    return tree


def save_all_args(dir_path, params, model_config, *other_json_args):
    """utils.py:135-148: next free integer run id under dir_path; config.json, params.npy and the
    extra (json_data, name) pairs.  Returns the run directory."""
    import os
    import numpy as np
    check_or_create_dir(dir_path)
    dir_list = sorted((d for d in os.listdir(dir_path) if d.isdigit()), key=lambda x: -int(x))
    key = "0" if not dir_list else f"{int(dir_list[0]) + 1}"
    full_path = os.path.join(dir_path, key)
    save_json(model_config, full_path, "config.json")
    np.save(os.path.join(full_path, "params.npy"), _to_numpy_tree(params), allow_pickle=True)
    for json_data, name in other_json_args:
        save_json(json_data, full_path, name)
    return full_path


def load_params(params_path, from_np=True, device="cuda"):
    """utils.py:151-156: the pickled pytree back as torch tensors on `device`."""
    import numpy as np
    if not from_np:
        raise NotImplementedError("params must be saved using numpy.")
    return _to_torch_tree(np.load(params_path, allow_pickle=True).item(), device)
