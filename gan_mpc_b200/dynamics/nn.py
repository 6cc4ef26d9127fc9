"""Dynamics network shell (reference dynamics/nn.py:10-34): x' = Dense(n)(relu-MLP([x, u])) + x.

The forward/adjoint arithmetic lives in libgmpc (csrc/plan_*.cuh); this class only carries the
hyper-parameters and builds flax-layout parameter pytrees."""

import numpy as np
import torch

from gan_mpc_b200 import base, synthetic


def dense_stack_params(rng, dims, device):
    """{"params": {"Dense_i": {"kernel": [in,out], "bias": [out]}}} with flax default init."""
    Ws, bs = synthetic.mlp_params(rng, dims)
    return {"params": {f"Dense_{i}": {"kernel": torch.from_numpy(W).to(device),
                                      "bias": torch.from_numpy(b).to(device)}
                       for i, (W, b) in enumerate(zip(Ws, bs))}}


def dense_stack_lists(params):
    """flax pytree -> (kernels, biases) lists in call order."""
    p = params["params"]
    n = len(p)
    return ([p[f"Dense_{i}"]["kernel"] for i in range(n)], [p[f"Dense_{i}"]["bias"] for i in range(n)])


class MLP(base.BaseDynamicsNN):
    def __init__(self, num_layers, num_hidden_units, x_out):
        self.num_layers = num_layers
        self.num_hidden_units = num_hidden_units
        self.x_out = x_out

    def get_carry(self, x):
        """dynamics/nn.py:15-17 -- the MLP has no carry: empty trailing axis."""
        return torch.empty(*x.shape[:-1], 0, device=x.device, dtype=torch.float32)

    def get_init_params(self, seed, u_size):
        return (seed, self.x_out, u_size)

    def init(self, seed, x_size, u_size, device="cuda"):
        rng = np.random.Generator(np.random.PCG64(seed))
        dims = synthetic.dyn_dims(x_size, u_size, self.num_layers, self.num_hidden_units)
        return dense_stack_params(rng, dims, device)


class LSTM(MLP):
    """dynamics/nn.py:37-57 -- out of scope (the north star names the MLP)."""

    def __init__(self, *a, **k):
        raise NotImplementedError("LSTM dynamics are out of scope of the B200 hot path (use: 'mlp')")
