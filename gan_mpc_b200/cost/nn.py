"""Cost network shell (reference cost/nn.py:9-29): ||Dense(fout)(relu-MLP(x))||^2."""

import numpy as np

from gan_mpc_b200 import base, synthetic
from gan_mpc_b200.dynamics.nn import dense_stack_params


class MLP(base.BaseCostNN):
    def __init__(self, num_layers, num_hidden_units, fout):
        self.num_layers = num_layers
        self.num_hidden_units = num_hidden_units
        self.fout = fout

    def get_init_params(self, seed, xc_size):
        return (seed, xc_size)

    def init(self, seed, xc_size, device="cuda"):
        rng = np.random.Generator(np.random.PCG64(seed + 17))
        dims = synthetic.cost_dims(xc_size, self.num_layers, self.num_hidden_units, self.fout)
        return dense_stack_params(rng, dims, device)

    def get_cost(self, params, x):
        raise NotImplementedError("the cost MLP is evaluated inside libgmpc (gmpc_objective_grad)")
