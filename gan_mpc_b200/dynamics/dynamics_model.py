"""DynamicsModel (reference dynamics/dynamics_model.py:11-48).  Inside a plan the step is fused into the
planner kernels; `predict` evaluates ONE step through the same kernels (gmpc_rollout on a horizon-1 handle)."""

import torch

from gan_mpc_b200 import _lib, base
from gan_mpc_b200.dynamics.nn import dense_stack_lists


class DynamicsModel(base.BaseDynamicsModel):
    def __init__(self, config, model):
        super().__init__(config)
        self.model = model

    def init(self, *args, device="cuda"):
        return self.model.init(*self.model.get_init_params(*args), device=device)

    def get_zero_carry(self, history_x):
        return self.model.get_carry(torch.zeros(history_x.shape[1], device=history_x.device))

    def get_history_carry(self, history_x, history_u, params):
        """dynamics_model.py:24-43 -- replaying history only updates the carry, which is empty
        for the MLP."""
        return self.get_zero_carry(history_x)

    def predict(self, xc, u, t, params):
        """dynamics_model.py:45-48: xc [n] (or [B,n]), u [m] -> next xc (the MLP has no carry, `t` is unused)."""
        batched = xc.dim() == 2
        x = (xc if batched else xc[None]).float().contiguous()
        uu = (u if batched else u[None]).float().contiguous()
        n, m, dev = x.shape[1], uu.shape[1], x.device
        if not hasattr(self, "_handles"):
            self._handles = {}
        key = (n, m, dev.index)
        if key not in self._handles:
            d = self.model
            self._handles[key] = _lib.Handle(n, m, 1, d.num_layers, d.num_hidden_units, 1, 1, 1, device=dev.index)
        h = self._handles[key]
        dW, db = dense_stack_lists(params)
        z = lambda *s_: torch.zeros(*s_, device=dev)
        h.set_weights([t_.contiguous() for t_ in dW], [t_.contiguous() for t_ in db], [z(n, 1)], [z(1)], z(3))
        X = h.rollout(x, uu[:, None, :].contiguous())
        return X[:, 1] if batched else X[0, 1]
