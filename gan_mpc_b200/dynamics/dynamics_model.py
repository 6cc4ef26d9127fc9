"""DynamicsModel shell (reference dynamics/dynamics_model.py:11-48)."""

import torch

from gan_mpc_b200 import base


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
        raise NotImplementedError(
            "DynamicsModel.predict is a structured closure: the step is fused inside libgmpc "
            "(gmpc_rollout / gmpc_plan); use policy.optimizers / the policy classes")
