"""Cost model shell (reference cost/cost_model.py:11-42): pseudo-Huber staging cost on (u, x-goal)
with sigmoid-squashed weights, learned terminal cost selected at t == horizon."""

from gan_mpc_b200 import base


class MujocoBasedModel(base.BaseCostModel):
    def __init__(self, config, model):
        super().__init__(config)
        self.model = model

    def init(self, *args, device="cuda"):
        return self.model.init(*self.model.get_init_params(*args), device=device)

    def get_cost(self, xc, u, t, params, weights, goal_X):
        raise NotImplementedError(
            "MujocoBasedModel.get_cost is a structured closure: the per-step cost is fused inside "
            "libgmpc (gmpc_objective_grad / gmpc_plan); use policy.optimizers.objective")
