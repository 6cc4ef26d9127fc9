"""Cost model (reference cost/cost_model.py:11-42): pseudo-Huber staging cost on (u, x - goal) with
sigmoid-squashed weights, learned terminal cost w2 |MLP(x)|^2 selected at t == horizon.

Inside a plan the cost is fused into the planner kernels.  `get_cost` evaluates ONE step cost through the
same kernels: a horizon-1 handle whose dynamics layer is zero (so x_1 = x_0) gives
    objective(x, [u], [goal, x])  with the terminal weight switched off   = staging cost of (x, u, goal)
    objective(x, [0], [x,   x])                                           = terminal cost of x
(the other term is exactly h(0) = sqrt(a^2) - a, at most one ulp of a = 1e-2)."""

import torch

from gan_mpc_b200 import _lib, base
from gan_mpc_b200.dynamics.nn import dense_stack_lists


class MujocoBasedModel(base.BaseCostModel):
    def __init__(self, config, model):
        super().__init__(config)
        self.model = model
        self._handles = {}

    def init(self, *args, device="cuda"):
        return self.model.init(*self.model.get_init_params(*args), device=device)

    def _handle(self, n, m, device):
        key = (n, m, device.index)
        if key not in self._handles:
            c = self.model
            self._handles[key] = _lib.Handle(n, m, 1, 1, 1, c.num_layers, c.num_hidden_units, c.fout,
                                             device=device.index)
        return self._handles[key]

    def get_cost(self, xc, u, t, params, weights, goal_X):
        """cost/cost_model.py:33-42.  xc [n] (or [B,n]), u [m], t int, params = cost params pytree,
        weights = raw mpc_weights [3], goal_X [T+1,n] (or [B,T+1,n]) -> scalar (or [B])."""
        batched = xc.dim() == 2
        x = (xc if batched else xc[None]).float().contiguous()
        uu = (u if batched else u[None]).float().contiguous()
        g = (goal_X if goal_X.dim() == 3 else goal_X[None]).float()
        B, n, m, dev = x.shape[0], x.shape[1], uu.shape[1], x.device
        h = self._handle(n, m, dev)
        cW, cb = dense_stack_lists(params)
        z = lambda *s: torch.zeros(*s, device=dev)
        terminal = int(t) == int(self.config.mpc.horizon)
        w = weights.float().clone()
        if not terminal:
            w[2] = -1.0e4                      # sigmoid -> 0: staging cost only
        h.set_weights([z(n + m, n)], [z(n)], [t_.contiguous() for t_ in cW], [t_.contiguous() for t_ in cb], w)
        if terminal:
            goal2 = torch.stack([x, x], dim=1)
            U = z(B, 1, m)
        else:
            goal2 = torch.stack([g[:, int(t)].expand(B, n), x], dim=1)
            U = uu[:, None, :]
        J, _, _, _ = h.objective_grad(x, U.contiguous(), goal2.contiguous(), want_grad=False, want_X=False)
        return J if batched else J[0]
