"""Stand-in for the reference's expert proposal network (expert/expert_model.py:60-91), which is
out of scope (it needs a pre-trained checkpoint the reference does not ship).  It exposes the two
methods EvalMPC.get_goal_states_init_actions calls and produces the SURVEY.md 8d synthetic
proposals: goal[0] = x then a 0.1-step random walk, actions = tanh(N(0,1))."""

import torch


class SyntheticExpert:
    def __init__(self, config, x_size, u_size, seed=0):
        self.config, self.x_size, self.u_size, self.seed = config, x_size, u_size, seed

    def init(self, *args):
        return {}

    def get_history_carry(self, history_x, xseq, params):
        del xseq, params
        return (history_x[..., -1, :],)

    def get_carry_next_state_and_action_seq(self, carry, xseq, params, teacher_forcing=False):
        del params, teacher_forcing
        x = carry[-1]
        T = xseq.shape[-2]
        g = torch.Generator(device=x.device)
        g.manual_seed(self.seed)
        steps = 0.1 * torch.randn(*x.shape[:-1], T, self.x_size, generator=g, device=x.device)
        goal = torch.cat([x[..., None, :], x[..., None, :] + torch.cumsum(steps, dim=-2)], dim=-2)
        useq = torch.tanh(torch.randn(*x.shape[:-1], T, self.u_size, generator=g, device=x.device))
        return carry, (goal, useq)
