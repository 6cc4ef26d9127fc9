"""ExpertModel with the reference's interface (expert/expert_model.py:11-91) over libgmpc's
gmpc_expert_propose.  The two calls the policies make (policy/eval.py:96-106) --
get_history_carry then get_carry_next_state_and_action_seq with teacher_forcing=False -- are ONE
kernel launch: the first returns a lazy carry (the history itself), the second runs the network.
Unbatched ([h+1,n]) or batched ([B,h+1,n]) histories."""

import os

import torch

from gan_mpc_b200 import utils
from gan_mpc_b200.expert import nn as expert_nn


class ExpertModel:
    def __init__(self, config, model):
        self.config = config
        self.model = model
        self._policy = None

    def bind(self, policy):
        """the policy that owns the libgmpc handles (called by EvalMPC.__init__)."""
        self._policy = policy

    @staticmethod
    def get_model(model_config, x_size, u_size):
        """expert_model.py:16-37."""
        if model_config.use == "lstm":
            c = model_config.lstm
            model = expert_nn.ScanLSTM(lstm_features=c.lstm_features, num_layers=c.num_layers,
                                       num_hidden_units=c.num_hidden_units, x_out=x_size, u_out=u_size)
        elif model_config.use == "mlp":
            c = model_config.mlp
            model = expert_nn.ScanMLP(num_layers=c.num_layers, num_hidden_units=c.num_hidden_units,
                                      x_out=x_size, u_out=u_size)
        else:
            raise ValueError("Choose either mlp or lstm model.")
        return expert_nn.StateAction(model)

    def init(self, load_params, *args, device="cuda"):
        """expert_model.py:39-49: load trained_models/expert/<type>/<name>/<id>/params.npy or
        random-init with args = (seed, batch_size, seqlen, x_size)."""
        if load_params:
            config = self.config
            path = os.path.join("trained_models", "expert", config.env.type, config.env.expert.name,
                                str(config.mpc.model.expert.load_id), "params.npy")
            return utils.load_params(path, device=device)
        return self.model.init(*args, device=device)

    def get_zero_carry(self, history_x, xseq, params):
        del history_x, params
        b = xseq if xseq.dim() == 3 else xseq[None]
        return self.model.get_init_carry(b)

    def get_history_carry(self, history_x, xseq, params):
        """expert_model.py:60-71 -- lazy: the history is consumed inside the kernel."""
        del xseq, params
        return ("history", history_x)

    def propose(self, policy, history_x, params):
        """policy/eval.py:87-107 in one launch: history_x [h+1,n] or [B,h+1,n] -> (goal_xseq, init_useq)."""
        batched = history_x.dim() == 3
        md = self.model.model
        hx = (history_x if batched else history_x[None]).to(policy.device, torch.float32).contiguous()
        h = policy._handle(md.x_out, md.u_out)
        goal, useq = h.expert_propose(hx, md.flatten(params), md.lstm_features, md.num_layers,
                                      md.num_hidden_units)
        return (goal, useq) if batched else (goal[0], useq[0])

    def get_carry_next_state_and_action_seq(self, carry, xseq, params, teacher_forcing=False):
        """expert_model.py:73-91 with teacher_forcing=False: (carry, (next_xseq [T+1,n], useq [T,m]))."""
        if teacher_forcing or not (isinstance(carry, tuple) and carry and carry[0] == "history"):
            raise NotImplementedError("the fused expert kernel serves the planning call sequence "
                                      "(get_history_carry -> free-running proposal) only")
        if self._policy is None:
            raise RuntimeError("ExpertModel is not bound to a policy (EvalMPC binds it)")
        if xseq.shape[-2] != self._policy.config.mpc.horizon:
            raise ValueError("xseq length must be the MPC horizon")
        return carry, self.propose(self._policy, carry[1], params)
