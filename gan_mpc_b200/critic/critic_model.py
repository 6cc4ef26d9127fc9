"""CriticModel with the reference's name and constructor (critic/critic_model.py:6-16).  The LSTM
discriminator is evaluated by libgmpc (csrc/critic.cuh, gmpc_critic_forward); this object owns the
network shell (hyper-parameters, flat <-> pytree layout), builds initial parameters and scores
trajectories."""

import torch

from gan_mpc_b200 import _lib, base


class CriticModel(base.BaseCriticModel):
    def __init__(self, config, model):
        super().__init__(config)
        self.model = model
        self._handles = {}

    def init(self, *args, device="cuda"):
        """args = (seed, x_size), as assembled by gan.runner.get_params."""
        init_args = self.model.get_init_params(*args)
        return self.model.init(*init_args, device=device)

    def predict(self, xseq, params):
        """critic/critic_model.py:15-16: xseq [T1,n] -> logit [1]  (or [B,T1,n] -> [B,1])."""
        batched = xseq.dim() == 3
        x = (xseq if batched else xseq[None]).float().contiguous()
        n, dev = x.shape[-1], x.device
        key = (n, dev.index)
        if key not in self._handles:
            c = self.model
            self._handles[key] = _lib.Handle(n, 1, max(1, x.shape[1] - 1), 1, 1, 1, 1, 1,
                                             critic_features=c.lstm_features, critic_layers=c.num_layers,
                                             critic_hidden=c.num_hidden_units, device=dev.index)
        logit = self._handles[key].critic_forward(x, self.model.flatten(params).float())
        return logit[:, None] if batched else logit
