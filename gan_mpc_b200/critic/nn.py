"""Critic network shell (reference critic/nn.py:10-42): scanned OptimizedLSTMCell over the
trajectory, final h, (num_layers-1) x relu(Dense(H)), Dense(1).

Parameters live in the flax pytree layout; `flatten` / `unflatten` map to the flat fp32 vector of
include/gmpc.h (Wi | Wh | bh | {Dk, Db}* | Wo | bo, gate order i,f,g,o)."""

import numpy as np
import torch

from gan_mpc_b200 import base, synthetic

GATES = ("i", "f", "g", "o")
CELL = "ScanOptimizedLSTMCell_0"


class LSTM(base.BaseNN):
    def __init__(self, lstm_features, num_layers, num_hidden_units, fout=1):
        if fout != 1:
            raise NotImplementedError("the critic head is Dense(1) (critic/nn.py:14)")
        self.lstm_features = lstm_features
        self.num_layers = num_layers
        self.num_hidden_units = num_hidden_units
        self.fout = fout

    def get_init_params(self, seed, xsize):
        return (seed, xsize)

    def init(self, seed, xsize, device="cuda"):
        flat = synthetic.critic_params_flat(seed, xsize, self.lstm_features, self.num_layers,
                                            self.num_hidden_units)
        return self.unflatten(torch.from_numpy(flat).to(device), xsize)

    def param_count(self, xsize):
        F, H = self.lstm_features, self.num_hidden_units
        cnt, d = xsize * 4 * F + F * 4 * F + 4 * F, F
        for _ in range(self.num_layers - 1):
            cnt += d * H + H
            d = H
        return cnt + d + 1

    def unflatten(self, flat, xsize):
        F, H = self.lstm_features, self.num_hidden_units
        o = 0

        def take(*shape):
            nonlocal o
            cnt = int(np.prod(shape))
            t = flat[o:o + cnt].reshape(*shape)
            o += cnt
            return t

        Wi, Wh, bh = take(xsize, 4 * F), take(F, 4 * F), take(4 * F)
        cell = {}
        for gi, g in enumerate(GATES):
            cell["i" + g] = {"kernel": Wi[:, gi * F:(gi + 1) * F]}
            cell["h" + g] = {"kernel": Wh[:, gi * F:(gi + 1) * F], "bias": bh[gi * F:(gi + 1) * F]}
        p = {CELL: cell}
        d = F
        for l in range(self.num_layers - 1):
            p[f"Dense_{l}"] = {"kernel": take(d, H), "bias": take(H)}
            d = H
        p[f"Dense_{self.num_layers - 1}"] = {"kernel": take(d, 1), "bias": take(1)}
        assert o == flat.numel()
        return {"params": p}

    def flatten(self, params):
        p = params["params"]
        cell = p[CELL]
        parts = [torch.cat([cell["i" + g]["kernel"] for g in GATES], dim=1).reshape(-1),
                 torch.cat([cell["h" + g]["kernel"] for g in GATES], dim=1).reshape(-1),
                 torch.cat([cell["h" + g]["bias"] for g in GATES])]
        for l in range(self.num_layers):
            parts += [p[f"Dense_{l}"]["kernel"].reshape(-1), p[f"Dense_{l}"]["bias"].reshape(-1)]
        return torch.cat(parts).contiguous()
