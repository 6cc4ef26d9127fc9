"""Expert network shells (reference expert/nn.py:10-163): ScanMLP / ScanLSTM wrapped by StateAction.

The arithmetic lives in libgmpc (csrc/expert.cuh); these classes carry the hyper-parameters, build
flax-layout parameter pytrees and map them to the flat fp32 vector of include/gmpc.h
(trunk | next_x head | action head; kernels [in,out]; LSTM gates i,f,g,o)."""

import numpy as np
import torch

from gan_mpc_b200 import base, synthetic

GATES = ("i", "f", "g", "o")


class ScanMLP(base.BaseNN):
    """expert/nn.py:63-95 -- scan of StackedMLPCell (:22-40)."""
    lstm_features = 0
    cell_name = "ScanStackedMLPCell_0"

    def __init__(self, num_layers, num_hidden_units, x_out, u_out):
        self.num_layers, self.num_hidden_units = num_layers, num_hidden_units
        self.x_out, self.u_out = x_out, u_out

    @property
    def head_layers(self):
        return self.num_layers - 1        # MLPCell(num_layers - 1, ...) (expert/nn.py:33-38)

    @property
    def trunk_out(self):
        return self.num_hidden_units

    def get_init_carry(self, batch_xseq):
        return (batch_xseq[:, 0],)

    def _shapes(self):
        """[(path, shape)] in flat order."""
        n, m, H, F = self.x_out, self.u_out, self.num_hidden_units, self.lstm_features
        out = []
        if F > 0:
            out += [(("lstm", "Wi"), (n, 4 * F)), (("lstm", "Wh"), (F, 4 * F)), (("lstm", "bh"), (4 * F,))]
        else:
            out += [(("Dense_0", "kernel"), (n, H)), (("Dense_0", "bias"), (H,))]
        for hd, dout_last in ((0, n), (1, m)):
            d = self.trunk_out
            for l in range(self.head_layers):
                dout = dout_last if l == self.head_layers - 1 else H
                out += [((f"MLPCell_{hd}", f"Dense_{l}", "kernel"), (d, dout)),
                        ((f"MLPCell_{hd}", f"Dense_{l}", "bias"), (dout,))]
                d = H
        return out

    def param_count(self):
        return int(sum(np.prod(s) for _, s in self._shapes()))

    def unflatten(self, flat):
        F = self.lstm_features
        cell, o = {}, 0
        for path, shape in self._shapes():
            cnt = int(np.prod(shape))
            t = flat[o:o + cnt].reshape(*shape)
            o += cnt
            if path[0] == "lstm":
                sub = cell.setdefault("OptimizedLSTMCell_0", {})
                for gi, g in enumerate(GATES):
                    if path[1] == "Wi":
                        sub.setdefault("i" + g, {})["kernel"] = t[:, gi * F:(gi + 1) * F]
                    elif path[1] == "Wh":
                        sub.setdefault("h" + g, {})["kernel"] = t[:, gi * F:(gi + 1) * F]
                    else:
                        sub.setdefault("h" + g, {})["bias"] = t[gi * F:(gi + 1) * F]
            else:
                d = cell
                for k in path[:-1]:
                    d = d.setdefault(k, {})
                d[path[-1]] = t
        assert o == flat.numel()
        return {"params": {"model": {self.cell_name: cell}}}

    def flatten(self, params):
        cell = params["params"]["model"][self.cell_name]
        parts = []
        if self.lstm_features > 0:
            c = cell["OptimizedLSTMCell_0"]
            parts += [torch.cat([c["i" + g]["kernel"] for g in GATES], dim=1).reshape(-1),
                      torch.cat([c["h" + g]["kernel"] for g in GATES], dim=1).reshape(-1),
                      torch.cat([c["h" + g]["bias"] for g in GATES])]
        else:
            parts += [cell["Dense_0"]["kernel"].reshape(-1), cell["Dense_0"]["bias"].reshape(-1)]
        for hd in (0, 1):
            for l in range(self.head_layers):
                dl = cell[f"MLPCell_{hd}"][f"Dense_{l}"]
                parts += [dl["kernel"].reshape(-1), dl["bias"].reshape(-1)]
        return torch.cat(parts).contiguous()


class ScanLSTM(ScanMLP):
    """expert/nn.py:98-131 -- scan of LSTMCell (:43-60)."""
    cell_name = "ScanLSTMCell_0"

    def __init__(self, lstm_features, num_layers, num_hidden_units, x_out, u_out):
        super().__init__(num_layers, num_hidden_units, x_out, u_out)
        self.lstm_features = lstm_features

    @property
    def head_layers(self):
        return self.num_layers            # MLPCell(num_layers, ...) (expert/nn.py:54-59)

    @property
    def trunk_out(self):
        return self.lstm_features

    def get_init_carry(self, batch_xseq):
        z = torch.zeros(batch_xseq.shape[0], self.lstm_features, device=batch_xseq.device)
        return ((z, z.clone()), batch_xseq[:, 0])


class StateAction(base.BaseNN):
    """expert/nn.py:134-163."""

    def __init__(self, model):
        self.model = model

    def get_init_carry(self, input):
        return self.model.get_init_carry(input)

    def get_init_params(self, seed, batch_size, seqlen, x_size):
        return (seed, batch_size, seqlen, x_size)

    def init(self, seed, batch_size=1, seqlen=1, x_size=None, device="cuda"):
        """flax defaults: lecun_normal Dense / LSTM input kernels, orthogonal recurrent kernels, zero
        biases (seeded numpy generator; JAX's threefry streams are not reproducible here)."""
        md = self.model
        flat = synthetic.expert_params_flat(seed, md._shapes(), md.lstm_features)
        return md.unflatten(torch.from_numpy(flat).to(device))
