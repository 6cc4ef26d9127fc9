"""CPU restatement of the expert proposal network as the planner calls it.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED.

expert/nn.py:22-60 (cells), :63-131 (scans), expert/expert_model.py:60-91 (history carry, then the
free-running proposal), policy/eval.py:87-107 (the call sequence).  Parameters come as the flat
vector of include/gmpc.h; `shapes` is [(path, shape)] in flat order (gan_mpc_b200/expert/nn.py)."""

import torch


def _take(flat, shapes):
    out, o = {}, 0
    for path, shape in shapes:
        cnt = 1
        for s in shape:
            cnt *= s
        out[path] = flat[o:o + cnt].reshape(*shape)
        o += cnt
    assert o == flat.numel()
    return out


def _head(p, hd, layers, y):
    a = y
    for l in range(layers - 1):
        a = torch.relu(a @ p[(f"MLPCell_{hd}", f"Dense_{l}", "kernel")] + p[(f"MLPCell_{hd}", f"Dense_{l}", "bias")])
    l = layers - 1
    return a @ p[(f"MLPCell_{hd}", f"Dense_{l}", "kernel")] + p[(f"MLPCell_{hd}", f"Dense_{l}", "bias")]


def propose(history_x, flat, shapes, F, head_layers, T):
    """history_x [B,h+1,n] -> goal_xseq [B,T+1,n], init_useq [B,T,m]."""
    p = _take(flat, shapes)
    B = history_x.shape[0]
    hist = history_x.shape[1] - 1
    if F > 0:
        c = torch.zeros(B, F, dtype=history_x.dtype)
        h = torch.zeros(B, F, dtype=history_x.dtype)

        def lstm(x, c, h):  # flax OptimizedLSTMCell, gates i,f,g,o
            z = x @ p[("lstm", "Wi")] + h @ p[("lstm", "Wh")] + p[("lstm", "bh")]
            i, f = torch.sigmoid(z[:, :F]), torch.sigmoid(z[:, F:2 * F])
            g, o = torch.tanh(z[:, 2 * F:3 * F]), torch.sigmoid(z[:, 3 * F:])
            c = f * c + i * g
            return c, o * torch.tanh(c)

        for r in range(hist):  # get_history_carry: teacher forcing, only the LSTM carry survives
            c, h = lstm(history_x[:, r], c, h)
    x = history_x[:, hist]
    xs, us = [x], []
    for _ in range(T):  # teacher_forcing=False: x = xprev
        if F > 0:
            c, h = lstm(x, c, h)
            y = h
        else:
            y = torch.relu(x @ p[("Dense_0", "kernel")] + p[("Dense_0", "bias")])
        x = _head(p, 0, head_layers, y) + x
        us.append(torch.tanh(_head(p, 1, head_layers, y)))
        xs.append(x)
    return torch.stack(xs, 1), torch.stack(us, 1)
