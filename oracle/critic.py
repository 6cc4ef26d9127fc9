"""CPU restatement of the JS-GAN critic path (torch).  TEST INFRASTRUCTURE ONLY.
PARITY UNPINNED -- see oracle/__init__.py.

Flat critic parameter layout shared with include/gmpc.h (fp32, row-major):
    Wi [n, 4F]   input kernels  ii|if|ig|io concatenated along columns (no bias)
    Wh [F, 4F]   recurrent kernels hi|hf|hg|ho
    bh [4F]      recurrent biases
    for each hidden Dense l < num_layers-1:  Dk_l [in, H], Db_l [H]
    Wo [in_last, 1], bo [1]
"""

import torch


def critic_param_count(n, F, num_layers, H):
    cnt = n * 4 * F + F * 4 * F + 4 * F
    d = F
    for _ in range(num_layers - 1):
        cnt += d * H + H
        d = H
    return cnt + d + 1


def unflatten(flat, n, F, num_layers, H):
    o = 0

    def take(*shape):
        nonlocal o
        cnt = 1
        for s in shape:
            cnt *= s
        t = flat[o:o + cnt].reshape(*shape)
        o += cnt
        return t

    p = {"Wi": take(n, 4 * F), "Wh": take(F, 4 * F), "bh": take(4 * F), "D": []}
    d = F
    for _ in range(num_layers - 1):
        p["D"].append((take(d, H), take(H)))
        d = H
    p["Wo"] = take(d, 1)
    p["bo"] = take(1)
    assert o == flat.numel()
    return p


def critic_logit(xseq, flat, n, F, num_layers, H):
    """critic/nn.py:27-42 -- scan OptimizedLSTMCell over the T+1 rows from zero (c,h),
    final h -> (num_layers-1) x relu(Dense(H)) -> Dense(1).  xseq [..., T+1, n] -> [...]."""
    p = unflatten(flat, n, F, num_layers, H)
    lead = xseq.shape[:-2]
    c = torch.zeros(*lead, F, dtype=xseq.dtype)
    h = torch.zeros(*lead, F, dtype=xseq.dtype)
    for t in range(xseq.shape[-2]):
        z = xseq[..., t, :] @ p["Wi"] + h @ p["Wh"] + p["bh"]
        i = torch.sigmoid(z[..., 0 * F:1 * F])
        f = torch.sigmoid(z[..., 1 * F:2 * F])
        g = torch.tanh(z[..., 2 * F:3 * F])
        o = torch.sigmoid(z[..., 3 * F:4 * F])
        c = f * c + i * g
        h = o * torch.tanh(c)
    out = h
    for W, b in p["D"]:
        out = torch.relu(out @ W + b)
    return (out @ p["Wo"] + p["bo"])[..., 0]


def critic_loss(xseq, label, flat, n, F, num_layers, H):
    """gan/js_policy.py:41-46 -- p = sigmoid(s); p = where(label>0, p, 1-p); -log p.
    Stable form: -log sigmoid(+-s) = softplus(-+s)."""
    s = critic_logit(xseq, flat, n, F, num_layers, H)
    return torch.nn.functional.softplus(torch.where(label > 0, -s, s))


def critic_loss_and_grad(batch_xseq, batch_label, flat, n, F, num_layers, H):
    """gan/js_policy.py:48-58 -- value_and_grad of the batch-mean loss w.r.t. critic params."""
    flat = flat.detach().clone().requires_grad_(True)
    loss = critic_loss(batch_xseq, batch_label, flat, n, F, num_layers, H).mean()
    (g,) = torch.autograd.grad(loss, flat)
    return loss.detach(), g


def generator_loss(xseq, flat, n, F, num_layers, H):
    """gan/js_policy.py:60-68 -- mean(-log p + log(1-p)) with p = sigmoid(s)  (== -s)."""
    return -critic_logit(xseq, flat, n, F, num_layers, H)


def clip_adam_step(params, grad, mom, vel, step, lr, max_norm=100.0,
                   b1=0.9, b2=0.999, eps=1e-8):
    """norm/runner.py:53-56 -- optax.chain(clip_by_global_norm(max_norm), adam(lr)) then apply_updates.
    `step` is the 1-based count after this update.  Returns (params, mom, vel)."""
    gn = torch.sqrt((grad * grad).sum())
    g = torch.where(gn < max_norm, grad, grad / gn * max_norm)
    mom = b1 * mom + (1.0 - b1) * g
    vel = b2 * vel + (1.0 - b2) * g * g
    mhat = mom / (1.0 - b1 ** step)
    vhat = vel / (1.0 - b2 ** step)
    return params - lr * mhat / (torch.sqrt(vhat) + eps), mom, vel


def train_critic_parameters(flat, mom, vel, step0, perm, X, Y, lr, n, F, num_layers, H):
    """gan/critic_trainer.py:48-65 -- sequential scan over minibatch index rows `perm`."""
    losses = []
    step = step0
    for p in perm:
        loss, g = critic_loss_and_grad(X[p], Y[p], flat, n, F, num_layers, H)
        step += 1
        flat, mom, vel = clip_adam_step(flat, g, mom, vel, step, lr)
        losses.append(loss)
    return flat, mom, vel, step, torch.stack(losses).mean()
