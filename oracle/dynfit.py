"""CPU restatement of the dynamics trainer's loss (norm/dynamics_trainer.py:13-44) and, through
torch autograd standing in for jax.value_and_grad (:64-79), its weight gradient.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED."""

import torch

from oracle import planner as pl


def predict_loss(params, xseq, useq, next_xseq, discount_factor, teacher_forcing):
    """batched windows xseq [B,S,n], useq [B,S,m], next_xseq [B,S,n] -> loss [B]."""
    S = xseq.shape[1]
    xprev = xseq[:, 0]
    loss = torch.zeros(xseq.shape[0], dtype=xseq.dtype)
    disc = 1.0
    for t in range(S):
        x = xseq[:, t] if teacher_forcing else xprev
        xprev = pl.dynamics_mlp(x, useq[:, t], params["dyn_W"], params["dyn_b"])
        loss = loss + disc * ((xprev - next_xseq[:, t]) ** 2).sum(-1)   # utils.discounted_sum, then sum
        disc = disc * discount_factor
    return loss


def loss_and_grad(params, xseq, useq, next_xseq, discount_factor, teacher_forcing):
    """mean over the batch and its gradient w.r.t. the dynamics kernels and biases."""
    Ws = [w.detach().clone().requires_grad_(True) for w in params["dyn_W"]]
    bs = [b.detach().clone().requires_grad_(True) for b in params["dyn_b"]]
    p2 = dict(params, dyn_W=Ws, dyn_b=bs)
    loss = predict_loss(p2, xseq, useq, next_xseq, discount_factor, teacher_forcing).mean()
    gs = torch.autograd.grad(loss, Ws + bs)
    return loss.detach(), list(gs[:len(Ws)]), list(gs[len(Ws):])
