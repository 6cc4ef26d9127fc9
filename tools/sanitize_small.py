"""Tiny invocations of every kernel added for the iLQR / bilevel / dynamics-fit / expert rows, meant to
run under compute-sanitizer (memcheck, racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_mpc_b200 import synthetic  # noqa: E402
from gan_mpc_b200.expert import nn as expert_nn  # noqa: E402
from tests import util  # noqa: E402

g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
for cfg, B in ((util.ODD, 35), (util.ODD, 2), (dict(util.SMALL, dyn_hidden=40, cost_hidden=24), 3)):
    p, x0, U0, goal = util.case(cfg, 1, B=B)
    h = util.make_handle(cfg, p, critic=dict(F=8, L=2, H=8))
    h.set_path("ffma")
    out = h.ilqr(g(x0), g(U0[:, 0].copy()), g(goal), maxiter=2, want_lqr=True)
    des = (goal + 0.05).astype(np.float32)
    o = h.bilevel_l2(g(x0), g(U0[:, 0].copy()), g(goal), g(des), maxiter=1, want_hessian=True)
    o2 = h.bilevel_l2(g(x0), o["U"], g(goal), g(des), maxiter=0, V=o["H"].contiguous())
    dl = torch.randn(B, cfg["T"] + 1, cfg["n"], device="cuda")
    o3 = h.bilevel_tail(g(x0), o["U"], g(goal), dl)
    S = 3
    xs = torch.randn(B, S, cfg["n"], device="cuda"); us = torch.randn(B, S, cfg["m"], device="cuda")
    dims = [p["dyn_W"][0].shape[0]] + [w.shape[1] for w in p["dyn_W"]]
    for tf in (True, False):
        h.dynamics_fit(xs, us, xs + 0.1, 0.9, tf, dims)
    for md in (expert_nn.ScanLSTM(8, 2, 12, cfg["n"], cfg["m"]), expert_nn.ScanMLP(3, 12, cfg["n"], cfg["m"])):
        flat = torch.from_numpy(synthetic.expert_params_flat(0, md._shapes(), md.lstm_features)).cuda()
        h.expert_propose(torch.randn(B, 3, cfg["n"], device="cuda"), flat, md.lstm_features, md.num_layers,
                         md.num_hidden_units)
    cf = torch.from_numpy(synthetic.critic_params_flat(0, cfg["n"], 8, 2, 8)).cuda()
    h.critic_input_grad(torch.randn(B, cfg["T"] + 1, cfg["n"], device="cuda"), cf)
    torch.cuda.synchronize()
    print("ok", cfg["n"], cfg["m"], B, float(out[2].sum()), float(o["H"].abs().sum()))
