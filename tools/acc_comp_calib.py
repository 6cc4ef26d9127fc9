"""Calibration of the accumulator un-biasing of plan_t128 (GMPC_T128_ACC_COMP, in units of 2^-24 per accumulate
event): rollout / plan errors against the fp64 oracle at C2 dims for the value in the environment.
usage: for c in 0 0.25 0.5 0.75 1; do GMPC_T128_ACC_COMP=$c python tools/acc_comp_calib.py; done"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_mpc_b200 import synthetic  # noqa: E402
from oracle import planner as oracle  # noqa: E402
from tests import util  # noqa: E402

cfg = dict(synthetic.CONFIGS["C2"], B=int(os.environ.get("CALIB_B", 1024)))
p = synthetic.planner_params(0, **cfg)
x0, U0, goal = synthetic.planner_inputs(0, **cfg)
h = util.make_handle({k: cfg[k] for k in ("n", "m", "T", "dyn_layers", "dyn_hidden", "cost_layers", "cost_hidden", "cost_fout")}, p)
h.set_path(os.environ.get("CALIB_PATH", "t128"))
dev = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
op = util.to_oracle(p)
oX, oJ, odU, olam = oracle.objective_grad(util.tt(x0), util.tt(U0[:, 0]), util.tt(goal), op, [None])
J, dU, X, lam = h.objective_grad(dev(x0), dev(U0[:, 0]), dev(goal), want_lam=True)
med = lambda a, b: float(util.rel_each(a, b).median())
# signed bias of the rollout: mean of (|X| - |oX|) / |oX| over the last state's large components
xe, xo = X[:, -1].double().cpu(), oX[:, -1]
big = xo.abs() > xo.abs().max(dim=1, keepdim=True).values * 0.1
bias = float((((xe.abs() - xo.abs()) / xo.abs())[big]).mean())
o64 = oracle.plan(util.tt(x0), util.tt(U0), util.tt(goal), op, "adam", 20, 1e-2)
Ub, Xb, Jb, idx, _ = h.plan(dev(x0), dev(U0), dev(goal), method="adam", iters=20, lr=1e-2)
eu = util.rel_each(Ub, o64[0])
print(f"acc_comp {os.environ.get('GMPC_T128_ACC_COMP', 'default')} path {h.last_path}: rollout X median {med(X, oX):.2e} signed bias of x_T {bias:+.2e} "
      f"| J {med(J[:, None], oJ[:, None]):.2e} dU {med(dU, odU):.2e} | plan: U median {float(eu.median()):.2e} rows>=1e-4 {int((eu >= 1e-4).sum())}/{len(eu)} "
      f"X {med(Xb, o64[1]):.2e} J {med(Jb[:, None], o64[2][:, None]):.2e}")
