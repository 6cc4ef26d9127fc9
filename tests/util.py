"""Shared helpers for the parity tests."""

import numpy as np
import torch

from gan_mpc_b200 import synthetic

SMALL = dict(n=3, m=1, T=5, dyn_layers=4, dyn_hidden=200, cost_layers=3, cost_hidden=128,
             cost_fout=10)                       # config/l2_hyperparameters.yaml dims (C1)
MID = dict(n=17, m=6, T=8, dyn_layers=4, dyn_hidden=200, cost_layers=3, cost_hidden=128,
           cost_fout=10)                         # C2 dims at a short horizon
ODD = dict(n=5, m=3, T=4, dyn_layers=3, dyn_hidden=50, cost_layers=2, cost_hidden=30,
           cost_fout=7)                          # widths that are not multiples of 4
WIDE = dict(n=17, m=6, T=6, dyn_layers=3, dyn_hidden=512, cost_layers=3, cost_hidden=512,
            cost_fout=10)                        # C4 dims at a short horizon


def to_oracle(p, dtype=torch.float64):
    """numpy param dict -> torch CPU dict for the oracle."""
    return {k: ([torch.from_numpy(w).to(dtype) for w in v] if isinstance(v, list)
                else torch.from_numpy(v).to(dtype)) for k, v in p.items()}


def tt(a, dtype=torch.float64):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dtype)


def make_handle(cfg, params, device=0, critic=None):
    from gan_mpc_b200 import _lib
    kw = {}
    if critic:
        kw = dict(critic_features=critic["F"], critic_layers=critic["L"], critic_hidden=critic["H"])
    h = _lib.Handle(cfg["n"], cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"],
                    cfg["cost_layers"], cfg["cost_hidden"], cfg["cost_fout"], device=device, **kw)
    dev = torch.device("cuda", device)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    h.set_weights([g(w) for w in params["dyn_W"]], [g(b) for b in params["dyn_b"]],
                  [g(w) for w in params["cost_W"]], [g(b) for b in params["cost_b"]],
                  g(params["mpc_weights"]))
    return h


def case(cfg, seed, B, K=1, bias_scale=0.1):
    p = synthetic.planner_params(seed, bias_scale=bias_scale, **cfg)
    x0, U0, goal = synthetic.planner_inputs(seed, B=B, K=K, **cfg)
    return p, x0, U0, goal


def rel_rows(a, b):
    """max over leading rows of ||a-b|| / ||b|| (trajectory-norm-wise relative error)."""
    a = a.double().cpu().reshape(a.shape[0], -1)
    b = b.double().cpu().reshape(b.shape[0], -1)
    return float(((a - b).norm(dim=1) / (b.norm(dim=1) + 1e-30)).max())


def rel_each(a, b):
    a = a.double().cpu().reshape(a.shape[0], -1)
    b = b.double().cpu().reshape(b.shape[0], -1)
    return (a - b).norm(dim=1) / (b.norm(dim=1) + 1e-30)


def assert_rows_close(name, k, o, tol=1e-4, outlier_frac=0.1, cap=5e-2, median_frac=0.25):
    """Row-wise (trajectory-norm-wise) parity after N planning iterations.

    The planner's trajectory through U-space is not continuous in the arithmetic: a hidden
    pre-activation within rounding distance of 0 flips a ReLU mask bit and an Adam step on a
    near-zero gradient flips sign, in ANY fp32 implementation (the oracle's own fp32-vs-fp64
    floor shows the same outliers).  So the bar is: every row within `cap` (nothing is garbage),
    median far below `tol`, and at most `outlier_frac` of the rows (min 2) above `tol`; the
    count is printed, never hidden."""
    e = rel_each(k, o)
    nbad = int((e >= tol).sum())
    print(f"{name}: rows {len(e)}, median {float(e.median()):.2e}, max {float(e.max()):.2e}, rows >= {tol:g}: {nbad}")
    assert float(e.max()) < cap, f"{name}: max row error {float(e.max()):.3e}"
    assert float(e.median()) < tol * median_frac, f"{name}: median row error {float(e.median()):.3e}"
    assert nbad <= max(2, int(outlier_frac * len(e))), f"{name}: {nbad} rows above {tol}"
