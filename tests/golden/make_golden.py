"""Generate the frozen golden vectors under tests/golden/ from the fp64 oracle.

The reference itself cannot run here (no jax/flax/optax/trajax, no network), so these vectors
are produced by the CPU restatement in oracle/ -- PARITY UNPINNED against the reference; they pin
the restatement (and through it the kernels) against drift.  Run from the repo root:
    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from gan_mpc_b200 import synthetic  # noqa: E402
from oracle import critic as ocritic  # noqa: E402
from oracle import ilqr as oilqr  # noqa: E402
from oracle import planner as oracle  # noqa: E402
from tests import util  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def planner_case(name, cfg, seed, B, K, iters, lr):
    p, x0, U0, goal = util.case(cfg, seed, B=B, K=K, bias_scale=0.1)
    op = util.to_oracle(p)
    tx0, tU0, tgoal = util.tt(x0), util.tt(U0), util.tt(goal)
    X, J, dU, lam = oracle.objective_grad(tx0, tU0[:, 0], tgoal, op)
    out = dict(cfg, iters=iters, lr=lr, x0=x0, U0=U0, goal=goal, mpc_weights=p["mpc_weights"],
               X=X.numpy(), J=J.numpy(), dU=dU.numpy(), lam=lam.numpy())
    for i, (w, b) in enumerate(zip(p["dyn_W"], p["dyn_b"])):
        out[f"dyn_W{i}"], out[f"dyn_b{i}"] = w, b
    for i, (w, b) in enumerate(zip(p["cost_W"], p["cost_b"])):
        out[f"cost_W{i}"], out[f"cost_b{i}"] = w, b
    for method in ("grad", "adam"):
        Ub, Xb, Jb, idx, Jall = oracle.plan(tx0, tU0, tgoal, op, method, iters, lr)
        out.update({f"{method}_U_best": Ub.numpy(), f"{method}_X_best": Xb.numpy(),
                    f"{method}_J_best": Jb.numpy(), f"{method}_idx": idx.numpy(),
                    f"{method}_J_all": Jall.numpy()})
    np.savez_compressed(os.path.join(HERE, f"planner_{name}.npz"), **out)


def critic_case():
    n, F, L, H, T1, D = 3, 16, 2, 8, 6, 6
    flat = synthetic.critic_params_flat(0, n, F, L, H)
    rng = np.random.Generator(np.random.PCG64(7))
    flat = flat + (0.05 * rng.standard_normal(flat.shape)).astype(np.float32)  # non-zero biases
    xs, lab = synthetic.critic_dataset(0, D, T1, n)
    tf, tx, tl = util.tt(flat), util.tt(xs), util.tt(lab)
    logit = ocritic.critic_logit(tx, tf, n, F, L, H)
    loss, g = ocritic.critic_loss_and_grad(tx, tl, tf, n, F, L, H)
    np.savez_compressed(os.path.join(HERE, "critic_small.npz"), n=n, F=F, L=L, H=H, flat=flat,
                        xseq=xs, label=lab, logit=logit.numpy(), loss=loss.numpy(), grad=g.numpy())


def ilqr_case(only_missing=False):
    """trajax iLQR restatement (oracle/ilqr.py) on the C1 dims: 6 trajectories, 8 iterations."""
    path = os.path.join(HERE, "ilqr_small.npz")
    if only_missing and os.path.exists(path):
        return
    seed, B, maxiter = 31, 6, 8
    p, x0, U0, goal = util.case(util.SMALL, seed, B=B)
    X, U, obj, g, lam, _, it = oilqr.ilqr(util.tt(x0), util.tt(U0[:, 0]), util.tt(goal),
                                          util.to_oracle(p), maxiter=maxiter)
    np.savez_compressed(path, seed=seed, B=B, maxiter=maxiter, X=X.numpy(), U=U.numpy(),
                        obj=obj.numpy(), gradient=g.numpy(), adjoints=lam.numpy(),
                        iteration=it.numpy())


if __name__ == "__main__":
    torch.set_num_threads(1)
    if "--ilqr-only" in sys.argv:
        ilqr_case()
        sys.exit(0)
    ilqr_case()
    planner_case("small", util.SMALL, seed=11, B=5, K=3, iters=6, lr=1e-2)
    planner_case("mid", util.MID, seed=12, B=4, K=2, iters=4, lr=1e-2)
    critic_case()
    print("golden vectors written to", HERE)
