"""GPU parity of the dynamics trainer (gmpc_dynamics_fit + the host mirror of
norm/dynamics_trainer.py:13-120) against oracle/dynfit.py (autograd)."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.norm import dynamics_trainer
from gan_mpc_b200.norm import runner as norm_runner
from oracle import dynfit as ofit
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


def _windows(cfg, seed, B, S):
    rng = np.random.Generator(np.random.PCG64(seed))
    n, m = cfg["n"], cfg["m"]
    xs = rng.standard_normal((B, S, n)).astype(np.float32)
    us = np.tanh(rng.standard_normal((B, S, m))).astype(np.float32)
    ys = (xs + 0.1 * rng.standard_normal((B, S, n))).astype(np.float32)
    return xs, us, ys


@pytest.mark.parametrize("tf", [True, False])
@pytest.mark.parametrize("cfg,B,S", [(util.SMALL, 37, 6), (util.MID, 40, 5), (util.ODD, 33, 3), (util.WIDE, 9, 3)])
def test_dynamics_fit_matches_autograd(cfg, B, S, tf, built_lib):
    p, *_ = util.case(cfg, 71, B=1)
    h = util.make_handle(cfg, p)
    op = util.to_oracle(p)
    xs, us, ys = _windows(cfg, 3, B, S)
    g = lambda a: torch.from_numpy(a).cuda()
    dims = [p["dyn_W"][0].shape[0]] + [w.shape[1] for w in p["dyn_W"]]
    loss, act, cot = h.dynamics_fit(g(xs), g(us), g(ys), 0.9, tf, dims)
    ol = ofit.predict_loss(op, util.tt(xs), util.tt(us), util.tt(ys), 0.9, tf)
    assert util.rel_rows(loss[:, None], ol[:, None]) < TOL
    _, oW, ob = ofit.loss_and_grad(op, util.tt(xs), util.tt(us), util.tt(ys), 0.9, tf)
    for l in range(len(act)):
        dW = (act[l] @ cot[l].t() / B).double().cpu()
        db = (cot[l].sum(1) / B).double().cpu()
        eW = float((dW - oW[l]).norm() / oW[l].norm())
        eb = float((db - ob[l]).norm() / ob[l].norm())
        print(f"layer {l}: dW rel err {eW:.2e}, db rel err {eb:.2e}")
        assert eW < TOL and eb < TOL


def test_dynamics_trainer_entry_points(built_lib):
    """predict_loss / train_per_update / train_params with the reference's signatures on the l2 YAML
    config; teacher-forcing switch; only the dynamics leaves move; the loss goes down."""
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    x_size, u_size, D, S = 3, 1, 96, 8
    policy, _, _ = norm_runner.get_policy(config, x_size, u_size)
    params = norm_runner.get_params(policy, config, x_size, u_size)
    xs, us, ys = _windows(dict(n=x_size, m=u_size), 5, D, S)
    X, U, Y = (torch.from_numpy(a).cuda() for a in (xs, us, ys))
    l1 = dynamics_trainer.predict_loss(policy, params, X[0], U[0], Y[0], 0.9, True)
    lb = dynamics_trainer.predict_loss(policy, params, X, U, Y, 0.9, True)
    assert l1.shape == () and lb.shape == (D,) and torch.equal(l1, lb[0])
    dopt, opt_state = norm_runner.get_optimizer(params, config.mpc.train.dynamics.no_grads, lr=1e-3)
    assert dopt.trained == ["dynamics_params"]
    before = utils.tree_clone(params)
    new_params, opt_state, losses = dynamics_trainer.train_params(
        (policy, dopt), opt_state, params, (X, U, Y), num_updates=4, batch_size=32, discount_factor=0.9,
        teacher_forcing_factor=0.5, key=0, id=0)
    assert len(losses) == 4 and opt_state["count"] == 12 and losses[1] < losses[0]
    assert torch.equal(params["dynamics_params"]["params"]["Dense_0"]["kernel"],
                       before["dynamics_params"]["params"]["Dense_0"]["kernel"])        # inputs not mutated
    assert not torch.equal(new_params["dynamics_params"]["params"]["Dense_0"]["kernel"],
                           before["dynamics_params"]["params"]["Dense_0"]["kernel"])
    assert torch.equal(new_params["cost_params"]["params"]["Dense_0"]["kernel"],
                       before["cost_params"]["params"]["Dense_0"]["kernel"])            # masked leaves
    assert torch.equal(new_params["mpc_weights"], before["mpc_weights"])


def test_dynamics_fit_full_size_finite_difference(built_lib):
    """C2 dims, 4096 windows of 8 steps, free running: the loss is the sum of per-step terms, and the
    weight gradient predicts the loss change of a small step along itself (directional finite
    difference through the kernel's own loss) -- no oracle run needed at this size."""
    from gan_mpc_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["C2"], K=1)
    p = synthetic.planner_params(0, bias_scale=0.1, **cfg)
    xs, us, ys = _windows(cfg, 7, 4096, 8)
    g = lambda a: torch.from_numpy(a).cuda()
    dims = [p["dyn_W"][0].shape[0]] + [w.shape[1] for w in p["dyn_W"]]
    h = util.make_handle(cfg, p)
    loss, act, cot = h.dynamics_fit(g(xs), g(us), g(ys), 0.9, False, dims)
    B = loss.shape[0]
    dW = [(a @ c.t()) / B for a, c in zip(act, cot)]
    gn2 = sum(float((w.double() ** 2).sum()) for w in dW)
    eps = 1e-3 / gn2 ** 0.5
    p2 = dict(p, dyn_W=[w + eps * d.cpu().numpy() for w, d in zip(p["dyn_W"], dW)])
    h2 = util.make_handle(cfg, p2)
    loss2, _, _ = h2.dynamics_fit(g(xs), g(us), g(ys), 0.9, False, dims)
    fd = (float(loss2.double().mean()) - float(loss.double().mean())) / eps
    print(f"directional derivative: finite difference {fd:.6e}, gradient norm^2 {gn2:.6e}")
    assert abs(fd - gn2) < 2e-2 * gn2
    tf_loss, _, _ = h.dynamics_fit(g(xs[:, :1].copy()), g(us[:, :1].copy()), g(ys[:, :1].copy()), 0.9, True, dims)
    fr_loss, _, _ = h.dynamics_fit(g(xs[:, :1].copy()), g(us[:, :1].copy()), g(ys[:, :1].copy()), 0.9, False, dims)
    assert torch.equal(tf_loss, fr_loss)        # a window of one step has nothing to force
