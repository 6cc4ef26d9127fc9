"""GPU parity of the expert proposal network (gmpc_expert_propose; reference expert/nn.py:22-131,
expert/expert_model.py:60-91, policy/eval.py:87-107) against oracle/expert.py, and the acting path
expert -> planner through the reference-named policy methods."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import synthetic, utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.expert import nn as expert_nn
from gan_mpc_b200.norm import runner as norm_runner
from oracle import expert as oexpert
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


@pytest.mark.parametrize("kind,F,L,H", [("lstm", 128, 3, 128), ("lstm", 24, 1, 16), ("lstm", 40, 2, 72),
                                        ("mlp", 0, 3, 128), ("mlp", 0, 2, 20)])
@pytest.mark.parametrize("cfg,hist,B", [(util.SMALL, 1, 37), (util.MID, 3, 9), (util.ODD, 0, 5)])
def test_expert_propose_matches_oracle(kind, F, L, H, cfg, hist, B, built_lib):
    n, m, T = cfg["n"], cfg["m"], cfg["T"]
    md = (expert_nn.ScanLSTM(F, L, H, n, m) if kind == "lstm" else expert_nn.ScanMLP(L, H, n, m))
    flat = synthetic.expert_params_flat(5, md._shapes(), md.lstm_features)
    rng = np.random.Generator(np.random.PCG64(9))
    flat = flat + (0.05 * rng.standard_normal(flat.shape)).astype(np.float32)   # non-zero biases
    hx = rng.standard_normal((B, hist + 1, n)).astype(np.float32)
    p, *_ = util.case(cfg, 1, B=1)
    h = util.make_handle(cfg, p)
    goal, useq = h.expert_propose(torch.from_numpy(hx).cuda(), torch.from_numpy(flat).cuda(), F, L, H)
    og, ou = oexpert.propose(util.tt(hx), util.tt(flat), md._shapes(), F, md.head_layers, T)
    assert goal.shape == (B, T + 1, n) and useq.shape == (B, T, m)
    assert torch.equal(goal[:, 0].cpu(), torch.from_numpy(hx[:, -1]))
    assert util.rel_rows(goal, og) < TOL and util.rel_rows(useq, ou) < TOL
    # pytree <-> flat round trip in the flax naming
    tree = md.unflatten(torch.from_numpy(flat))
    assert torch.equal(md.flatten(tree), torch.from_numpy(flat))
    with pytest.raises(ValueError):
        h.expert_propose(torch.from_numpy(hx).cuda(), torch.from_numpy(flat[:-1].copy()).cuda(), F, L, H)


def test_acting_with_the_expert_network(built_lib):
    """get_policy(..., expert_model="network"): EvalMPC.get_optimal_action = expert proposal ->
    plan, unbatched and batched, with the reference's method names."""
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    x_size, u_size = 3, 1
    train_policy, eval_policy, _ = norm_runner.get_policy(config, x_size, u_size, expert_model="network")
    with pytest.raises(FileNotFoundError):
        norm_runner.get_params(train_policy, config, x_size, u_size)       # no checkpoint is shipped
    params = norm_runner.get_params(train_policy, config, x_size, u_size, load_expert=False)
    cell = params["expert_params"]["params"]["model"]["ScanLSTMCell_0"]
    assert cell["OptimizedLSTMCell_0"]["hi"]["kernel"].shape == (128, 128)
    assert cell["MLPCell_1"]["Dense_2"]["kernel"].shape == (128, u_size)
    T = config.mpc.horizon
    hx = torch.randn(6, 2, x_size, generator=torch.Generator().manual_seed(2)).cuda()
    goal, init_u = eval_policy.get_goal_states_init_actions(hx, params)
    assert goal.shape == (6, T + 1, x_size) and init_u.shape == (6, T, u_size)
    assert float(init_u.abs().max()) <= 1.0 and torch.equal(goal[:, 0], hx[:, -1])
    g1, u1 = eval_policy.get_goal_states_init_actions(hx[2], params)
    assert torch.equal(g1, goal[2]) and torch.equal(u1, init_u[2])
    # the reference's two-call sequence gives the same proposal
    em = eval_policy.expert_model
    carry = em.get_history_carry(hx[2], None, params["expert_params"])
    _, (g2, u2) = em.get_carry_next_state_and_action_seq(carry, torch.zeros(T, x_size), params["expert_params"])
    assert torch.equal(g2, g1) and torch.equal(u2, u1)
    u0 = eval_policy.get_optimal_action(params, hx, torch.zeros(6, 1, u_size).cuda())
    assert u0.shape == (6, u_size) and bool(torch.isfinite(u0).all())
    X, U, *_ = train_policy.get_optimal_values(params, hx)
    assert torch.equal(U[:, 0], u0)
