"""GPU tests of the drop-in Python API (EvalMPC / L2MPC / JS_MPC / policy.optimizers /
critic_trainer) against the oracle -- written the way a test of the reference would read."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import optim, utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.gan import critic_trainer
from gan_mpc_b200.gan import runner as gan_runner
from gan_mpc_b200.norm import cost_trainer
from gan_mpc_b200.norm import runner as norm_runner
from gan_mpc_b200.policy import optimizers as opt
from oracle import critic as ocritic
from oracle import planner as oracle
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


def cfg(name):
    return utils.get_config(os.path.join(load_config.CONFIG_DIR, name))


def oracle_params(params, dtype=torch.float64):
    def lists(tree):
        p = tree["params"]
        return ([p[f"Dense_{i}"]["kernel"].cpu().to(dtype) for i in range(len(p))],
                [p[f"Dense_{i}"]["bias"].cpu().to(dtype) for i in range(len(p))])
    dW, db = lists(params["dynamics_params"])
    cW, cb = lists(params["cost_params"])
    return dict(dyn_W=dW, dyn_b=db, cost_W=cW, cost_b=cb, mpc_weights=params["mpc_weights"].cpu().to(dtype))


def test_l2_policy_plan_and_act(built_lib):
    """config/l2_hyperparameters.yaml defaults (BASELINE config 1): n=3, m=1, T=5, single state."""
    config = cfg("l2_hyperparameters.yaml")
    x_size, u_size = 3, 1
    train_policy, eval_policy, _ = norm_runner.get_policy(config, x_size, u_size)
    # the YAML default is the reference's planner (trajax iLQR, test_policy_with_trajax_ilqr_method and
    # test_default_planner_is_the_reference_ilqr); this test exercises the north star's first-order planner
    assert eval_policy.planner_kwargs["method"] == "ilqr"
    for pol in (train_policy, eval_policy):
        pol.planner_kwargs["method"] = "adam"
    params = norm_runner.get_params(train_policy, config, x_size, u_size)
    assert set(params) == {"mpc_weights", "cost_params", "dynamics_params", "expert_params"}
    assert params["dynamics_params"]["params"]["Dense_0"]["kernel"].shape == (4, 200)
    history_x = torch.randn(2, x_size, generator=torch.Generator().manual_seed(0)).cuda()
    history_u = torch.zeros(1, u_size).cuda()
    with pytest.warns(UserWarning, match="first-order planner"):
        X, U, obj, gradient, adjoints, lqr, iteration = eval_policy.get_optimal_values(params, history_x, history_u)
    T = config.mpc.horizon
    assert X.shape == (T + 1, x_size) and U.shape == (T, u_size) and obj.shape == ()
    assert gradient.shape == (T, u_size) and adjoints.shape == (T + 1, x_size) and lqr is None
    assert int(iteration) == config.mpc.planner.iters
    u0 = eval_policy.get_optimal_action(params, history_x, history_u)
    assert torch.equal(u0, U[0])
    # same plan from the oracle, fed the expert's proposals
    goal, init_u = eval_policy.get_goal_states_init_actions(history_x, params)
    pk = eval_policy.planner_kwargs
    oU, oX, oJ, _, _ = oracle.plan(history_x[-1].cpu().double()[None], init_u.cpu().double()[None, None],
                                   goal.cpu().double()[None], oracle_params(params), pk["method"],
                                   pk["iters"], pk["learning_rate"])
    assert util.rel_rows(U[None], oU) < TOL and util.rel_rows(X[None], oX) < TOL
    assert abs(float(obj) - float(oJ)) < TOL * abs(float(oJ))
    # train policy (BaseMPC signature: no history_u) gives the same plan
    X2, U2, *_ = train_policy.get_optimal_values(params, history_x)
    assert torch.equal(U2, U) and torch.equal(X2, X)


def test_default_planner_is_the_reference_ilqr(built_lib):
    """Defaults: the policy plans with trajax iLQR like the reference (and like BaseMPC's bilevel gradient
    assumes); two different same-shape parameter pytrees staged back to back are never confused."""
    config = cfg("l2_hyperparameters.yaml")
    policy, eval_policy, _ = norm_runner.get_policy(config, 3, 1)
    assert policy.planner_kwargs["method"] == "ilqr" and eval_policy.planner_kwargs["method"] == "ilqr"
    params = norm_runner.get_params(policy, config, 3, 1)
    hx = torch.randn(4, 2, 3, generator=torch.Generator().manual_seed(5)).cuda()
    X, U, obj, grad, lam, lqr, it = policy.get_optimal_values(params, hx)
    assert policy.last_plan_info["path"] == "ilqr" and it.dtype == torch.int32 and int(it.max()) <= 100
    # a functionally created replacement pytree (every leaf a NEW tensor, version 0, quite possibly at the
    # addresses of the old leaves once those are freed) must be re-staged
    import copy

    def scaled(tree, f):
        return {k: scaled(v, f) for k, v in tree.items()} if isinstance(tree, dict) else tree * f
    params2 = dict(params)
    params2["cost_params"] = scaled(params["cost_params"], 1.5)
    obj2 = policy.get_optimal_values(params2, hx)[2].clone()
    del params
    params3 = dict(params2)
    params3["cost_params"] = scaled(params2["cost_params"], 1.0 / 1.5)
    obj3 = policy.get_optimal_values(params3, hx)[2]
    assert not torch.allclose(obj2, obj) and torch.allclose(obj3, obj, rtol=1e-4)


def test_model_entry_points_are_callable(built_lib):
    """DynamicsModel.predict (dynamics/dynamics_model.py:45-48), MujocoBasedModel.get_cost
    (cost/cost_model.py:33-42: staging cost, terminal cost at t == horizon) and CriticModel.predict
    (critic/critic_model.py:15-16) evaluate one step / one trajectory through the kernels, vs the oracle."""
    from oracle import critic as ocritic
    config = cfg("gan_hyperparameters.yaml")
    n, m, B = 3, 1, 9
    policy, _, _ = gan_runner.get_policy(config, n, m)
    params = gan_runner.get_params(policy, config, n, m)
    T = config.mpc.horizon
    gen = torch.Generator().manual_seed(11)
    x = torch.randn(B, n, generator=gen).cuda()
    u = torch.tanh(torch.randn(B, m, generator=gen)).cuda()
    goal = torch.randn(B, T + 1, n, generator=gen).cuda()
    op = oracle_params(params)
    d = lambda t_: t_.cpu().double()
    # one dynamics step, batched and unbatched
    xn = policy.dynamics_model.predict(x, u, 0, params["dynamics_params"])
    oxn = oracle.dynamics_mlp(d(x), d(u), op["dyn_W"], op["dyn_b"])
    assert util.rel_rows(xn, oxn) < TOL
    assert torch.equal(policy.dynamics(x[2], u[2], 0, params), xn[2])
    # one step cost: staging (t < horizon) and terminal (t == horizon), through policy.cost as the reference calls it
    for t in (0, 2, T):
        c = policy.cost(x, u, t, params, goal)
        oc = torch.stack([oracle.step_cost(d(x[b]), d(u[b]), t, T, op, d(goal[b])) for b in range(B)])
        assert util.rel_rows(c[:, None], oc[:, None]) < TOL, t
    assert abs(float(policy.cost(x[1], u[1], T, params, goal[1])) - float(oc[1])) < TOL * abs(float(oc[1]))
    # critic score
    cm = policy.critic_model
    xs = torch.randn(B, T + 1, n, generator=gen).cuda()
    s = cm.predict(xs, params["critic_params"])
    c_ = cm.model
    os_ = ocritic.critic_logit(d(xs), d(c_.flatten(params["critic_params"])), n, c_.lstm_features, c_.num_layers,
                               c_.num_hidden_units)
    assert s.shape == (B, 1) and util.rel_rows(s, os_.reshape(B, 1)) < TOL
    assert cm.predict(xs[3], params["critic_params"]).shape == (1,)


def test_batched_policy_and_optimizer_functions(built_lib):
    config = cfg("l2_hyperparameters.yaml")
    x_size, u_size, B = 3, 1, 70
    policy, _, _ = norm_runner.get_policy(config, x_size, u_size)
    policy.planner_kwargs["method"] = "adam"      # first-order planner (the default is "ilqr")
    params = norm_runner.get_params(policy, config, x_size, u_size)
    hx = torch.randn(B, 2, x_size, generator=torch.Generator().manual_seed(1)).cuda()
    X, U, obj, grad, lam, _, it = policy.get_optimal_values(params, hx)
    T = config.mpc.horizon
    assert X.shape == (B, T + 1, x_size) and U.shape == (B, T, u_size) and obj.shape == (B,)
    goal, init_u = policy.get_goal_states_init_actions(hx, params)
    op = oracle_params(params)
    x0 = hx[:, -1]
    # objective / rollout / gradient through the reference-named functions
    cost = opt.bind(policy.cost, params, (goal,))
    dyn = opt.bind(policy.dynamics, params)
    J = opt.objective(cost, dyn, init_u, x0)
    oX, oJ, odU, _ = oracle.objective_grad(x0.cpu().double(), init_u.cpu().double(), goal.cpu().double(), op)
    assert util.rel_rows(J[:, None], oJ[:, None]) < TOL
    assert util.rel_rows(opt.rollout(dyn, init_u, x0), oX) < TOL
    J2, dU, X2, _ = opt.objective_and_grad(cost, dyn, init_u, x0)
    util.assert_rows_close("dJ/dU", dU, odU, TOL)   # a ReLU-kink-adjacent row may exceed 1e-4 (see util)
    # ilqr_solve with the reference's argument order
    out = opt.ilqr_solve(policy.cost, policy.dynamics, x0, init_u, params, (goal,), (), policy.trajax_ilqr_kwargs)
    assert torch.equal(out[1], U)
    # L2 loss and its control gradient
    desired = goal
    loss = policy.loss(X, U, params, desired)
    assert util.rel_rows(loss[:, None], oracle.l2_loss(X.cpu().double(), desired.cpu().double())[:, None]) < TOL
    g = opt.loss_grad_wrt_control(policy.loss, dyn, x0, U, (params, desired))
    og = oracle.loss_grad_wrt_control_l2(x0.cpu().double(), U.cpu().double(), desired.cpu().double(), op)
    util.assert_rows_close("dL2/dU", g, og, TOL)
    # cost_trainer.calculate_loss = mean loss of the plans
    test_loss = cost_trainer.calculate_loss(policy, params, (hx, desired))
    assert abs(float(test_loss) - float(loss.mean())) < 1e-5 * abs(float(loss.mean()))
    # loss_and_grad (the bilevel gradient) returns a params-shaped pytree; parity in test_gpu_bilevel.py
    bl_loss, bl_grads = policy.loss_and_grad(hx, params, (desired,))
    assert bl_loss.shape == () and set(bl_grads) == set(params)


def test_restaging_follows_parameter_updates(built_lib):
    config = cfg("l2_hyperparameters.yaml")
    policy, _, _ = norm_runner.get_policy(config, 3, 1)
    params = norm_runner.get_params(policy, config, 3, 1)
    hx = torch.randn(5, 2, 3, generator=torch.Generator().manual_seed(2)).cuda()
    J1 = policy.get_optimal_values(params, hx)[2].clone()
    params["cost_params"]["params"]["Dense_2"]["kernel"].mul_(2.0)        # in-place update
    J2 = policy.get_optimal_values(params, hx)[2].clone()
    assert not torch.allclose(J1, J2)
    params["mpc_weights"] = params["mpc_weights"] + 1.0                   # replaced tensor
    J3 = policy.get_optimal_values(params, hx)[2]
    assert not torch.allclose(J2, J3)


def test_js_policy_losses_and_critic_training(built_lib):
    """gan_hyperparameters.yaml (BASELINE config 3 at a small dataset): critic BCE loss/grad,
    generator loss, get_dataset (planner on every sample) and the minibatch scan."""
    config = cfg("gan_hyperparameters.yaml")
    x_size, u_size, D = 3, 1, 96
    policy, _, _ = gan_runner.get_policy(config, x_size, u_size)
    params = gan_runner.get_params(policy, config, x_size, u_size)
    assert "critic_params" in params
    cm = policy.critic_model.model
    F, L, H, T = cm.lstm_features, cm.num_layers, cm.num_hidden_units, config.mpc.horizon
    gen = torch.Generator().manual_seed(3)
    X = torch.randn(D, 2, x_size, generator=gen).cuda()
    true_Y = (X[:, -1:, :] + 0.1 * torch.cumsum(torch.randn(D, T + 1, x_size, generator=gen).cuda(), 1)).contiguous()
    split = 64
    true_dataset = ((X[:split], true_Y[:split]), (X[split:], true_Y[split:]))
    flat = policy.critic_flat(params)
    # loss / grad pytree
    lab = torch.cat([torch.ones(8), -torch.ones(8)]).cuda()
    loss, grads = policy.critic_loss_and_grad(true_Y[:16], lab, params)
    oloss, ograd = ocritic.critic_loss_and_grad(true_Y[:16].cpu().double(), lab.cpu().double(),
                                                flat.cpu().double(), x_size, F, L, H)
    assert abs(float(loss) - float(oloss)) < 1e-5
    assert float((cm.flatten(grads["critic_params"]).cpu().double() - ograd).norm() / ograd.norm()) < TOL
    assert float(grads["cost_params"]["params"]["Dense_0"]["kernel"].abs().sum()) == 0.0
    assert abs(float(policy.critic_loss(true_Y[:16], lab, params)) - float(loss)) == 0.0
    gl = policy.generator_loss(true_Y[:4], None, params, true_Y[:4])
    assert torch.allclose(gl.cpu().double(), ocritic.generator_loss(true_Y[:4].cpu().double(), flat.cpu().double(),
                                                                     x_size, F, L, H), atol=1e-5)
    # dataset: expert (+1) and planned (-1) trajectories, train part permuted
    (trX, trY), (teX, teY) = critic_trainer.get_dataset(policy, params, true_dataset, key=0)
    assert trX.shape == (2 * split, T + 1, x_size) and teX.shape == (2 * (D - split), T + 1, x_size)
    assert float(trY.sum()) == 0.0 and set(trY.tolist()) == {1.0, -1.0}
    planned = policy.get_optimal_values(params, X[split:])[0]
    assert torch.equal(teX[D - split:], planned) and torch.equal(teX[:D - split], true_Y[split:])
    # the scan: reference signature, deterministic key
    copt, copt_state = gan_runner.get_optimizer(params, config.mpc.train.critic.no_grads, lr=1e-3)
    assert copt.trained == ["critic_params"]
    perm = torch.randint(0, trX.shape[0], (3, 32), generator=torch.Generator().manual_seed(5)).to(torch.int32).cuda()
    new_params, copt_state, mean_loss = critic_trainer.train_critic_parameters(
        (policy, copt), copt_state, params, perm, (trX, trY))
    of, _, _, _, ol = ocritic.train_critic_parameters(
        flat.cpu().double(), torch.zeros(flat.numel(), dtype=torch.float64), torch.zeros(flat.numel(), dtype=torch.float64),
        0, perm.cpu().long(), trX.cpu().double(), trY.cpu().double(), 1e-3, x_size, F, L, H)
    assert float((policy.critic_flat(new_params).cpu().double() - of).abs().max()) < 2e-5
    assert abs(float(mean_loss) - float(ol)) < 1e-5 and copt_state["count"] == 3
    assert torch.equal(policy.critic_flat(params), flat)                     # inputs are not mutated
    # full train(): (params, opt_state, train_losses, test_losses, minutes)
    copt, copt_state = gan_runner.get_optimizer(params, config.mpc.train.critic.no_grads, lr=1e-3)
    out = critic_trainer.train((policy, copt), copt_state, params, true_dataset, num_updates=2,
                               batch_size=32, key=0, id=1)
    assert len(out) == 5 and len(out[2]) == 2 and len(out[3]) == 2 and out[4] >= 0.0
    assert out[2][1] < out[2][0]                                            # the critic learns


def test_policy_with_trajax_ilqr_method(built_lib):
    """planner method "ilqr" = the reference's own step (trajax iLQR under TRAJAX_iLQR_KWARGS,
    policy/eval.py:10-20): same entry points, 7-tuple with per-state iteration counts."""
    from oracle import ilqr as oilqr
    config = cfg("l2_hyperparameters.yaml")
    x_size, u_size, B = 3, 1, 40
    policy, eval_policy, _ = norm_runner.get_policy(config, x_size, u_size)
    params = norm_runner.get_params(policy, config, x_size, u_size)
    policy.planner_kwargs["method"] = "ilqr"
    hx = torch.randn(B, 2, x_size, generator=torch.Generator().manual_seed(4)).cuda()
    goal, init_u = policy.get_goal_states_init_actions(hx, params)
    kw = dict(policy.trajax_ilqr_kwargs, maxiter=2)
    X, U, obj, grad, lam, lqr, it = opt.ilqr_solve(policy.cost, policy.dynamics, hx[:, -1], init_u, params,
                                                   (goal,), (), kw)
    assert lqr is None and it.dtype == torch.int32 and int(it.max()) <= 2
    o = oilqr.ilqr(hx[:, -1].cpu().double(), init_u.cpu().double(), goal.cpu().double(), oracle_params(params),
                   **kw)
    util.assert_rows_close("ilqr U", U, o[1], TOL, outlier_frac=0.1, cap=2.0)
    util.assert_rows_close("ilqr obj", obj[:, None], o[2][:, None], TOL, outlier_frac=0.1, cap=1.0)
    assert int((it.cpu() == o[6]).sum()) >= B - 2
    # unbatched call through the policy (full 100-iteration default options): shapes of the reference
    Xs, Us, objs, gs, lams, _, its = policy.get_optimal_values(params, hx[0])
    T = config.mpc.horizon
    assert Xs.shape == (T + 1, x_size) and Us.shape == (T, u_size) and objs.shape == () and its.shape == ()
    J0 = opt.objective(opt.bind(policy.cost, params, (goal[0],)), opt.bind(policy.dynamics, params), init_u[0], hx[0, -1])
    assert float(objs) <= float(J0) * (1 + 1e-6)
