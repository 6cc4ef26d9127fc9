"""GPU parity of the bilevel (cost-training) gradient -- gmpc_bilevel_l2 + the host mirror of
policy/optimizers.py:34-105 and policy/base.py:87-128 -- against oracle/bilevel.py, which restates
those lines by literal autodiff (torch.autograd standing in for jax.grad / jax.hessian).

The tail is checked AT THE KERNEL'S OWN planned U (the oracle tail is evaluated there), so the
comparison is not polluted by iLQR accept decisions; the iLQR part has its own tests.
Tolerances: loss, B, Hessian 1e-4 relative (row-wise); H = solve(A, B) and what is derived from it
1e-4 x cond-number headroom, stated per test -- the reference solves the same fp32 system without
regularisation (policy/optimizers.py:67)."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.norm import cost_trainer
from gan_mpc_b200.norm import runner as norm_runner
from gan_mpc_b200.policy import optimizers as opt
from oracle import bilevel as obl
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(cfg, seed, B):
    p, x0, U0, goal = util.case(cfg, seed, B=B)
    rng = np.random.Generator(np.random.PCG64(seed + 7))
    desired = (goal + 0.05 * rng.standard_normal(goal.shape)).astype(np.float32)
    return p, x0, U0[:, 0].copy(), goal, desired


@pytest.mark.parametrize("cfg,B,maxiter", [(util.SMALL, 37, 0), (util.ODD, 33, 0), (util.MID, 8, 0),
                                           (util.SMALL, 12, 3), (util.ODD, 6, 2)])
def test_bilevel_tail_matches_autodiff_oracle(cfg, B, maxiter, built_lib):
    p, x0, U0, goal, desired = _case(cfg, 61, B)
    h = util.make_handle(cfg, p)
    op = util.to_oracle(p)
    o = h.bilevel_l2(dev(x0), dev(U0), dev(goal), dev(desired), maxiter=maxiter, want_hessian=True)
    T, m, n = cfg["T"], cfg["m"], cfg["n"]
    U = o["U"].double().cpu()
    if maxiter == 0:
        assert torch.equal(o["U"].cpu(), torch.from_numpy(U0))
    worst = dict(loss=0.0, B=0.0, hess=0.0)
    nb = min(B, 6)
    for b in range(nb):
        loss, Bv, A, H, grad = obl.bilevel_tail(util.tt(x0[b]), U[b], util.tt(goal[b]), util.tt(desired[b]), op)
        rel = lambda k, a, ref: worst.__setitem__(k, max(worst[k], float((a.double().cpu() - ref).norm() / (ref.norm() + 1e-30))))
        rel("loss", o["loss"][b:b + 1], loss[None])
        rel("B", o["B"][b], Bv)
        rel("hess", o["hessian"][b], A)
        cond = float(torch.linalg.cond(A))
        e_H = float((o["H"][b].double().cpu() - H).norm() / H.norm())
        e_g = float((o["grad_mpc_weights"][b].double().cpu() - grad["mpc_weights"]).norm() / grad["mpc_weights"].norm())
        print(f"b={b}: cond(A) {cond:.2e}, H rel err {e_H:.2e}, grad mpc_weights rel err {e_g:.2e}")
        # fp32 LU: error <= ~ cond * 2^-24 * growth; headroom 20x
        bound = max(TOL, 20 * cond * 6e-8)
        assert e_H < bound, (e_H, bound)
        assert e_g < 2 * bound, (e_g, bound)
    print("worst relative errors:", worst)
    assert worst["loss"] < TOL and worst["B"] < TOL and worst["hess"] < TOL


def test_loss_and_grad_and_optimizer_entry_points(built_lib):
    """BaseMPC.loss_and_grad / bilevel_optimization / cost_hessian_wrt_control / cost_vjp with the
    reference's signatures on the l2 YAML config (C1 dims), against the autodiff oracle at the
    kernel's own plans; dynamics / expert leaves get exactly zero (Appendix D.3)."""
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    x_size, u_size, B = 3, 1, 16
    policy, _, _ = norm_runner.get_policy(config, x_size, u_size)
    params = norm_runner.get_params(policy, config, x_size, u_size)
    gen = torch.Generator().manual_seed(9)
    hx = torch.randn(B, 2, x_size, generator=gen).cuda()
    T = config.mpc.horizon
    goal, init_u = policy.get_goal_states_init_actions(hx, params)
    batch_y = (goal + 0.1 * torch.randn(B, T + 1, x_size, generator=gen).cuda()).contiguous()
    policy.trajax_ilqr_kwargs = dict(policy.trajax_ilqr_kwargs, maxiter=4)
    loss, grads = policy.loss_and_grad(hx, params, (batch_y,))
    out = policy.last_bilevel
    assert set(grads) == set(params)
    assert float(grads["dynamics_params"]["params"]["Dense_0"]["kernel"].abs().sum()) == 0.0
    from tests.test_gpu_api import oracle_params
    op = oracle_params(params)
    acc_m, acc_W, losses = 0.0, [0.0] * 3, []
    for b in range(B):
        l, _, _, Hh, g = obl.bilevel_tail(hx[b, -1].cpu().double(), out["U"][b].cpu().double(),
                                          goal[b].cpu().double(), batch_y[b].cpu().double(), op)
        losses.append(l)
        acc_m = acc_m + g["mpc_weights"] / B
        acc_W = [a + w / B for a, w in zip(acc_W, g["cost_W"])]
    assert abs(float(loss) - float(torch.stack(losses).mean())) < TOL * abs(float(loss))
    e = float((grads["mpc_weights"].cpu().double() - acc_m).norm() / acc_m.norm())
    print("mean grad mpc_weights rel err", e)
    assert e < 1e-3
    for i in range(3):
        k = grads["cost_params"]["params"][f"Dense_{i}"]["kernel"].cpu().double()
        e = float((k - acc_W[i]).norm() / acc_W[i].norm())
        print(f"mean grad cost Dense_{i} kernel rel err", e)
        assert e < 1e-3
    # bilevel_optimization, unbatched, reference argument order
    l0, low, high, itr = opt.bilevel_optimization(policy.cost, policy.dynamics, policy.loss, hx[0, -1], init_u[0],
                                                  params, (goal[0],), (), (batch_y[0],), policy.trajax_ilqr_kwargs)
    assert l0.shape == () and low.shape == (T, u_size) and itr.shape == ()
    assert torch.equal(l0, out["loss"][0]) and high["mpc_weights"].shape == (3,)
    # cost_hessian_wrt_control / cost_vjp at a fixed U
    cost = opt.bind(policy.cost, params, (goal[0],))
    dyn = opt.bind(policy.dynamics, params)
    Hs = opt.cost_hessian_wrt_control(cost, dyn, hx[0, -1], init_u[0])
    oH = obl.cost_hessian_wrt_control(hx[0, -1].cpu().double(), init_u[0].cpu().double(), goal[0].cpu().double(), op)
    assert Hs.shape == (T, u_size, T, u_size)
    assert float((Hs.cpu().double() - oH).norm() / oH.norm()) < TOL
    V = torch.randn(T * u_size, generator=gen).cuda()
    gv = opt.cost_vjp(policy.cost, dyn, V, hx[0, -1], init_u[0], params, (goal[0],))
    og = obl.cost_vjp(V.cpu().double(), hx[0, -1].cpu().double(), init_u[0].cpu().double(), goal[0].cpu().double(), op)
    assert float((gv["mpc_weights"].cpu().double() - og["mpc_weights"]).norm() / og["mpc_weights"].norm()) < TOL
    k = gv["cost_params"]["params"]["Dense_1"]["kernel"].cpu().double()
    assert float((k - og["cost_W"][1]).norm() / og["cost_W"][1].norm()) < TOL


def test_cost_trainer_train(built_lib):
    """norm/cost_trainer.train with the reference's signature: losses per update, Polyak blend,
    untouched masked leaves, inputs not mutated."""
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    x_size, u_size, D = 3, 1, 48
    policy, _, _ = norm_runner.get_policy(config, x_size, u_size)
    params = norm_runner.get_params(policy, config, x_size, u_size)
    policy.trajax_ilqr_kwargs = dict(policy.trajax_ilqr_kwargs, maxiter=5)
    policy.planner_kwargs["method"] = "ilqr"
    gen = torch.Generator().manual_seed(11)
    X = torch.randn(D, 2, x_size, generator=gen).cuda()
    T = config.mpc.horizon
    Y = (X[:, -1:, :] + 0.1 * torch.cumsum(torch.randn(D, T + 1, x_size, generator=gen).cuda(), 1)).contiguous()
    dataset = ((X[:32], Y[:32]), (X[32:], Y[32:]))
    copt, opt_state = norm_runner.get_optimizer(params, config.mpc.train.cost.no_grads, lr=1e-3)
    assert sorted(copt.trained) == ["cost_params", "mpc_weights"]
    before = utils.tree_clone(params)
    out = cost_trainer.train((policy, copt), opt_state, params, dataset, num_updates=2, batch_size=16,
                             polyak_factor=0.9, key=0, id=1)
    new_params, opt_state, train_losses, test_losses, minutes = out
    assert len(train_losses) == 2 and len(test_losses) == 2 and minutes >= 0.0
    assert opt_state["count"] == 4
    assert torch.equal(params["mpc_weights"], before["mpc_weights"])          # inputs not mutated
    # masked leaves get a zero update; the Polyak blend 0.9 x + 0.1 x (cost_trainer.py:88-92 blends ALL
    # leaves) only re-rounds them
    assert torch.allclose(new_params["dynamics_params"]["params"]["Dense_0"]["kernel"],
                          before["dynamics_params"]["params"]["Dense_0"]["kernel"], rtol=1e-6, atol=0)
    d = (new_params["mpc_weights"] - before["mpc_weights"]).abs()
    assert float(d.max()) > 0 and float(d.max()) < 4 * 1e-3 * 0.1 * 1.01     # 4 Adam steps of lr 1e-3, Polyak 0.1
    assert all(np.isfinite(train_losses)) and all(np.isfinite(test_losses))


def test_generator_loss_bilevel_through_the_critic(built_lib):
    """JS_MPC.generator_loss_and_grad (gan/js_policy.py:60-74) = loss_and_grad with the generator
    loss: gmpc_ilqr -> gmpc_critic_input_grad (BPTT of the LSTM critic to its inputs) ->
    gmpc_bilevel_tail, against the autodiff oracle with loss = -critic_logit(X)."""
    from gan_mpc_b200.gan import runner as gan_runner
    from oracle import critic as ocritic
    from tests.test_gpu_api import oracle_params
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "gan_hyperparameters.yaml"))
    x_size, u_size, B = 3, 1, 12
    policy, _, _ = gan_runner.get_policy(config, x_size, u_size)
    params = gan_runner.get_params(policy, config, x_size, u_size)
    cm = policy.critic_model.model
    F, L, H = cm.lstm_features, cm.num_layers, cm.num_hidden_units
    T = config.mpc.horizon
    gen = torch.Generator().manual_seed(21)
    hx = torch.randn(B, 2, x_size, generator=gen).cuda()
    flat = policy.critic_flat(params)
    # the critic's input gradient on its own
    xs = torch.randn(B, T + 1, x_size, generator=gen).cuda()
    score, dx = policy.critic_handle(x_size).critic_input_grad(xs, flat)
    xo = xs.cpu().double().requires_grad_(True)
    so = ocritic.critic_logit(xo, flat.cpu().double(), x_size, F, L, H)
    (dxo,) = torch.autograd.grad(so.sum(), xo)
    assert util.rel_rows(score[:, None], so.detach()[:, None]) < TOL
    assert util.rel_rows(dx, dxo) < TOL
    # the bilevel gradient with the generator loss
    policy.trajax_ilqr_kwargs = dict(policy.trajax_ilqr_kwargs, maxiter=3)
    actual = torch.zeros(B, T + 1, x_size).cuda()   # generator_loss only reads its last-axis size
    loss, grads = policy.generator_loss_and_grad(hx, params, (actual,))
    goal, init_u = policy.get_goal_states_init_actions(hx, params)
    _, out = policy._bilevel(hx[:, -1], init_u, params, goal, actual)
    op = oracle_params(params)
    fo = flat.cpu().double()
    loss_fn = lambda X: -ocritic.critic_logit(X, fo, x_size, F, L, H)
    acc_m, losses = 0.0, []
    for b in range(B):
        l, Bv, A, Hh, g = obl.bilevel_tail(hx[b, -1].cpu().double(), out["U"][b].cpu().double(),
                                           goal[b].cpu().double(), None, op, loss_fn)
        losses.append(l)
        acc_m = acc_m + g["mpc_weights"] / B
        assert float((out["B"][b].cpu().double() - Bv).norm() / Bv.norm()) < TOL
        assert float((out["H"][b].cpu().double() - Hh).norm() / Hh.norm()) < max(TOL, 20 * float(torch.linalg.cond(A)) * 6e-8)
    assert abs(float(loss) - float(torch.stack(losses).mean())) < TOL * max(1.0, abs(float(loss)))
    e = float((grads["mpc_weights"].cpu().double() - acc_m).norm() / acc_m.norm())
    print("generator-loss mean grad mpc_weights rel err", e)
    assert e < 1e-3
    assert float(cm.flatten(grads["critic_params"]).abs().sum()) == 0.0   # cost side only (Appendix D.3)


def test_bilevel_solve_residual_at_c2_dims(built_lib):
    """C2 dims (192 x 192 Hessian per state): size-independent properties of the tail -- the Hessian is
    symmetric, H satisfies A H = B to fp32 LU accuracy, and the given-direction mode (cost_vjp's V = H)
    reproduces the tangent quantities of the solve mode."""
    from gan_mpc_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["C2"], K=1, B=40)
    p = synthetic.planner_params(0, **cfg)
    x0, U0, goal = synthetic.planner_inputs(0, **cfg)
    rng = np.random.Generator(np.random.PCG64(3))
    desired = (goal + 0.05 * rng.standard_normal(goal.shape)).astype(np.float32)
    h = util.make_handle(cfg, p)
    o = h.bilevel_l2(dev(x0), dev(U0[:, 0].copy()), dev(goal), dev(desired), maxiter=1, want_hessian=True)
    A, H, Bv = o["hessian"].double(), o["H"].double().reshape(40, -1), o["B"].double().reshape(40, -1)
    assert float((A - A.transpose(1, 2)).abs().max()) == 0.0
    res = (torch.einsum("bij,bj->bi", A, H) - Bv).norm(dim=1) / Bv.norm(dim=1)
    cond = torch.linalg.cond(A)
    print(f"solve residual max {float(res.max()):.2e}, cond(A) median {float(cond.median()):.2e}")
    assert float(res.max()) < 1e-3
    o2 = h.bilevel_l2(dev(x0), o["U"], dev(goal), dev(desired), maxiter=0, V=o["H"].contiguous())
    assert util.rel_rows(o2["dxT"], o["dxT"].double().cpu()) < 1e-4
    assert util.rel_rows(o2["grad_mpc_weights"], o["grad_mpc_weights"].double().cpu()) < 1e-4


def test_gan_generator_training_through_cost_trainer(built_lib):
    """The GAN's generator half (gan/runner.py:161 calls cost_trainer.train with the JS policy):
    loss_and_grad = bilevel gradient of the generator loss; cost side moves, critic does not."""
    from gan_mpc_b200.gan import runner as gan_runner
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "gan_hyperparameters.yaml"))
    x_size, u_size, D = 3, 1, 40
    policy, _, _ = gan_runner.get_policy(config, x_size, u_size)
    params = gan_runner.get_params(policy, config, x_size, u_size)
    policy.trajax_ilqr_kwargs = dict(policy.trajax_ilqr_kwargs, maxiter=3)
    policy.planner_kwargs["method"] = "ilqr"
    gen = torch.Generator().manual_seed(31)
    T = config.mpc.horizon
    X = torch.randn(D, 2, x_size, generator=gen).cuda()
    Y = (X[:, -1:, :] + 0.1 * torch.cumsum(torch.randn(D, T + 1, x_size, generator=gen).cuda(), 1)).contiguous()
    copt, opt_state = gan_runner.get_optimizer(params, config.mpc.train.cost.no_grads, lr=1e-3)
    assert "critic_params" not in copt.trained and "cost_params" in copt.trained
    before = utils.tree_clone(params)
    new_params, opt_state, tr, te, minutes = cost_trainer.train(
        (policy, copt), opt_state, params, ((X[:32], Y[:32]), (X[32:], Y[32:])), num_updates=1, batch_size=16,
        polyak_factor=0.9, key=0, id=1)
    assert len(tr) == 1 and len(te) == 1 and np.isfinite(tr[0]) and np.isfinite(te[0])
    k = lambda p: p["cost_params"]["params"]["Dense_1"]["kernel"]
    assert float((k(new_params) - k(before)).abs().max()) > 0
    assert torch.allclose(policy.critic_flat(new_params), policy.critic_flat(before), rtol=1e-6, atol=0)
