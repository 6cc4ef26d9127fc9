"""CPU self-checks that pin the oracle restatement: the reference ships no
tests, so these are the pins -- autograd vs hand adjoint, finite differences, closed forms,
and the frozen golden vectors."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import synthetic
from oracle import critic as ocritic
from oracle import planner as oracle
from tests import util

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


@pytest.mark.parametrize("cfg", [util.SMALL, util.MID, util.ODD])
def test_adjoint_matches_autograd_fp64(cfg):
    p, x0, U0, goal = util.case(cfg, 3, B=6)
    op = util.to_oracle(p)
    x0, U, goal = util.tt(x0), util.tt(U0[:, 0]), util.tt(goal)
    Ua = U.clone().requires_grad_(True)
    X, J = oracle.objective(x0, Ua, goal, op)
    (g_auto,) = torch.autograd.grad(J.sum(), Ua)
    X2, J2, dU, lam = oracle.objective_grad(x0, U, goal, op)
    assert torch.allclose(J, J2, rtol=1e-12, atol=0)
    assert torch.allclose(X, X2, rtol=1e-12, atol=1e-14)
    assert torch.allclose(g_auto, dU, rtol=1e-9, atol=1e-12)
    # adjoints: lam[t] = dJ/dX[t] holding U fixed downstream; check lam[0] via autograd on x0
    x0a = x0.clone().requires_grad_(True)
    _, Jx = oracle.objective(x0a, U, goal, op)
    (gx,) = torch.autograd.grad(Jx.sum(), x0a)
    assert torch.allclose(gx, lam[:, 0], rtol=1e-9, atol=1e-12)


def test_adjoint_matches_finite_differences():
    cfg = util.SMALL
    p, x0, U0, goal = util.case(cfg, 5, B=2)
    op = util.to_oracle(p)
    x0, U, goal = util.tt(x0), util.tt(U0[:, 0]), util.tt(goal)
    _, _, dU, _ = oracle.objective_grad(x0, U, goal, op)
    eps = 1e-6
    for t in range(cfg["T"]):
        for j in range(cfg["m"]):
            Up, Um = U.clone(), U.clone()
            Up[:, t, j] += eps
            Um[:, t, j] -= eps
            fd = (oracle.objective(x0, Up, goal, op)[1] - oracle.objective(x0, Um, goal, op)[1]) / (2 * eps)
            assert torch.allclose(fd, dU[:, t, j], rtol=1e-5, atol=1e-8)


def test_objective_term_by_term():
    """J = sum_{t<T} stage(X[t],U[t],goal[t]) + w2*||costMLP(X[T])||^2, staging at t=0 included."""
    cfg = util.ODD
    p, x0, U0, goal = util.case(cfg, 1, B=3)
    op = util.to_oracle(p)
    x0, U, goal = util.tt(x0), util.tt(U0[:, 0]), util.tt(goal)
    X, J = oracle.objective(x0, U, goal, op)
    w = torch.sigmoid(op["mpc_weights"])
    a = oracle.ALPHA
    ref = torch.zeros(3, dtype=torch.float64)
    for t in range(cfg["T"]):
        ref += w[0] * (torch.sqrt((U[:, t] ** 2).sum(-1) + a * a) - a)
        ref += w[1] * (torch.sqrt(((X[:, t] - goal[:, t]) ** 2).sum(-1) + a * a) - a)
    ref += w[2] * oracle.cost_mlp(X[:, -1], op["cost_W"], op["cost_b"])
    assert torch.allclose(J, ref, rtol=1e-12)
    assert torch.allclose(w, torch.tensor([0.11920292, 0.95257413, 0.04742587], dtype=torch.float64), atol=1e-7)


def test_dense_relu_hand_example():
    """flax Dense: y = x @ kernel + bias with kernel[in,out]; residual dynamics."""
    W0 = torch.tensor([[1.0, -1.0], [0.5, 2.0], [1.0, 1.0]], dtype=torch.float64)  # [3 in, 2 out]
    b0 = torch.tensor([0.0, -1.0], dtype=torch.float64)
    W1 = torch.tensor([[2.0, 0.0], [1.0, -1.0]], dtype=torch.float64)              # [2, n=2]
    b1 = torch.tensor([0.5, 0.5], dtype=torch.float64)
    x = torch.tensor([1.0, 2.0], dtype=torch.float64)
    u = torch.tensor([-1.0], dtype=torch.float64)
    # q=[1,2,-1]; z0 = [1+1-1, -1+4-1-1] = [1, 1]; relu -> [1,1]; out = [2+1+.5, -1+.5] = [3.5,-.5]
    nx = oracle.dynamics_mlp(x, u, [W0, W1], [b0, b1])
    assert torch.allclose(nx, torch.tensor([4.5, 1.5], dtype=torch.float64))
    c = oracle.cost_mlp(x, [W0[:2], W1], [b0, b1])
    # z0 = [1+1, -1+4-1] = [2,2]; y = [4+2+.5, -2+.5] = [6.5,-1.5]; y.y = 44.5
    assert torch.allclose(c, torch.tensor(44.5, dtype=torch.float64))


def test_adam_first_step_closed_form():
    """optax adam step 1: update = -lr * g / (|g| + eps)  (bias-corrected m = g, v = g^2)."""
    cfg = util.SMALL
    p, x0, U0, goal = util.case(cfg, 2, B=4, K=2)
    op = util.to_oracle(p)
    x0, U0, goal = util.tt(x0), util.tt(U0), util.tt(goal)
    U1, *_ = oracle.plan(x0, U0, goal, op, "adam", 1, 1e-2)
    x0k = x0[:, None].expand(4, 2, -1)
    gk = goal[:, None].expand(4, 2, -1, -1)
    _, J0, g, _ = oracle.objective_grad(x0k, U0, gk, op)
    exp_U = U0 - 1e-2 * g / (g.abs() + 1e-8)
    _, J1 = oracle.objective(x0k, exp_U, gk, op)
    idx = J1.argmin(1)
    assert torch.allclose(U1, exp_U[torch.arange(4), idx], rtol=1e-9, atol=1e-12)


def test_plan_grad_decreases_cost_and_selects_argmin():
    cfg = util.SMALL
    p, x0, U0, goal = util.case(cfg, 4, B=8, K=3)
    op = util.to_oracle(p)
    x0, U0, goal = util.tt(x0), util.tt(U0), util.tt(goal)
    _, _, J0, _, J0_all = oracle.plan(x0, U0, goal, op, "grad", 0, 1e-2)
    Ub, Xb, Jb, idx, J_all = oracle.plan(x0, U0, goal, op, "grad", 10, 1e-2)
    assert (J_all <= J0_all + 1e-12).all()
    assert torch.equal(idx.long(), J_all.argmin(1))
    assert torch.allclose(Jb, J_all.min(1).values)
    assert torch.allclose(oracle.rollout(x0, Ub, op), Xb, rtol=1e-12)


def test_l2_loss_and_grad():
    cfg = util.ODD
    p, x0, U0, goal = util.case(cfg, 6, B=3)
    op = util.to_oracle(p)
    x0, U, des = util.tt(x0), util.tt(U0[:, 0]), util.tt(goal)
    Ua = U.clone().requires_grad_(True)
    loss = oracle.l2_loss(oracle.rollout(x0, Ua, op), des)
    (g,) = torch.autograd.grad(loss.sum(), Ua)
    assert torch.allclose(g, oracle.loss_grad_wrt_control_l2(x0, U, des, op), rtol=1e-9, atol=1e-13)
    X = oracle.rollout(x0, U, op)
    ref = ((X - des) ** 2).mean(1).sum(-1)
    assert torch.allclose(loss.detach(), ref)


def test_lstm_cell_hand_example():
    """OptimizedLSTMCell math on a 1-feature, 1-input, 2-step example."""
    n, F = 1, 1
    # Wi = [ii, if, ig, io], Wh = [hi, hf, hg, ho], bh
    flat = torch.tensor([0.5, -0.5, 1.0, 2.0, 0.1, 0.2, 0.3, 0.4, 0.0, 1.0, 0.0, -1.0, 3.0, 0.25],
                        dtype=torch.float64)
    assert ocritic.critic_param_count(n, F, 1, 1) == flat.numel()
    xs = torch.tensor([[1.0], [-2.0]], dtype=torch.float64)
    sig = lambda v: 1 / (1 + np.exp(-v))
    c = h = 0.0
    for x in (1.0, -2.0):
        i = sig(0.5 * x + 0.1 * h + 0.0)
        f = sig(-0.5 * x + 0.2 * h + 1.0)
        g = np.tanh(1.0 * x + 0.3 * h + 0.0)
        o = sig(2.0 * x + 0.4 * h - 1.0)
        c = f * c + i * g
        h = o * np.tanh(c)
    want = 3.0 * h + 0.25
    got = ocritic.critic_logit(xs, flat, n, F, 1, 1)
    assert abs(float(got) - want) < 1e-12


def test_critic_loss_forms_agree():
    """softplus form == -log(where(label>0, p, 1-p)) (gan/js_policy.py:41-46); generator loss == -s."""
    n, F, L, H = 3, 8, 2, 6
    flat = util.tt(synthetic.critic_params_flat(0, n, F, L, H))
    xs, lab = synthetic.critic_dataset(0, 5, 6, n)
    xs, lab = util.tt(xs), util.tt(lab)
    s = ocritic.critic_logit(xs, flat, n, F, L, H)
    p = torch.sigmoid(s)
    ref = -torch.log(torch.where(lab > 0, p, 1 - p))
    assert torch.allclose(ocritic.critic_loss(xs, lab, flat, n, F, L, H), ref, rtol=1e-10)
    gen_ref = -torch.log(p) + torch.log(1 - p)
    assert torch.allclose(ocritic.generator_loss(xs, flat, n, F, L, H), gen_ref, rtol=1e-9, atol=1e-12)


def test_clip_adam_matches_optax_semantics():
    g = torch.tensor([300.0, -400.0], dtype=torch.float64)   # norm 500 -> scaled to norm 100
    p = torch.zeros(2, dtype=torch.float64)
    p1, m1, v1 = ocritic.clip_adam_step(p, g, torch.zeros(2, dtype=torch.float64),
                                        torch.zeros(2, dtype=torch.float64), 1, 1e-3)
    gc = g / 500.0 * 100.0
    assert torch.allclose(m1, 0.1 * gc) and torch.allclose(v1, 0.001 * gc * gc)
    assert torch.allclose(p1, -1e-3 * gc / (gc.abs() + 1e-8))
    # below the threshold the gradient is untouched
    p2, m2, _ = ocritic.clip_adam_step(p, g / 100, torch.zeros(2, dtype=torch.float64),
                                       torch.zeros(2, dtype=torch.float64), 1, 1e-3)
    assert torch.allclose(m2, 0.1 * g / 100)


def test_fp32_noise_floor_reported():
    """oracle32 vs oracle64 on the C2 dims: the floor the kernel is judged against."""
    cfg = dict(util.MID, T=32)
    p, x0, U0, goal = util.case(cfg, 0, B=16, bias_scale=0.0)
    o64 = oracle.plan(util.tt(x0), util.tt(U0), util.tt(goal), util.to_oracle(p), "adam", 20, 1e-2)
    f = torch.float32
    o32 = oracle.plan(util.tt(x0, f), util.tt(U0, f), util.tt(goal, f), util.to_oracle(p, f), "adam", 20, 1e-2)
    floor = max(util.rel_rows(o32[0], o64[0]), util.rel_rows(o32[1], o64[1]),
                util.rel_rows(o32[2][:, None], o64[2][:, None]))
    print("fp32 noise floor (U, X, J rel):", floor)
    assert floor < 1e-4


@pytest.mark.parametrize("name", ["small", "mid"])
def test_golden_vectors(name):
    """Frozen oracle outputs (tests/golden/make_golden.py): pins the restatement against drift."""
    z = np.load(os.path.join(GOLDEN, f"planner_{name}.npz"))
    cfg = {k: int(z[k]) for k in ("n", "m", "T", "dyn_layers", "dyn_hidden", "cost_layers",
                                  "cost_hidden", "cost_fout")}
    L, Lc = cfg["dyn_layers"], cfg["cost_layers"]
    p = dict(dyn_W=[z[f"dyn_W{i}"] for i in range(L)], dyn_b=[z[f"dyn_b{i}"] for i in range(L)],
             cost_W=[z[f"cost_W{i}"] for i in range(Lc)], cost_b=[z[f"cost_b{i}"] for i in range(Lc)],
             mpc_weights=z["mpc_weights"])
    op = util.to_oracle(p)
    x0, U0, goal = util.tt(z["x0"]), util.tt(z["U0"]), util.tt(z["goal"])
    X, J, dU, lam = oracle.objective_grad(x0, U0[:, 0], goal, op)
    assert np.allclose(X.numpy(), z["X"], rtol=1e-10, atol=1e-12)
    assert np.allclose(J.numpy(), z["J"], rtol=1e-10)
    assert np.allclose(dU.numpy(), z["dU"], rtol=1e-9, atol=1e-12)
    assert np.allclose(lam.numpy(), z["lam"], rtol=1e-9, atol=1e-12)
    for method in ("grad", "adam"):
        Ub, Xb, Jb, idx, Jall = oracle.plan(x0, U0, goal, op, method, int(z["iters"]), float(z["lr"]))
        assert np.allclose(Ub.numpy(), z[f"{method}_U_best"], rtol=1e-8, atol=1e-11)
        assert np.allclose(Jall.numpy(), z[f"{method}_J_all"], rtol=1e-8)
        assert np.array_equal(idx.numpy(), z[f"{method}_idx"])


def test_golden_critic():
    z = np.load(os.path.join(GOLDEN, "critic_small.npz"))
    n, F, L, H = (int(z[k]) for k in ("n", "F", "L", "H"))
    flat, xs, lab = util.tt(z["flat"]), util.tt(z["xseq"]), util.tt(z["label"])
    assert np.allclose(ocritic.critic_logit(xs, flat, n, F, L, H).numpy(), z["logit"], rtol=1e-10)
    loss, g = ocritic.critic_loss_and_grad(xs, lab, flat, n, F, L, H)
    assert np.allclose(float(loss), float(z["loss"]), rtol=1e-10)
    assert np.allclose(g.numpy(), z["grad"], rtol=1e-8, atol=1e-12)


def test_dynfit_oracle_teacher_forcing_and_discount():
    """oracle/dynfit.py: with teacher forcing the window is S independent one-step regressions;
    discount 0 keeps only step 0; free running equals teacher forcing on a window of one step."""
    from oracle import dynfit as ofit
    p, *_ = util.case(util.ODD, 3, B=1)
    op = util.to_oracle(p)
    g = torch.Generator().manual_seed(0)
    xs = torch.randn(4, 5, 5, generator=g, dtype=torch.float64)
    us = torch.randn(4, 5, 3, generator=g, dtype=torch.float64)
    ys = torch.randn(4, 5, 5, generator=g, dtype=torch.float64)
    tf = ofit.predict_loss(op, xs, us, ys, 0.9, True)
    ref = sum(0.9 ** t * ((oracle.dynamics_mlp(xs[:, t], us[:, t], op["dyn_W"], op["dyn_b"]) - ys[:, t]) ** 2).sum(-1)
              for t in range(5))
    assert torch.allclose(tf, ref)
    assert torch.allclose(ofit.predict_loss(op, xs, us, ys, 0.0, True), ofit.predict_loss(op, xs[:, :1], us[:, :1], ys[:, :1], 0.9, False))
    assert not torch.allclose(tf, ofit.predict_loss(op, xs, us, ys, 0.9, False))
