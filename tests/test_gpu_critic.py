"""GPU parity of the critic path (LSTM forward, BCE, flat gradient, clip+Adam) vs the oracle."""

import os

import numpy as np
import pytest
import torch

from gan_mpc_b200 import synthetic
from oracle import critic as ocritic
from tests import util

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def handle(n, F, L, H, T):
    cfg = dict(util.SMALL, n=n, T=T)
    p = synthetic.planner_params(0, **cfg)
    return util.make_handle(cfg, p, critic=dict(F=F, L=L, H=H))


@pytest.mark.parametrize("n,F,L,H,T1,Bc", [(3, 64, 1, 64, 6, 128), (17, 64, 1, 64, 33, 128),
                                           (5, 24, 3, 40, 9, 37), (3, 16, 2, 8, 6, 12)])
def test_critic_loss_and_grad(n, F, L, H, T1, Bc, built_lib):
    h = handle(n, F, L, H, T1 - 1)
    flat = synthetic.critic_params_flat(1, n, F, L, H)
    rng = np.random.default_rng(3)
    flat = flat + (0.05 * rng.standard_normal(flat.shape)).astype(np.float32)
    assert h.critic_param_count == flat.size == ocritic.critic_param_count(n, F, L, H)
    xs, lab = synthetic.critic_dataset(2, (Bc + 1) // 2, T1, n)
    xs, lab = xs[:Bc], lab[:Bc]
    lab = lab[rng.permutation(Bc)].copy()
    logit = h.critic_forward(dev(xs), dev(flat))
    ologit = ocritic.critic_logit(util.tt(xs), util.tt(flat), n, F, L, H)
    assert float((logit.double().cpu() - ologit).abs().max()) < 1e-4 * float(ologit.abs().max() + 1)
    loss, grad = h.critic_loss_grad(dev(xs), dev(lab), dev(flat))
    oloss, ograd = ocritic.critic_loss_and_grad(util.tt(xs), util.tt(lab), util.tt(flat), n, F, L, H)
    assert abs(float(loss) - float(oloss)) < 1e-5 * abs(float(oloss))
    assert float((grad.double().cpu() - ograd).norm() / ograd.norm()) < 1e-4
    # loss only (calculate_loss, gan/critic_trainer.py:41-45)
    loss2, none = h.critic_loss_grad(dev(xs), dev(lab), dev(flat), want_grad=False)
    assert none is None and float(loss2) == float(loss)
    # gathered minibatch == explicit gather (gan/critic_trainer.py:55-56)
    perm = rng.integers(0, Bc, size=Bc).astype(np.int32)          # with replacement
    l3, g3 = h.critic_loss_grad(dev(xs), dev(lab), dev(flat), perm=dev(perm))
    l4, g4 = h.critic_loss_grad(dev(xs[perm]), dev(lab[perm]), dev(flat))
    assert torch.equal(l3, l4) and torch.equal(g3, g4)
    # deterministic reduction
    l5, g5 = h.critic_loss_grad(dev(xs), dev(lab), dev(flat))
    assert torch.equal(loss, l5) and torch.equal(grad, g5)


def test_clip_adam_step(built_lib):
    h = handle(3, 16, 1, 8, 5)
    rng = np.random.default_rng(0)
    P = 5000
    for scale in (1.0, 1000.0):      # below / above the global-norm threshold of 100
        prm = rng.standard_normal(P).astype(np.float32)
        g = (scale * rng.standard_normal(P)).astype(np.float32)
        mom = (0.1 * rng.standard_normal(P)).astype(np.float32)
        vel = np.abs(0.1 * rng.standard_normal(P)).astype(np.float32)
        dp, dm, dv = dev(prm), dev(mom), dev(vel)
        h.clip_adam_step(dp, dev(g), dm, dv, step=7, lr=1e-3)
        op, om, ov = ocritic.clip_adam_step(util.tt(prm), util.tt(g), util.tt(mom), util.tt(vel), 7, 1e-3)
        assert float((dm.double().cpu() - om).norm() / om.norm()) < 1e-6
        assert float((dv.double().cpu() - ov).norm() / ov.norm()) < 1e-6
        assert float((dp.double().cpu() - op).abs().max()) < 1e-6


def test_train_critic_scan(built_lib):
    """gan/critic_trainer.py:48-65: sequential minibatch steps (gather -> loss/grad -> clip+adam)."""
    n, F, L, H, T1, Bc, steps = 3, 64, 1, 64, 6, 32, 5
    h = handle(n, F, L, H, T1 - 1)
    flat = synthetic.critic_params_flat(5, n, F, L, H)
    xs, lab = synthetic.critic_dataset(5, 64, T1, n)
    rng = np.random.default_rng(1)
    perm = rng.integers(0, xs.shape[0], size=(steps, Bc)).astype(np.int32)
    dflat = dev(flat)
    dm, dv = torch.zeros_like(dflat), torch.zeros_like(dflat)
    dx, dl, dperm = dev(xs), dev(lab), dev(perm)
    losses = []
    for s in range(steps):
        loss, g = h.critic_loss_grad(dx, dl, dflat, perm=dperm[s])
        h.clip_adam_step(dflat, g, dm, dv, step=s + 1, lr=1e-3)
        losses.append(loss)
    of, om, ov, _, oloss = ocritic.train_critic_parameters(
        util.tt(flat), torch.zeros(flat.size, dtype=torch.float64), torch.zeros(flat.size, dtype=torch.float64),
        0, torch.from_numpy(perm).long(), util.tt(xs), util.tt(lab), 1e-3, n, F, L, H)
    assert float((dflat.double().cpu() - of).abs().max()) < 2e-5
    assert abs(float(torch.stack(losses).mean()) - float(oloss)) < 1e-5
    # the single-call scan (gmpc_critic_train_scan) enqueues exactly the same kernels
    sflat = dev(flat)
    sm, sv = torch.zeros_like(sflat), torch.zeros_like(sflat)
    slosses = h.critic_train_scan(dx, dl, dperm, sflat, sm, sv, step0=0, lr=1e-3)
    # (same reduction order; the fused tail kernel may contract FMAs differently: last-bit tolerance)
    for a, b in ((sflat, dflat), (sm, dm), (sv, dv), (slosses, torch.stack(losses).reshape(-1))):
        assert torch.allclose(a, b, rtol=2e-6, atol=1e-9)
    assert float((sflat.double().cpu() - of).abs().max()) < 2e-5


def test_golden_critic(built_lib):
    z = np.load(os.path.join(GOLDEN, "critic_small.npz"))
    n, F, L, H = (int(z[k]) for k in ("n", "F", "L", "H"))
    h = handle(n, F, L, H, z["xseq"].shape[1] - 1)
    loss, grad = h.critic_loss_grad(dev(z["xseq"]), dev(z["label"]), dev(z["flat"]))
    assert abs(float(loss) - float(z["loss"])) < 1e-5 * abs(float(z["loss"]))
    assert float((grad.double().cpu() - util.tt(z["grad"])).norm() / np.linalg.norm(z["grad"])) < 1e-4
