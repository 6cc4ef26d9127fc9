"""GPU parity of gmpc_ilqr (the reference's own planner step, trajax iLQR: policy/optimizers.py:10-21)
against oracle/ilqr.py on identical seeded inputs, through the C ABI.

iLQR is a discrete algorithm (accept/reject of line-search trials, stopping tests), and on the
random-init ReLU dynamics its trajectory through U-space is not continuous in the arithmetic, in
ANY fp32 implementation.  So parity is layered:
  * one iteration from the same start (maxiter 0 and 1): 1e-4 relative, row-wise, on X, U, obj,
    gradient, adjoints and the Jacobians; iteration counts exact;
  * a few iterations: the same bar with counted outliers (rows whose accept decision flipped);
  * a full run: invariants that hold for any correct implementation (descent, obj = J(U) and
    gradient = dJ/dU recomputed by the fp64 oracle at the RETURNED U, iteration bounds) and
    agreement with the oracle in distribution."""

import os

import numpy as np
import pytest
import torch

from oracle import ilqr as oilqr
from oracle import planner as oracle
from tests import util

pytestmark = pytest.mark.gpu
TOL = 1e-4
GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _setup(cfg, seed, B):
    p, x0, U0, goal = util.case(cfg, seed, B=B)
    h = util.make_handle(cfg, p)
    return h, util.to_oracle(p), x0, U0[:, 0].copy(), goal


@pytest.mark.parametrize("cfg,B", [(util.SMALL, 1), (util.SMALL, 37), (util.MID, 40), (util.ODD, 33),
                                   (util.WIDE, 8)])
def test_ilqr_maxiter0_is_rollout_and_linearisation(cfg, B, built_lib):
    """No iteration: X = rollout, obj = J, gradient/adjoints from the explicit Jacobians (trajax
    `adjoint`), lqr A/B = dynamics Jacobians."""
    h, op, x0, U0, goal = _setup(cfg, 41, B)
    X, U, obj, g, lam, (A, Bm), it = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=0, want_lqr=True)
    margin = [None]
    oX, oJ, odU, olam = oracle.objective_grad(util.tt(x0), util.tt(U0), util.tt(goal), op, margin)
    assert torch.equal(U.cpu(), torch.from_numpy(U0))
    assert int(it.abs().max()) == 0
    assert util.rel_rows(X, oX) < TOL and util.rel_rows(obj[:, None], oJ[:, None]) < TOL
    away = margin[0] > 1e-6
    print(f"rows within 1e-6 of a ReLU kink: {int((~away).sum())} of {B}")
    assert int(away.sum()) >= (B + 1) // 2
    aw = away.cuda()
    assert util.rel_rows(g[aw], odU[away]) < TOL
    assert util.rel_rows(lam[aw], olam[away]) < TOL
    *_, oA, oB = oilqr.lqr_params(oX, util.tt(U0), util.tt(goal), op)
    assert util.rel_rows(A[aw], oA[away]) < TOL
    assert util.rel_rows(Bm[aw], oB[away]) < TOL


@pytest.mark.parametrize("cfg,B,maxiter", [(util.SMALL, 37, 1), (util.MID, 40, 1), (util.ODD, 33, 1),
                                           (util.SMALL, 64, 3), (util.MID, 33, 3)])
def test_ilqr_iterations_match_oracle(cfg, B, maxiter, built_lib):
    h, op, x0, U0, goal = _setup(cfg, 43, B)
    X, U, obj, g, lam, _, it = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=maxiter)
    oX, oU, oobj, og, olam, _, oit = oilqr.ilqr(util.tt(x0), util.tt(U0), util.tt(goal), op, maxiter=maxiter)
    same_it = (it.cpu() == oit)
    print(f"iteration counts equal: {int(same_it.sum())} of {B}")
    assert int(same_it.sum()) >= B - max(1, B // 16)
    frac = 0.1 if maxiter == 1 else 0.2
    util.assert_rows_close("U", U, oU, tol=TOL, outlier_frac=frac, cap=2.0)
    util.assert_rows_close("X", X, oX, tol=TOL, outlier_frac=frac, cap=2.0)
    util.assert_rows_close("obj", obj[:, None], oobj[:, None], tol=TOL, outlier_frac=frac, cap=1.0)


def test_ilqr_golden(built_lib):
    z = np.load(os.path.join(GOLDEN, "ilqr_small.npz"))
    cfg = util.SMALL
    p, x0, U0, goal = util.case(cfg, int(z["seed"]), B=int(z["B"]))
    h = util.make_handle(cfg, p)
    X, U, obj, g, lam, _, it = h.ilqr(dev(x0), dev(U0[:, 0]), dev(goal), maxiter=int(z["maxiter"]))
    e = util.rel_each(obj[:, None], torch.from_numpy(z["obj"])[:, None])
    print("golden obj rel err", e.tolist(), "iterations", it.tolist(), z["iteration"].tolist())
    assert float(e.median()) < TOL and float(e.max()) < 5e-2


@pytest.mark.parametrize("cfg,B,maxiter", [(util.SMALL, 70, 100), (util.MID, 48, 40)])
def test_ilqr_full_run_invariants(cfg, B, maxiter, built_lib):
    h, op, x0, U0, goal = _setup(cfg, 47, B)
    X, U, obj, g, lam, _, it = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=maxiter)
    _, J0 = oracle.objective(util.tt(x0), util.tt(U0), util.tt(goal), op)
    # descent: accepted steps are strict decreases of the fp32 objective
    assert bool((obj.double().cpu() <= J0 * (1 + 1e-5)).all())
    assert 0 <= int(it.min()) and int(it.max()) <= maxiter
    # the returned tuple is self-consistent at the returned U (fp64 oracle as the judge)
    margin = [None]
    oX, oJ, odU, olam = oracle.objective_grad(util.tt(x0), U.double().cpu(), util.tt(goal), op, margin)
    assert util.rel_rows(X, oX) < TOL
    assert util.rel_rows(obj[:, None], oJ[:, None]) < TOL
    away = margin[0] > 1e-6
    assert int(away.sum()) >= (B + 1) // 2
    assert util.rel_rows(g[away.cuda()], odU[away]) < 5e-4
    # agreement with the oracle's own run in distribution (the two trajectories through U-space
    # separate at the first flipped accept decision; both descend to comparable objectives)
    _, _, oobj, _, _, _, oit = oilqr.ilqr(util.tt(x0), util.tt(U0), util.tt(goal), op, maxiter=maxiter)
    ratio = obj.double().cpu() / oobj
    print(f"obj / oracle obj: median {float(ratio.median()):.4f}, min {float(ratio.min()):.3f}, "
          f"max {float(ratio.max()):.3f}; iterations kernel median {float(it.float().median())}, "
          f"oracle median {float(oit.float().median())}; improvement median "
          f"{float((obj.double().cpu() / J0).median()):.3f}")
    assert 0.9 < float(ratio.median()) < 1.1


def test_ilqr_lane_independence(built_lib):
    """vmap semantics: a trajectory's result does not depend on its tile mates (bitwise)."""
    cfg = util.SMALL
    h, op, x0, U0, goal = _setup(cfg, 53, 40)
    full = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=10)
    for b in (0, 35):
        one = h.ilqr(dev(x0[b:b + 1]), dev(U0[b:b + 1]), dev(goal[b:b + 1]), maxiter=10)
        for i in (0, 1, 2, 3, 4, 6):
            assert torch.equal(full[i][b], one[i][0]), i


def test_ilqr_rejects_unsupported_options(built_lib):
    h, op, x0, U0, goal = _setup(util.SMALL, 3, 2)
    with pytest.raises(NotImplementedError):
        h.ilqr(dev(x0), dev(U0), dev(goal), make_psd=True)
    with pytest.raises(TypeError):
        h.ilqr(dev(x0), dev(U0), dev(goal), bogus=1)


TINY = dict(n=2, m=3, T=1, dyn_layers=2, dyn_hidden=12, cost_layers=2, cost_hidden=9, cost_fout=3)  # T=1, m > n
LONG = dict(n=4, m=2, T=24, dyn_layers=3, dyn_hidden=64, cost_layers=3, cost_hidden=32, cost_fout=5)


@pytest.mark.parametrize("cfg,B,maxiter", [(TINY, 35, 4), (LONG, 20, 2), (util.WIDE, 8, 2)])
def test_ilqr_edge_shapes(cfg, B, maxiter, built_lib):
    """horizon 1 with more actions than states (G = R + B^T P B has a direction held only by the 1e-5
    eigenvalue of the pseudo-Huber Hessian), a long horizon, hidden width 512 (two output tiles per
    thread).  These are ill-conditioned: the ORACLE's own fp32 run misses the fp64 one by more than
    1e-4, so the bar is the fp32 noise floor measured by the oracle itself (x4 median, x8 max), and
    iteration counts exact up to one row in eight."""
    h, op, x0, U0, goal = _setup(cfg, 59, B)
    X, U, obj, g, lam, _, it = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=maxiter)
    oX, oU, oobj, og, olam, _, oit = oilqr.ilqr(util.tt(x0), util.tt(U0), util.tt(goal), op, maxiter=maxiter)
    f32 = torch.float32
    p32 = {k: ([w.to(f32) for w in v] if isinstance(v, list) else v.to(f32)) for k, v in op.items()}
    _, sU, sobj, *_ = oilqr.ilqr(util.tt(x0, f32), util.tt(U0, f32), util.tt(goal, f32), p32, maxiter=maxiter)
    assert int((it.cpu() == oit).sum()) >= B - max(1, B // 8)
    for name, k, s32, o64 in (("U", U, sU, oU), ("obj", obj[:, None], sobj[:, None], oobj[:, None])):
        ek, es = util.rel_each(k, o64), util.rel_each(s32, o64)
        print(f"{name}: kernel median {float(ek.median()):.2e} max {float(ek.max()):.2e}; "
              f"fp32 oracle median {float(es.median()):.2e} max {float(es.max()):.2e}")
        assert float(ek.median()) <= max(TOL / 4, 4 * float(es.median()))
        assert float(ek.max()) <= max(TOL, 8 * float(es.max()))


def test_ilqr_alpha0_below_alpha_min_and_bad_shapes(built_lib):
    """alpha_0 <= alpha_min: no trial can run, one iteration is counted and U is unchanged (the
    while_loop of line_search_ddp never enters); a state size whose Riccati tile does not fit
    shared memory is refused, not silently truncated."""
    from gan_mpc_b200 import _lib
    h, op, x0, U0, goal = _setup(util.SMALL, 61, 5)
    X, U, obj, g, lam, _, it = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=7, alpha_0=1e-5, alpha_min=1e-4)
    o = oilqr.ilqr(util.tt(x0), util.tt(U0), util.tt(goal), op, maxiter=7, alpha_0=1e-5, alpha_min=1e-4)
    assert torch.equal(U.cpu(), torch.from_numpy(U0)) and torch.equal(it.cpu(), o[6]) and int(it.max()) == 0
    big = dict(util.SMALL, n=40, m=2)
    p, bx0, bU0, bgoal = util.case(big, 1, B=2)
    hb = util.make_handle(big, p)
    with pytest.raises(_lib.GmpcError, match="shared memory"):
        hb.ilqr(dev(bx0), dev(bU0[:, 0].copy()), dev(bgoal), maxiter=1)


def test_ilqr_full_size_properties(built_lib):
    """BASELINE config C2 at full size (4096 states, n=17, m=6, T=32): properties that need no oracle
    run -- descent, the returned objective / gradient agree with the first-order kernels evaluated at
    the returned U (cross-kernel consistency), iteration bounds, bitwise determinism."""
    from gan_mpc_b200 import synthetic
    cfg = dict(synthetic.CONFIGS["C2"], K=1)
    p = synthetic.planner_params(0, **cfg)
    x0, U0, goal = synthetic.planner_inputs(0, **cfg)
    h = util.make_handle(cfg, p)
    dx0, dU0, dgoal = dev(x0), dev(U0[:, 0].copy()), dev(goal)
    h.set_path("ffma")
    J0, *_ = h.objective_grad(dx0, dU0, dgoal, want_grad=False, want_X=False)
    X, U, obj, g, lam, _, it = h.ilqr(dx0, dU0, dgoal, maxiter=3)
    assert bool((obj <= J0 * (1 + 1e-5)).all()) and float((obj / J0).median()) < 0.9
    assert int(it.min()) >= 0 and int(it.max()) <= 3
    J1, dU1, X1, lam1 = h.objective_grad(dx0, U.contiguous(), dgoal, want_lam=True)
    assert util.rel_rows(X, X1.double().cpu()) < 1e-5
    assert util.rel_rows(obj[:, None], J1.double().cpu()[:, None]) < 1e-5
    util.assert_rows_close("gradient vs BPTT kernel", g, dU1.double().cpu(), tol=1e-3, outlier_frac=0.02, cap=1.0)
    again = h.ilqr(dx0, dU0, dgoal, maxiter=3)
    assert torch.equal(again[1], U) and torch.equal(again[2], obj) and torch.equal(again[6], it)


def test_ilqr_host_buffers_and_empty_batch(built_lib):
    """gmpc_ilqr_host (host pointers, synchronous) returns what the device call returns; B = 0 is a
    no-op for the iLQR, bilevel, dynamics-fit and expert entry points."""
    h, op, x0, U0, goal = _setup(util.SMALL, 67, 35)
    dv = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=4)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a))
    ho = h.ilqr_host(t(x0), t(U0), t(goal), maxiter=4)
    for i in (0, 1, 2, 3, 4, 6):
        assert torch.equal(ho[i], dv[i].cpu()), i
    e = lambda *s: torch.empty(*s, device="cuda")
    out = h.ilqr(e(0, 3), e(0, 5, 1), e(0, 6, 3), maxiter=3)
    assert out[1].shape == (0, 5, 1)
    o = h.bilevel_l2(e(0, 3), e(0, 5, 1), e(0, 6, 3), e(0, 6, 3), maxiter=1)
    assert o["H"].shape == (0, 5, 1)
    loss, act, cot = h.dynamics_fit(e(0, 4, 3), e(0, 4, 1), e(0, 4, 3), 0.9, True, [4, 200, 200, 200, 3])
    assert loss.shape == (0,) and act[0].shape == (4, 0)


@pytest.mark.parametrize("cfg,B,maxiter", [(util.SMALL, 37, 1), (util.SMALL, 64, 3), (util.MID, 33, 2)])
def test_ilqr_gradient_lag_option(cfg, B, maxiter, built_lib):
    """gmpc_ilqr_options::gradient_lag -- trajax's loop body as recalled (`adjoint` on the lqr tuple unpacked before
    the step): the returned gradient / adjoints are those of the iterate BEFORE the last step and the grad_norm test
    lags with them; trajectory, objective and lqr are unchanged for lanes that stop on maxiter or alpha."""
    h, op, x0, U0, goal = _setup(cfg, 43, B)
    a = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=maxiter)
    b = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=maxiter, gradient_lag=True)
    o = oilqr.ilqr(util.tt(x0), util.tt(U0), util.tt(goal), op, maxiter=maxiter, gradient_lag=True)
    # no lane converges on the gradient norm within these few iterations: same steps, same trajectories
    assert torch.equal(a[6], b[6]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
    same_it = (b[6].cpu() == o[6])
    assert int(same_it.sum()) >= B - max(1, B // 16)
    frac = 0.1 if maxiter == 1 else 0.2
    util.assert_rows_close("gradient (lagging)", b[3], o[3], tol=TOL, outlier_frac=frac, cap=2.0)
    util.assert_rows_close("adjoints (lagging)", b[4], o[4], tol=TOL, outlier_frac=frac, cap=2.0)
    # with one iteration the lagging gradient is the gradient at U0 itself
    if maxiter == 1:
        g0 = h.ilqr(dev(x0), dev(U0), dev(goal), maxiter=0)[3]
        moved = (b[6] == 1)
        assert torch.equal(b[3][moved], g0[moved])


def test_ilqr_gradient_lag_runs_one_more_iteration_on_convergence(built_lib):
    """A problem iLQR solves (zero weights: linear dynamics x' = x, the staging cost alone): without the lag the loop
    stops as soon as the gradient at the new iterate is below the threshold, with it one iteration later."""
    cfg = util.SMALL
    p, x0, U0, goal = util.case(cfg, 5, B=16)
    for k in ("dyn_W", "cost_W"):
        p[k] = [np.zeros_like(w) for w in p[k]]
    h = util.make_handle(cfg, p)
    U0s = (0.01 * U0[:, 0]).astype(np.float32)
    a = h.ilqr(dev(x0), dev(U0s), dev(goal), maxiter=50, grad_norm_threshold=1e-3)
    b = h.ilqr(dev(x0), dev(U0s), dev(goal), maxiter=50, grad_norm_threshold=1e-3, gradient_lag=True)
    o = oilqr.ilqr(util.tt(x0), util.tt(U0s), util.tt(goal), util.to_oracle(p), maxiter=50, grad_norm_threshold=1e-3,
                   gradient_lag=True)
    print("iterations without / with lag / oracle with lag:", a[6].tolist(), b[6].tolist(), o[6].tolist())
    conv = (a[6] < 50).cpu()
    assert bool(conv.any())
    assert bool((b[6].cpu()[conv] >= a[6].cpu()[conv]).all())
    assert int((b[6].cpu() == o[6]).sum()) >= 14
