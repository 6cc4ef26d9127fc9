"""Seeded synthetic planner inputs and flax-default weight initialisers (host side, numpy).

JAX's threefry streams cannot be reproduced without JAX, so the parity contract is "the same
arrays are fed to the oracle and to the kernels" (SURVEY.md 8d).  Generator:
numpy.random.Generator(PCG64(seed)); seed 0 is the reference's `seed` (config/*.yaml:3).
Everything is returned as float32 numpy arrays.
"""

import numpy as np

# BASELINE.json configs.  lr / method are builder-chosen (the reference planner
# is trajax iLQR); they are recorded in every benchmark line.
CONFIGS = {
    "C1": dict(B=1, K=1, n=3, m=1, T=5, iters=20, dyn_layers=4, dyn_hidden=200, cost_layers=3,
               cost_hidden=128, cost_fout=10),
    "C2": dict(B=4096, K=1, n=17, m=6, T=32, iters=20, dyn_layers=4, dyn_hidden=200,
               cost_layers=3, cost_hidden=128, cost_fout=10),
    "C3": dict(B=8192, K=1, n=3, m=1, T=5, iters=20, dyn_layers=4, dyn_hidden=200, cost_layers=3,
               cost_hidden=128, cost_fout=10, critic_features=64, critic_layers=1,
               critic_hidden=64, critic_batch=128),
    "C4": dict(B=16384, K=8, n=17, m=6, T=64, iters=20, dyn_layers=3, dyn_hidden=512,
               cost_layers=3, cost_hidden=512, cost_fout=10),
    "C5": dict(B=262144, K=1, n=17, m=6, T=32, iters=50, dyn_layers=4, dyn_hidden=200,
               cost_layers=3, cost_hidden=128, cost_fout=10),
}
MPC_WEIGHTS = (-2.0, 3.0, -3.0)  # action, state, terminal (config/*.yaml:44-46)


def _trunc_normal(rng, shape):
    """standard normal truncated to [-2, 2] (rejection sampling)."""
    out = rng.standard_normal(shape)
    bad = np.abs(out) > 2.0
    while bad.any():
        out[bad] = rng.standard_normal(int(bad.sum()))
        bad = np.abs(out) > 2.0
    return out


def lecun_normal(rng, fan_in, fan_out):
    """flax nn.Dense default kernel init: truncated normal (+-2 sigma), variance 1/fan_in."""
    std = np.sqrt(1.0 / fan_in) / 0.87962566103423978
    return (_trunc_normal(rng, (fan_in, fan_out)) * std).astype(np.float32)


def orthogonal(rng, n):
    """flax recurrent kernel default: orthogonal."""
    a = rng.standard_normal((n, n))
    q, r = np.linalg.qr(a)
    return (q * np.sign(np.diag(r))).astype(np.float32)


def mlp_params(rng, dims):
    """Dense stack with flax defaults: lecun_normal kernels [in,out], zero biases."""
    Ws = [lecun_normal(rng, dims[i], dims[i + 1]) for i in range(len(dims) - 1)]
    bs = [np.zeros(dims[i + 1], np.float32) for i in range(len(dims) - 1)]
    return Ws, bs


def dyn_dims(n, m, layers, hidden):
    return [n + m] + [hidden] * (layers - 1) + [n]


def cost_dims(n, layers, hidden, fout):
    return [n] + [hidden] * (layers - 1) + [fout]


def planner_params(seed, n, m, dyn_layers, dyn_hidden, cost_layers, cost_hidden, cost_fout,
                   bias_scale=0.0, **_):
    """dict(dyn_W, dyn_b, cost_W, cost_b, mpc_weights) of float32 arrays.
    bias_scale > 0 draws non-zero biases (tests only; flax init is zeros)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    dW, db = mlp_params(rng, dyn_dims(n, m, dyn_layers, dyn_hidden))
    cW, cb = mlp_params(rng, cost_dims(n, cost_layers, cost_hidden, cost_fout))
    if bias_scale > 0:
        db = [(bias_scale * rng.standard_normal(b.shape)).astype(np.float32) for b in db]
        cb = [(bias_scale * rng.standard_normal(b.shape)).astype(np.float32) for b in cb]
    return dict(dyn_W=dW, dyn_b=db, cost_W=cW, cost_b=cb,
                mpc_weights=np.asarray(MPC_WEIGHTS, np.float32))


def planner_inputs(seed, B, K, n, m, T, **_):
    """x0 ~ N(0,1); goal[0]=x0 then a 0.1-step random walk; U0 = tanh(N(0,1))  (SURVEY.md 8d)."""
    rng = np.random.Generator(np.random.PCG64(seed + 1000003))
    x0 = rng.standard_normal((B, n)).astype(np.float32)
    steps = 0.1 * rng.standard_normal((B, T, n))
    goal = np.concatenate([x0[:, None, :], x0[:, None, :] + np.cumsum(steps, axis=1)], axis=1)
    U0 = np.tanh(rng.standard_normal((B, K, T, m)))
    return x0, U0.astype(np.float32), goal.astype(np.float32)


def critic_params_flat(seed, n, F, layers, hidden):
    """Flat critic vector in the include/gmpc.h layout with flax OptimizedLSTMCell defaults:
    input kernels lecun_normal, recurrent kernels orthogonal, zero biases; Dense head lecun."""
    rng = np.random.Generator(np.random.PCG64(seed + 2000003))
    Wi = np.concatenate([lecun_normal(rng, n, F) for _ in range(4)], axis=1)
    Wh = np.concatenate([orthogonal(rng, F) for _ in range(4)], axis=1)
    parts = [Wi.ravel(), Wh.ravel(), np.zeros(4 * F, np.float32)]
    d = F
    for _ in range(layers - 1):
        parts += [lecun_normal(rng, d, hidden).ravel(), np.zeros(hidden, np.float32)]
        d = hidden
    parts += [lecun_normal(rng, d, 1).ravel(), np.zeros(1, np.float32)]
    return np.concatenate(parts).astype(np.float32)


def expert_params_flat(seed, shapes, lstm_features):
    """Flat expert-network vector (include/gmpc.h layout) with flax defaults.  `shapes` is the
    [(path, shape)] list of expert/nn.py:_shapes(): LSTM input kernels lecun_normal, recurrent kernels
    orthogonal (per gate), Dense kernels lecun_normal, biases zero."""
    rng = np.random.Generator(np.random.PCG64(seed + 4000003))
    F = lstm_features
    parts = []
    for path, shape in shapes:
        if len(shape) == 1:
            parts.append(np.zeros(shape, np.float32))
        elif path[:2] == ("lstm", "Wh"):
            parts.append(np.concatenate([orthogonal(rng, F) for _ in range(4)], axis=1).ravel())
        elif path[:2] == ("lstm", "Wi"):
            parts.append(np.concatenate([lecun_normal(rng, shape[0], F) for _ in range(4)], axis=1).ravel())
        else:
            parts.append(lecun_normal(rng, shape[0], shape[1]).ravel())
    return np.concatenate(parts).astype(np.float32)


def critic_dataset(seed, D, T1, n):
    """Labelled trajectories: +1 goal-style random walks, -1 a second family with drift (a
    stand-in for planner outputs when none are supplied).  Returns xseq [2D,T1,n], label [2D]."""
    rng = np.random.Generator(np.random.PCG64(seed + 3000003))
    x0 = rng.standard_normal((2 * D, 1, n))
    steps = 0.1 * rng.standard_normal((2 * D, T1 - 1, n))
    steps[D:] += 0.05
    xs = np.concatenate([x0, x0 + np.cumsum(steps, axis=1)], axis=1).astype(np.float32)
    label = np.concatenate([np.ones(D), -np.ones(D)]).astype(np.float32)
    return xs, label
