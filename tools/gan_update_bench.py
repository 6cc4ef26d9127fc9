"""BASELINE config C3 through the reference's trainer entry points (gan_hyperparameters.yaml dims:
n=3, m=1, T=5, critic LSTM 64, batch 128): critic_trainer.get_dataset (planner on every sample),
critic_trainer.train (2 updates of 2D/128 minibatch steps), cost_trainer.train_cost_parameters
(generator: bilevel gradient of the generator loss, full trajax iLQR options) and
dynamics_trainer.train_params.  Wall-clock per call after a warm-up call, one B200.
    python tools/gan_update_bench.py [--D 8192]"""
import argparse
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_mpc_b200 import utils  # noqa: E402
from gan_mpc_b200.config import load_config  # noqa: E402
from gan_mpc_b200.gan import critic_trainer, runner as gan_runner  # noqa: E402
from gan_mpc_b200.norm import cost_trainer, dynamics_trainer  # noqa: E402


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        out = fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--D", type=int, default=8192)
    a = ap.parse_args()
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "gan_hyperparameters.yaml"))
    n, m, T, D = 3, 1, config.mpc.horizon, a.D
    policy, _, _ = gan_runner.get_policy(config, n, m)
    params = gan_runner.get_params(policy, config, n, m)
    gen = torch.Generator().manual_seed(0)
    X = torch.randn(D, 2, n, generator=gen).cuda()
    Y = (X[:, -1:, :] + 0.1 * torch.cumsum(torch.randn(D, T + 1, n, generator=gen).cuda(), 1)).contiguous()
    split = int(0.9 * D)
    true_dataset = ((X[:split], Y[:split]), (X[split:], Y[split:]))
    res = {}
    dt, _ = timed(lambda: critic_trainer.get_dataset(policy, params, true_dataset, key=0))
    res["get_dataset"] = dict(seconds=dt, states_planned=D, planner=policy.planner_kwargs["method"])
    copt, cstate = gan_runner.get_optimizer(params, config.mpc.train.critic.no_grads, lr=config.mpc.train.critic.learning_rate)
    bs = config.mpc.train.critic.batch_size
    dt, out = timed(lambda: critic_trainer.train((policy, copt), cstate, params, true_dataset, num_updates=2,
                                                 batch_size=bs, key=0, id=1), reps=1)
    steps = 2 * (2 * split // bs)
    res["critic_train"] = dict(seconds=dt, minibatch_steps=steps, batch=bs, steps_per_s=steps / dt,
                               includes="get_dataset (planner) + 2 updates + 2 test losses")
    gopt, gstate = gan_runner.get_optimizer(params, config.mpc.train.cost.no_grads, lr=config.mpc.train.cost.learning_rate)
    perm = torch.randint(0, split, (4, 128), generator=torch.Generator().manual_seed(1)).cuda()
    dt, out = timed(lambda: cost_trainer.train_cost_parameters((policy, gopt), gstate, params, perm,
                                                               true_dataset[0]), reps=1)
    res["generator_train"] = dict(seconds=dt, minibatch_steps=4, batch=128, steps_per_s=4 / dt,
                                  includes="iLQR (maxiter 100) + critic input gradient + bilevel tail + cost-MLP "
                                           "mixed VJP + clipped Adam per step")
    S = 16
    xs = torch.randn(D, S, n, generator=gen).cuda()
    us = torch.tanh(torch.randn(D, S, m, generator=gen)).cuda()
    ys = (xs + 0.1 * torch.randn(D, S, n, generator=gen).cuda()).contiguous()
    dopt, dstate = gan_runner.get_optimizer(params, config.mpc.train.dynamics.no_grads, lr=config.mpc.train.dynamics.learning_rate)
    dt, out = timed(lambda: dynamics_trainer.train_params((policy, dopt), dstate, params, (xs, us, ys), num_updates=1,
                                                          batch_size=128, discount_factor=0.9, teacher_forcing_factor=0.0,
                                                          key=0, id=1), reps=1)
    res["dynamics_train"] = dict(seconds=dt, minibatch_steps=D // 128, batch=128, window=S, steps_per_s=(D // 128) / dt,
                                 includes="free-running windows: rollout + BPTT kernel, 4 gmpc_gemm_nt, clipped Adam per step")
    print(json.dumps(dict(workload=f"C3: GAN YAML dims n={n}, m={m}, T={T}, {D} trajectories", results=res)))


if __name__ == "__main__":
    main()
