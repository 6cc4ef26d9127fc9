"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard -> per-rank work -> gather of the
best plans, and the data-parallel critic step (per-rank partial gradients, sum all-reduce, the
same update on every rank) against the single-process result."""

import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gan_mpc_b200 import parallel, synthetic
from oracle import critic as ocritic


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, fn, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ret[rank] = fn(rank, world)
    finally:
        dist.destroy_process_group()


def _run(fn, world=2):
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), fn, ret), nprocs=world, join=True)
    return [ret[r] for r in range(world)]


def _gather_job(rank, world):
    B, T, m = 37, 4, 3
    full_U = torch.arange(B * T * m, dtype=torch.float32).reshape(B, T, m)
    full_J = torch.arange(B, dtype=torch.float32) * 0.5
    full_idx = (torch.arange(B) % 5).to(torch.int32)
    lo, hi = parallel.shard_range(B)
    U, J, idx = parallel.gather_best_plans(full_U[lo:hi], full_J[lo:hi], full_idx[lo:hi], B)
    return bool(torch.equal(U, full_U) and torch.equal(J, full_J) and torch.equal(idx, full_idx))


def test_gather_best_plans_reassembles_the_batch():
    assert _run(_gather_job) == [True, True]


def _critic_dp_job(rank, world):
    """each rank: gradient of its slice of the minibatch scaled by 1/global_batch, sum
    all-reduce, identical clip+Adam -- the structure of critic_trainer.train_critic_parameters."""
    n, F, L, H, T1, Bc = 3, 8, 1, 4, 5, 16
    flat = torch.from_numpy(synthetic.critic_params_flat(0, n, F, L, H)).double()
    xs, lab = synthetic.critic_dataset(0, Bc // 2, T1, n)
    xs, lab = torch.from_numpy(xs).double(), torch.from_numpy(lab).double()
    lo, hi = parallel.shard_range(Bc)
    p = flat.clone().requires_grad_(True)
    part = ocritic.critic_loss(xs[lo:hi], lab[lo:hi], p, n, F, L, H).sum() / Bc
    (g,) = torch.autograd.grad(part, p)
    loss = part.detach().clone().reshape(1)
    parallel.allreduce_sum_(g)
    parallel.allreduce_sum_(loss)
    new, _, _ = ocritic.clip_adam_step(flat, g, torch.zeros_like(flat), torch.zeros_like(flat), 1, 1e-3)
    return new.numpy(), float(loss)


def test_data_parallel_critic_step_equals_single_process():
    n, F, L, H, T1, Bc = 3, 8, 1, 4, 5, 16
    flat = torch.from_numpy(synthetic.critic_params_flat(0, n, F, L, H)).double()
    xs, lab = synthetic.critic_dataset(0, Bc // 2, T1, n)
    loss, g = ocritic.critic_loss_and_grad(torch.from_numpy(xs).double(), torch.from_numpy(lab).double(),
                                           flat, n, F, L, H)
    want, _, _ = ocritic.clip_adam_step(flat, g, torch.zeros_like(flat), torch.zeros_like(flat), 1, 1e-3)
    out = _run(_critic_dp_job)
    for new, l in out:
        assert np.allclose(new, want.numpy(), rtol=1e-12, atol=1e-15)
        assert abs(l - float(loss)) < 1e-12
    assert np.array_equal(out[0][0], out[1][0])          # replicas stay bit-identical
