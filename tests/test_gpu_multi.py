"""Multi-GPU (one process per GPU, NCCL): sharded planning (first-order and iLQR) == single-GPU
planning bit for bit, the data-parallel critic step == the single-GPU full-batch step, and the
data-parallel bilevel gradient == the single-GPU batch mean.  Skipped with < 2 GPUs."""

import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys
sys.path.insert(0, os.environ["GMPC_ROOT"])
import numpy as np, torch, torch.distributed as dist
from gan_mpc_b200 import parallel, synthetic
from tests import util
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = util.SMALL
p, x0, U0, goal = util.case(cfg, 3, B=101, K=2)
h = util.make_handle(cfg, p, device=local, critic=dict(F=16, L=1, H=8))
t = lambda a: torch.from_numpy(a)
U, J, idx = parallel.plan_sharded(h, t(x0), t(U0), t(goal), iters=4)
dev = torch.device("cuda", local)
full = h.plan(t(x0).to(dev), t(U0).to(dev), t(goal).to(dev), iters=4)
ok_plan = torch.equal(U, full[0]) and torch.equal(J, full[2]) and torch.equal(idx, full[3])
# data-parallel critic step
n, F, L, H, T1, Bc = cfg["n"], 16, 1, 8, 6, 64
flat = t(synthetic.critic_params_flat(0, n, F, L, H)).to(dev)
xs, lab = synthetic.critic_dataset(0, Bc // 2, T1, n)
xs, lab = t(xs).to(dev), t(lab).to(dev)
perm = torch.arange(Bc, dtype=torch.int32, device=dev)
lo, hi = parallel.shard_range(Bc)
loss, g = h.critic_loss_grad(xs, lab, flat, inv_count=1.0 / Bc, perm=perm[lo:hi].contiguous())
parallel.allreduce_sum_(g); parallel.allreduce_sum_(loss)
loss1, g1 = h.critic_loss_grad(xs, lab, flat)
ok_critic = float((g - g1).abs().max()) < 1e-6 * float(g1.abs().max() + 1e-30) + 1e-9 and abs(float(loss) - float(loss1)) < 1e-6
# sharded iLQR == single-GPU iLQR (lanes are independent of their tile mates)
Ui, Ji, iti = parallel.ilqr_sharded(h, t(x0), t(U0[:, 0].copy()), t(goal), maxiter=5)
fi = h.ilqr(t(x0).to(dev), t(U0[:, 0].copy()).to(dev), t(goal).to(dev), maxiter=5)
ok_ilqr = torch.equal(Ui, fi[1]) and torch.equal(Ji, fi[2]) and torch.equal(iti, fi[6])
# data-parallel bilevel gradient (BaseMPC.loss_and_grad): global mean on every rank == one-GPU mean
from gan_mpc_b200 import utils
from gan_mpc_b200.config import load_config
from gan_mpc_b200.norm import runner as norm_runner
config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
# the expert NETWORK proposes per row (the synthetic stand-in draws per batch position)
policy, _, _ = norm_runner.get_policy(config, 3, 1, expert_model="network")
policy.trajax_ilqr_kwargs = dict(policy.trajax_ilqr_kwargs, maxiter=3)
params = norm_runner.get_params(policy, config, 3, 1, load_expert=False)
gen = torch.Generator().manual_seed(5)
hx = torch.randn(37, 2, 3, generator=gen).to(dev)
by = torch.randn(37, config.mpc.horizon + 1, 3, generator=gen).to(dev)
l_dp, g_dp = policy.loss_and_grad(hx, params, (by,))
saved = (dist.get_rank, dist.get_world_size)
parallel.rank_world = lambda: (0, 1)       # the same call without sharding
parallel.allreduce_sum_ = lambda x: x
l_1, g_1 = policy.loss_and_grad(hx, params, (by,))
k = lambda g: g["cost_params"]["params"]["Dense_1"]["kernel"]
ok_bl = (abs(float(l_dp) - float(l_1)) < 1e-5 * abs(float(l_1)) and
         float((k(g_dp) - k(g_1)).norm() / k(g_1).norm()) < 1e-4 and
         float((g_dp["mpc_weights"] - g_1["mpc_weights"]).norm() / g_1["mpc_weights"].norm()) < 1e-4)
flag = torch.tensor([int(ok_plan), int(ok_critic), int(ok_ilqr), int(ok_bl)], device=dev)
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("RESULT", flag.tolist())
dist.destroy_process_group()
'''


def test_sharded_plan_and_dp_critic_match_single_gpu(built_lib, tmp_path):
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    world = 2
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    env = dict(os.environ, GMPC_ROOT=ROOT)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1",
                          f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
                          "--master-port", str(port), str(script)],
                         capture_output=True, text=True, env=env, timeout=600)
    assert "RESULT [1, 1, 1, 1]" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
