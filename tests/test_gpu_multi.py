"""Multi-GPU (one process per GPU, NCCL): sharded planning == single-GPU planning bit for bit,
and the data-parallel critic step == the single-GPU full-batch step.  Skipped with < 2 GPUs."""

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
flag = torch.tensor([int(ok_plan), int(ok_critic)], device=dev)
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
    assert "RESULT [1, 1]" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
