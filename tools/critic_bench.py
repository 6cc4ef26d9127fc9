"""Critic (JS-GAN discriminator) update throughput on one B200 -- BASELINE config C3.
Dataset: 16 384 labelled trajectories (8192 "expert" random walks +1, 8192 planner-style -1), T+1 = 6
rows of n = 3 (gan_hyperparameters.yaml dims), LSTM F = 64, minibatch 128 sampled with replacement,
128 sequential steps per update (gan/critic_trainer.py:48-65).  Each step = fused gather + LSTM
forward/backward + BCE + flat gradient, deterministic reduction, clip_by_global_norm(100) + Adam.
Prints one JSON line (steps/s, us per step, launches per step) and checks the first step's loss
and gradient against the CPU oracle.   usage: python tools/critic_bench.py [--updates 4]"""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from gan_mpc_b200 import _lib, synthetic
from oracle import critic as ocritic

ap = argparse.ArgumentParser()
ap.add_argument("--updates", type=int, default=4)
args = ap.parse_args()
cfg = synthetic.CONFIGS["C3"]
n, T1, F, L, H, Bc = cfg["n"], cfg["T"] + 1, cfg["critic_features"], cfg["critic_layers"], cfg["critic_hidden"], cfg["critic_batch"]
D = 2 * cfg["B"]
rng = np.random.default_rng(0)
X = np.cumsum(0.1 * rng.standard_normal((D, T1, n)), axis=1).astype(np.float32) + rng.standard_normal((D, 1, n)).astype(np.float32)
Y = np.concatenate([np.ones(D // 2), -np.ones(D // 2)]).astype(np.float32)
dev = torch.device("cuda", 0)
h = _lib.Handle(n, cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"], cfg["cost_layers"], cfg["cost_hidden"],
                cfg["cost_fout"], critic_features=F, critic_layers=L, critic_hidden=H, device=0)
P = h.critic_param_count
flat0 = torch.from_numpy(synthetic.critic_params_flat(0, n, F, L, H)) if hasattr(synthetic, "critic_params_flat") else \
    torch.from_numpy((0.1 * rng.standard_normal(P)).astype(np.float32))
dX, dY = torch.from_numpy(X).to(dev), torch.from_numpy(Y).to(dev)
steps = D // Bc
g = torch.Generator(device=dev); g.manual_seed(1)

ap_mode = os.environ.get("CRITIC_BENCH_MODE", "scan")   # "scan": one C-ABI call per update; "steps": 2 calls per step


def update(flat, mom, vel, count, perm):
    if ap_mode == "scan":
        losses = h.critic_train_scan(dX, dY, perm, flat, mom, vel, step0=count, lr=1e-5, max_norm=100.0)
        return count + perm.shape[0], losses.mean()
    losses = []
    for s in range(perm.shape[0]):
        loss, grad = h.critic_loss_grad(dX, dY, flat, inv_count=1.0 / Bc, perm=perm[s])
        count += 1
        h.clip_adam_step(flat, grad, mom, vel, step=count, lr=1e-5, max_norm=100.0)
        losses.append(loss)
    return count, torch.stack(losses).mean()

flat = flat0.to(dev).clone(); mom = torch.zeros_like(flat); vel = torch.zeros_like(flat)
perm = torch.randint(0, D, (steps, Bc), generator=g, device=dev, dtype=torch.int32)
# parity of the first step against the oracle (loss and gradient)
loss, grad = h.critic_loss_grad(dX, dY, flat, inv_count=1.0 / Bc, perm=perm[0])
idx = perm[0].cpu().long()
ol, og = ocritic.critic_loss_and_grad(torch.from_numpy(X)[idx].double(), torch.from_numpy(Y)[idx].double(), flat0.double(), n, F, L, H)
rel = float((grad.double().cpu() - og).norm() / og.norm())
assert abs(float(loss) - float(ol)) < 1e-5 * max(1.0, abs(float(ol))) and rel < 1e-4, (float(loss), float(ol), rel)
count, _ = update(flat, mom, vel, 0, perm)          # warm-up update
torch.cuda.synchronize()
l0 = h.launch_count
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.updates):
    perm = torch.randint(0, D, (steps, Bc), generator=g, device=dev, dtype=torch.int32)
    count, ml = update(flat, mom, vel, count, perm)
e1.record(); torch.cuda.synchronize()
wall = time.perf_counter() - t0
ms = e0.elapsed_time(e1)
nsteps = args.updates * steps
print(json.dumps({"metric": "critic minibatch steps/sec (gather + LSTM fwd/bwd + BCE + clip + Adam)", "value": nsteps / (ms * 1e-3),
                  "unit": "steps/s", "us_per_step": 1e3 * ms / nsteps, "trajectories_per_s": nsteps * Bc / (ms * 1e-3),
                  "updates": args.updates, "steps_per_update": steps, "batch": Bc, "dataset": D,
                  "launches_per_step": (h.launch_count - l0) / nsteps, "wall_s": wall, "mode": ap_mode,
                  "config": {"workload": "C3 critic: n=3, T+1=6, LSTM F=64, num_layers=1", "params": int(P)},
                  "parity_first_step": {"loss_abs_err": abs(float(loss) - float(ol)), "grad_rel_err": rel},
                  "final_mean_loss": float(ml)}))
