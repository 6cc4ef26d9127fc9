#!/usr/bin/env python
"""bench.py -- planned states/sec of the fused rollout+BPTT+update planner (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload C2]

A "step" is one gmpc_plan call over one batch of synthetic start states (workload C2 of
BASELINE.json by default: 4096 states, n=17, m=6, T=32, 20 Adam planning iterations, random-init
flax-default MLPs).  `value` is whole-job states/s with inputs resident in HBM (CUDA events, L2
flushed between timed steps); `e2e` is the same metric through the host-buffer C-ABI call
gmpc_plan_host (pinned host tensors, H2D + D2H inside the timed region).  With N>1 (torchrun, one
rank per GPU) every rank plans its own B states (weak scaling, no data-path collective) and the
best plans are all-gathered over NCCL inside the step, as the north star prescribes.

The timed path proves its own validity: the fp16 operand-range counter of the tensor-core kernels is read
on every rank after the timed loop (`clamped_ranks`, must be 0: no `value` is printed otherwise), `path` is
the kernel the timed loop ran, and at N > 1 rank 0 re-plans a slice of rank 1's shard and compares it
bitwise with what rank 1 gathered (`shard_parity`).  `parity` compares a sample of the timed plans with the
fp64 oracle (rank 0, N = 1).  `configs` holds one or two timed steps of the other BASELINE configs (C5 as
262 144 / N states per rank: the strong-scaling split the north star names) and `critic` the data-parallel
critic step (gather + LSTM fwd/bwd + BCE -> NCCL all-reduce of the flat gradient -> clip + Adam).

`--impl reference` times the reference's CPU path: the JAX stack cannot be installed offline, so
this is the fp32 oracle port (oracle/planner.py, whole-batch matmuls, all host threads) on a
bounded sample of the same workload.
"""

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

from gan_mpc_b200 import synthetic  # noqa: E402

METRIC = "planned states/sec (fused rollout+BPTT+update)"
METRIC_ILQR = "planned states/sec (trajax-iLQR mode, gmpc_ilqr)"
UNIT = "states/s"
L2_FLUSH_BYTES = 256 << 20


def flops_per_state(c):
    """true dims, 1 MAC = 2 FLOP, forward + input-adjoint backward."""
    dd = synthetic.dyn_dims(c["n"], c["m"], c["dyn_layers"], c["dyn_hidden"])
    cd = synthetic.cost_dims(c["n"], c["cost_layers"], c["cost_hidden"], c["cost_fout"])
    m_dyn = sum(a * b for a, b in zip(dd[:-1], dd[1:]))
    m_cost = sum(a * b for a, b in zip(cd[:-1], cd[1:]))
    f_iter = 4 * c["T"] * m_dyn + 4 * m_cost
    f_final = 2 * c["T"] * m_dyn + 2 * m_cost
    return c["K"] * (c["iters"] * f_iter + f_final)


def bytes_per_state(c):
    n, m, T, K = c["n"], c["m"], c["T"], c["K"]
    return 4 * (n + (T + 1) * n + K * T * m) + 4 * (T * m + (T + 1) * n + 1) + 4


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return dict(hbm_gbs=p["hbm_gbs"], bf16_tflops=p["bf16_tflops"],
                    bf16_tflops_sustained=p.get("bf16_tflops_sustained"), source="measured")
    return dict(hbm_gbs=6650.0, bf16_tflops=1590.0, bf16_tflops_sustained=1400.0, source="fallback")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.thread.join(timeout=2)
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4)
                          if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


def oracle_params(p, dtype):
    return {k: ([torch.from_numpy(w).to(dtype) for w in v] if isinstance(v, list)
                else torch.from_numpy(v).to(dtype)) for k, v in p.items()}


def time_cpu_port(cfg, params, x0, U0, goal, lr, target_s=15.0, reps=1):
    """fp32 oracle port on a bounded sample of the workload, all host threads."""
    from oracle import planner as oracle
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    op = oracle_params(params, torch.float32)
    f = lambda a, b: torch.from_numpy(np.ascontiguousarray(a[:b]))
    probe = min(64, x0.shape[0])
    t0 = time.perf_counter()
    oracle.plan(f(x0, probe), f(U0, probe), f(goal, probe), op, "adam", cfg["iters"], lr)
    dt = time.perf_counter() - t0
    sample = int(min(x0.shape[0], max(probe, probe * target_s / max(dt, 1e-3))))
    sample = max(probe, (sample // 64) * 64)
    # the sample is at most the whole workload; when that takes less than the target, repeat it
    # so that the figure still comes from about target_s of CPU work (mean over the repeats)
    tot, n = 0.0, 0
    while n < reps or (tot < 0.6 * target_s and n < 20):
        t0 = time.perf_counter()
        oracle.plan(f(x0, sample), f(U0, sample), f(goal, sample), op, "adam", cfg["iters"], lr)
        tot += time.perf_counter() - t0
        n += 1
    best = tot / n
    return dict(value=sample / best, unit=UNIT, cores=cores, kind="port",
                sample=f"{sample} of {x0.shape[0]} states of workload, {n} pass(es) of {best:.2f} s each "
                       f"(fp32 torch-CPU oracle port, all {cores} host threads; JAX reference not "
                       f"installable offline)"), best, sample


ILQR_KW = dict(maxiter=100, grad_norm_threshold=1e-4, alpha_0=1.0, alpha_min=0.00005)  # policy/eval.py:10-20


def time_cpu_ilqr(cfg, params, x0, U0, goal, target_s=15.0):
    """fp32 oracle port of trajax iLQR (oracle/ilqr.py) on a bounded sample, all host threads."""
    from oracle import ilqr as oilqr
    cores = len(os.sched_getaffinity(0))
    torch.set_num_threads(cores)
    op = oracle_params(params, torch.float32)
    f = lambda a, b: torch.from_numpy(np.ascontiguousarray(a[:b]))
    probe = min(8, x0.shape[0])
    t0 = time.perf_counter()
    oilqr.ilqr(f(x0, probe), f(U0[:, 0], probe), f(goal, probe), op, **ILQR_KW)
    dt = time.perf_counter() - t0
    sample = int(min(x0.shape[0], max(probe, 8 * int(probe * target_s / max(dt, 1e-3) / 8))))
    t0 = time.perf_counter()
    oilqr.ilqr(f(x0, sample), f(U0[:, 0], sample), f(goal, sample), op, **ILQR_KW)
    best = time.perf_counter() - t0
    return dict(value=sample / best, unit=UNIT, cores=cores, kind="port",
                sample=f"{sample} of {x0.shape[0]} states of workload, one pass of {best:.2f} s (fp32 torch-CPU "
                       f"oracle port of trajax iLQR, all {cores} host threads; JAX/trajax not installable offline)"), best, sample


def run_reference(args, cfg, rank):
    if rank != 0:
        return
    lr = args.lr
    if args.planner == "ilqr":
        params = synthetic.planner_params(0, **cfg)
        x0, U0, goal = synthetic.planner_inputs(0, **cfg)
        tot_t = tot_s = 0.0
        for i in range(args.warmup + args.steps):
            cb, dt, sample = time_cpu_ilqr(cfg, params, x0, U0, goal, target_s=args.ref_seconds)
            if i >= args.warmup:
                tot_t, tot_s = tot_t + dt, tot_s + sample
        value = tot_s / tot_t
        cb["value"] = value
        print(json.dumps({"impl": "reference", "metric": METRIC_ILQR, "value": value, "unit": UNIT, "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * tot_t / max(args.steps, 1),
                          "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                          "data": "synthetic", "config": workload_config(args, cfg, lr), "path": "cpu-oracle-port",
                          "cpu_baseline": cb,
                          "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}),
              flush=True)
        return
    params = synthetic.planner_params(0, **cfg)
    x0, U0, goal = synthetic.planner_inputs(0, **cfg)
    times = []
    cb = None
    for i in range(args.warmup + args.steps):
        cb, dt, sample = time_cpu_port(cfg, params, x0, U0, goal, lr,
                                       target_s=args.ref_seconds, reps=1)
        if i >= args.warmup:
            times.append((dt, sample))
    tot_t = sum(t for t, _ in times)
    tot_s = sum(s for _, s in times)
    value = tot_s / tot_t
    cb["value"] = value
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT,
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * tot_t / len(times), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, lr), "path": "cpu-oracle-port",
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, lr, path=None):
    return {"workload": f"{args.workload}: B={cfg['B']} start states x K={cfg['K']} candidates, "
                        f"n={cfg['n']}, m={cfg['m']}, horizon T={cfg['T']}, {cfg['iters']} planning "
                        f"iterations, dyn MLP {synthetic.dyn_dims(cfg['n'], cfg['m'], cfg['dyn_layers'], cfg['dyn_hidden'])}, "
                        f"cost MLP {synthetic.cost_dims(cfg['n'], cfg['cost_layers'], cfg['cost_hidden'], cfg['cost_fout'])}",
            "states_per_gpu": cfg["B"],
            "planner": ({"method": "ilqr", **ILQR_KW} if args.planner == "ilqr" else
                        {"method": "adam", "lr": lr, "b1": 0.9, "b2": 0.999, "eps": 1e-8}),
            "note": ("the reference's own planner step: trajax iLQR with the options of policy/eval.py:10-20"
                     if args.planner == "ilqr" else
                     "first-order planner on the reference's objective (reference planner is trajax iLQR)"),
            "weights": "random-init flax defaults (lecun_normal, zero bias), seed 0",
            "cache": f"L2 flushed between timed steps ({L2_FLUSH_BYTES >> 20} MiB write)",
            "parallelism": f"dp{args.gpus} (start states sharded, no data-path collective; "
                                         "best plans all-gathered over NCCL)" if args.gpus > 1 else "single GPU"}


def make_handle(cfg, params, local, critic=False):
    from gan_mpc_b200 import _lib
    dev = torch.device("cuda", local)
    kw = {}
    if critic:
        kw = dict(critic_features=cfg["critic_features"], critic_layers=cfg["critic_layers"],
                  critic_hidden=cfg["critic_hidden"])
    h = _lib.Handle(cfg["n"], cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"],
                    cfg["cost_layers"], cfg["cost_hidden"], cfg["cost_fout"], device=local, **kw)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    h.set_weights([g(w) for w in params["dyn_W"]], [g(b) for b in params["dyn_b"]],
                  [g(w) for w in params["cost_W"]], [g(b) for b in params["cost_b"]],
                  g(params["mpc_weights"]))
    return h


def timed_plans(h, d_x0, d_U0, d_goal, iters, lr, steps, warmup, flush):
    """`steps` device-timed gmpc_plan calls (CUDA events, L2 flushed in between); returns ms per call."""
    B, K = d_U0.shape[0], d_U0.shape[1]
    out = h.alloc_plan_outputs(B, K, want_J_all=False)
    for _ in range(warmup):
        h.plan(d_x0, d_U0, d_goal, method="adam", iters=iters, lr=lr, out=out, check_range=False)
        flush.fill_(1.0)
    ms = []
    for _ in range(steps):
        flush.fill_(0.0)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        h.plan(d_x0, d_U0, d_goal, method="adam", iters=iters, lr=lr, out=out, check_range=False)
        e1.record()
        e1.synchronize()
        ms.append(e0.elapsed_time(e1))
    return ms, out


def parity_summary(cfg, params, x0, U0, goal, lr, got, sample):
    """The timed plans of the first `sample` states vs the fp64 oracle (and the oracle's own fp32-vs-fp64
    floor): row-wise relative errors of the planned actions and of the final plan cost."""
    from oracle import planner as oracle
    f = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a[:sample])).to(dt)
    res = {}
    o = {}
    for name, dt in (("f64", torch.float64), ("f32", torch.float32)):
        o[name] = oracle.plan(f(x0, dt), f(U0, dt), f(goal, dt), oracle_params(params, dt), "adam", cfg["iters"], lr)

    def rows(a, b):
        a = a.double().reshape(a.shape[0], -1)
        b = b.double().reshape(b.shape[0], -1)
        return (a - b).norm(dim=1) / (b.norm(dim=1) + 1e-30)

    for key, i in (("U", 0), ("X", 1), ("J", 2)):
        ours = got[i][:sample].cpu()
        ref = o["f64"][i]
        if key == "J":
            ours, ref, flo = ours[:, None], ref[:, None], o["f32"][i][:, None]
        else:
            flo = o["f32"][i]
        e, fl = rows(ours, ref), rows(flo, ref)
        res[key] = {"median": float(e.median()), "max": float(e.max()), "rows_ge_1e-4": int((e >= 1e-4).sum()),
                    "fp32_floor_median": float(fl.median()), "fp32_floor_max": float(fl.max()),
                    "fp32_floor_rows_ge_1e-4": int((fl >= 1e-4).sum())}
    res["idx_equal"] = bool(torch.equal(got[3][:sample].cpu(), o["f64"][3]))
    res["rows"] = sample
    res["oracle"] = "oracle/planner.py fp64 (floor: the same oracle in fp32); row = one planned state"
    return res


def other_configs(args, local, world, rank, flush, peaks, dist):
    """One or two device-timed steps of the other BASELINE configs.  N = 1: C1 (latency), C3 planner, C4, C5;
    N > 1: C5 only, split as 262144 / N states per rank (strong scaling, max over ranks)."""
    dev = torch.device("cuda", local)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = {}
    names = ["C1", "C3", "C4", "C5"] if world == 1 else ["C5"]
    for name in names:
        cfg = dict(synthetic.CONFIGS[name])
        if name == "C5":
            cfg["B"] = cfg["B"] // world
        params = synthetic.planner_params(0, **cfg)
        x0, U0, goal = synthetic.planner_inputs(rank, **cfg)
        h = make_handle(cfg, params, local)
        steps = 5 if name in ("C1", "C3") else 1
        ms, _ = timed_plans(h, g(x0), g(U0), g(goal), cfg["iters"], args.lr, steps, 1, flush)
        clamped = h.range_overflow()
        t = torch.tensor([float(np.mean(ms)), float(clamped > 0)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t[:1], op=dist.ReduceOp.MAX)
            dist.all_reduce(t[1:], op=dist.ReduceOp.SUM)
        ms_step, nclamp = float(t[0]), int(t[1])
        fl = flops_per_state(cfg) * cfg["B"]
        tf = fl / (ms_step * 1e-3) / 1e12
        res[name] = {"states": cfg["B"] * world, "candidates": cfg["K"], "T": cfg["T"], "iters": cfg["iters"],
                     "hidden": cfg["dyn_hidden"], "ms_per_step": ms_step,
                     "states_per_s": cfg["B"] * world / (ms_step * 1e-3), "tflops_per_gpu": tf,
                     "frac_of_bf16_peak": tf / peaks["bf16_tflops"], "path": h.last_path, "clamped_ranks": nclamp,
                     "steps": steps}
        h.close()
        del h
        torch.cuda.empty_cache()
    return res


def acting_latency(local, calls=200):
    """C1 (BASELINE configs[0]): p50 / p99 of consecutive single-state get_optimal_action calls through the
    policy API (utils.run_dm_policy's loop, reference utils.py:254-290), for both planners."""
    import warnings
    from gan_mpc_b200 import utils
    from gan_mpc_b200.config import load_config
    from gan_mpc_b200.norm import runner as norm_runner
    config = utils.get_config(os.path.join(load_config.CONFIG_DIR, "l2_hyperparameters.yaml"))
    res = {}
    for method in ("ilqr", "adam"):
        _, policy, _ = norm_runner.get_policy(config, 3, 1)
        policy.planner_kwargs["method"] = method
        params = norm_runner.get_params(policy, config, 3, 1)
        hx = torch.randn(2, 3, generator=torch.Generator().manual_seed(0)).cuda(local)
        hu = torch.zeros(1, 1).cuda(local)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for _ in range(10):
                policy.get_optimal_action(params, hx, hu)
            torch.cuda.synchronize()
            ts = []
            for _ in range(calls):
                t0 = time.perf_counter()
                u = policy.get_optimal_action(params, hx, hu)
                u.cpu()                       # the action leaves the device every environment step
                ts.append((time.perf_counter() - t0) * 1e3)
        res[method] = {"p50_ms": float(np.percentile(ts, 50)), "p99_ms": float(np.percentile(ts, 99)), "calls": calls}
    return res


def critic_dp(args, local, world, rank, dist, steps=128):
    """Data-parallel critic step at C3 dims (gan/critic_trainer.py:48-65): every rank takes Bc / N samples of the
    minibatch (fused gather + LSTM forward/backward + BCE + flat gradient), the flat gradient is summed over NCCL,
    every rank applies the same clip_by_global_norm(100) + Adam step.  Device time, max over ranks."""
    cfg = dict(synthetic.CONFIGS["C3"])
    dev = torch.device("cuda", local)
    n, T1, F, L, H, Bc = cfg["n"], cfg["T"] + 1, cfg["critic_features"], cfg["critic_layers"], cfg["critic_hidden"], cfg["critic_batch"]
    D = cfg["B"]
    xs, lab = synthetic.critic_dataset(0, D, T1, n)
    h = make_handle(cfg, synthetic.planner_params(0, **cfg), local, critic=True)
    dX, dY = torch.from_numpy(xs).to(dev), torch.from_numpy(lab).to(dev)
    flat = torch.from_numpy(synthetic.critic_params_flat(0, n, F, L, H)).to(dev)
    mom, vel = torch.zeros_like(flat), torch.zeros_like(flat)
    gen = torch.Generator(device="cpu").manual_seed(1)     # the same permutation on every rank
    perm = torch.randint(0, 2 * D, (steps + 8, Bc), generator=gen, dtype=torch.int32).to(dev)
    lo, hi = (Bc * rank) // world, (Bc * (rank + 1)) // world

    def one(sidx, count):
        loss, grad = h.critic_loss_grad(dX, dY, flat, inv_count=1.0 / Bc, perm=perm[sidx, lo:hi].contiguous())
        if world > 1:
            dist.all_reduce(grad, op=dist.ReduceOp.SUM)
        h.clip_adam_step(flat, grad, mom, vel, step=count, lr=1e-5, max_norm=100.0)

    for i in range(8):
        one(steps + i, i + 1)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        one(i, 9 + i)
    e1.record()
    e1.synchronize()
    t = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    chk = flat.double().sum().reshape(1)
    same = True
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        allc = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(allc, chk)
        same = all(bool(torch.equal(c, allc[0])) for c in allc)     # replicated parameters stay bit-identical
    ms = float(t[0])
    return {"steps_per_s": steps / (ms * 1e-3), "us_per_step": 1e3 * ms / steps, "minibatch": Bc,
            "samples_per_rank": hi - lo, "ranks": world, "param_floats": int(flat.numel()),
            "allreduce": "NCCL sum of the flat gradient between gmpc_critic_loss_grad_gather and gmpc_clip_adam_step"
                         if world > 1 else "none (one rank)",
            "params_identical_on_all_ranks": same, "steps": steps}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(synthetic.CONFIGS))
    ap.add_argument("--path", default="auto", choices=["auto", "ffma", "tc16", "tc16s", "t128"])
    ap.add_argument("--planner", default="adam", choices=["adam", "ilqr"],
                    help="adam: the north-star first-order planner (the BASELINE metric, default); ilqr: the "
                         "reference's own step, trajax iLQR with its full options (gmpc_ilqr; secondary line)")
    ap.add_argument("--lr", type=float, default=1e-2)
    ap.add_argument("--batch", type=int, default=0, help="override states per GPU")
    ap.add_argument("--ref-seconds", type=float, default=15.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs / critic records of the C2 line")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    cfg = dict(synthetic.CONFIGS[args.workload])
    if args.batch:
        cfg["B"] = args.batch
    if args.impl == "reference":
        run_reference(args, cfg, rank)
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback); "
                         "use --impl reference for the CPU baseline arm")
    from gan_mpc_b200 import _lib
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    # each rank plans its own shard: different seed per rank, same weights everywhere
    params = synthetic.planner_params(0, **cfg)
    x0, U0, goal = synthetic.planner_inputs(rank, **cfg)
    h = _lib.Handle(cfg["n"], cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"],
                    cfg["cost_layers"], cfg["cost_hidden"], cfg["cost_fout"], device=local)
    g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    h.set_weights([g(w) for w in params["dyn_W"]], [g(b) for b in params["dyn_b"]],
                  [g(w) for w in params["cost_W"]], [g(b) for b in params["cost_b"]],
                  g(params["mpc_weights"]))
    h.set_path(args.path)
    d_x0, d_U0, d_goal = g(x0), g(U0), g(goal)
    B, K = cfg["B"], cfg["K"]
    out = h.alloc_plan_outputs(B, K, want_J_all=False)
    flush = torch.empty(L2_FLUSH_BYTES // 4, device=dev, dtype=torch.float32)
    gather_bufs = None
    if world > 1:
        gather_bufs = [torch.empty(world * t.numel(), device=dev, dtype=t.dtype)
                       for t in (out[0], out[2], out[3])]

    ilqr = args.planner == "ilqr"
    d_U0i = d_U0[:, 0].contiguous() if ilqr else None
    if ilqr and K != 1:
        raise SystemExit("--planner ilqr plans one action sequence per state (K = 1 workloads)")

    def step():
        if ilqr:
            res = h.ilqr(d_x0, d_U0i, d_goal, **ILQR_KW)
            plans = (res[1], res[2], res[6])
        else:
            # check_range=False: no host sync inside the timed region; the counter is read after the loop
            h.plan(d_x0, d_U0, d_goal, method="adam", iters=cfg["iters"], lr=args.lr, out=out, check_range=False)
            plans = (out[0], out[2], out[3])
        if world > 1:  # gather of the best plans (U*, J*, idx / iteration) -- 776 B/state at C2
            for buf, t in zip(gather_bufs, plans):
                dist.all_gather_into_tensor(buf, t.reshape(-1))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    for _ in range(max(args.warmup, 3)):
        step()
        flush.fill_(1.0)
    barrier()
    h.range_overflow()  # reset the operand-range counter: what it holds after the timed loop belongs to that loop
    if ilqr:
        h.ilqr_stats()  # reset the work counters: the timed steps are counted below
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = h.launch_count
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(args.steps)]
    barrier()
    wall0 = time.perf_counter()
    for e0, e1 in evs:
        flush.fill_(0.0)          # L2 flush, outside the per-step event pair
        e0.record()
        step()
        e1.record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = h.launch_count - launches0
    timed_path = "ilqr" if ilqr else h.last_path
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    clamped = 0 if ilqr else h.range_overflow()
    stat = torch.tensor([sum(step_ms), float(clamped > 0)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stat[:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(stat[1:], op=dist.ReduceOp.SUM)
    total_s = float(stat[0].item()) * 1e-3
    clamped_ranks = int(stat[1].item())
    clocks = sampler.stop() if rank == 0 else None

    # ---- shard parity (N > 1): rank 0 re-plans the head of rank 1's shard from the same seeded inputs and
    # compares with what rank 1 computed and gathered -- bitwise (the kernels are deterministic per tile)
    shard_parity = None
    if world > 1 and not ilqr:
        ns = min(256, B)
        if rank == 0:
            x1, U1, g1 = synthetic.planner_inputs(1, **cfg)
            o1 = h.plan(g(x1[:ns]), g(U1[:ns]), g(g1[:ns]), method="adam", iters=cfg["iters"], lr=args.lr,
                        want_J_all=False)
            Ug = gather_bufs[0].reshape(world, B, cfg["T"], cfg["m"])[1, :ns]
            Jg = gather_bufs[1].reshape(world, B)[1, :ns]
            shard_parity = {"states": ns, "of_rank": 1, "U_bitwise_equal": bool(torch.equal(o1[0], Ug)),
                            "J_bitwise_equal": bool(torch.equal(o1[2], Jg))}
        barrier()

    # ---- parity of the timed plans vs the fp64 oracle (rank 0, one GPU: about 5 s of CPU work)
    parity = None
    if rank == 0 and world == 1 and not ilqr and not args.no_cpu_baseline:
        parity = parity_summary(cfg, params, x0, U0, goal, args.lr, out, sample=min(B, 512))

    # ---- e2e through the host-buffer C-ABI call (pinned host tensors in and out)
    ilqr_work = h.ilqr_stats() if ilqr else None   # (tile outer iterations, tile rollouts) of the timed steps
    pin = lambda a: torch.from_numpy(np.ascontiguousarray(a)).pin_memory()
    h_x0, h_U0, h_goal = pin(x0), pin(U0), pin(goal)
    if ilqr:
        h_U0 = pin(U0[:, 0])
        host_call = lambda: h.ilqr_host(h_x0, h_U0, h_goal, **ILQR_KW)
    else:
        h_out = h.alloc_plan_outputs(B, K, True, device=torch.device("cpu"), pin=True)
        host_call = lambda: h.plan_host(h_x0, h_U0, h_goal, method="adam", iters=cfg["iters"], lr=args.lr,
                                        out=h_out)
    for _ in range(2):
        h_res = host_call()
    barrier()
    e2e_steps = max(3, min(args.steps, 10))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_res = host_call()
    my_e2e = time.perf_counter() - t0       # host calls are synchronous: this rank's own time
    e2e_path = "ilqr" if ilqr else h.last_path
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    per_rank = torch.zeros(world, device=dev, dtype=torch.float64)
    per_rank[rank] = 1e3 * my_e2e / e2e_steps
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
        dist.all_reduce(per_rank, op=dist.ReduceOp.SUM)
    e2e_value = world * B * e2e_steps / float(e2e_s.item())
    h2d = 4 * (x0.size + h_U0.numel() + goal.size)
    d2h = sum(t.numel() * t.element_size() for t in h_res if t is not None)

    peaks = measured_peaks()
    extra = {}
    if not ilqr and args.workload == "C2" and not args.no_configs:
        extra["configs"] = other_configs(args, local, world, rank, flush, peaks, dist)
        extra["critic"] = critic_dp(args, local, world, rank, dist)
        if world == 1 and rank == 0:
            extra["configs"]["C1"]["acting_latency_get_optimal_action"] = acting_latency(local)

    if rank == 0:
        value = world * B * args.steps / total_s
        ms_per_step = 1e3 * total_s / args.steps
        fl = flops_per_state(cfg) * B                    # algorithmic FLOPs of one launch (one GPU)
        if ilqr:
            # data-dependent work, counted by the kernel: per 32-lane tile, every linearisation is T (1 + n)
            # dynamics passes + (1 + fout) cost passes, every rollout T dynamics passes + 1 cost pass;
            # a pass is 2 FLOP x MACs x the trajectories of the tile (algorithmic: B / tiles on average)
            dd = synthetic.dyn_dims(cfg["n"], cfg["m"], cfg["dyn_layers"], cfg["dyn_hidden"])
            cd = synthetic.cost_dims(cfg["n"], cfg["cost_layers"], cfg["cost_hidden"], cfg["cost_fout"])
            m_dyn = sum(a * b for a, b in zip(dd[:-1], dd[1:]))
            m_cost = sum(a * b for a, b in zip(cd[:-1], cd[1:]))
            sms = torch.cuda.get_device_properties(dev).multi_processor_count
            tile_traj = max(1, -(-B // sms))       # the host's tile rule (gmpc_ilqr): small batches get small tiles
            tile_traj = 32 if tile_traj > 16 else tile_traj
            tiles = -(-B // tile_traj)
            n_lin = ilqr_work[0] / args.steps + tiles
            n_roll = ilqr_work[1] / args.steps
            lanes = B / tiles
            fl = 2.0 * lanes * (n_lin * (cfg["T"] * (1 + cfg["n"]) * m_dyn + (1 + cfg["cost_fout"]) * m_cost)
                                + n_roll * (cfg["T"] * m_dyn + m_cost))
        kernel_s = (sum(step_ms) / len(step_ms)) * 1e-3  # the plan kernel IS the step (K=1: +1 memset)
        achieved = fl / kernel_s / 1e12
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(f"{args.workload}:{timed_path}")
        fp32_peak = _lib.measure_fp32_peak(local)
        f16_peak = _lib.measure_f16_mma_peak(local)
        line = {
            "metric": METRIC_ILQR if ilqr else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, args.lr),
            "path": timed_path, "clamped_ranks": clamped_ranks,
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                         "traffic": traffic,
                         "peak_source": f"{peaks['source']} cuBLAS bf16 burst (MEASURED_PEAKS.json)",
                         "ceiling": 1.0 / 3.0,
                         "ceiling_note": "fp32-accurate contraction on the kind::f16 pipe = three products of the "
                                         "fp16 hi/lo split (ah Wh + al Wh + ah Wl): at most a third of the peak",
                         "frac_of_ceiling": 3.0 * achieved / peaks["bf16_tflops"],
                         "f16_mma_peak_tflops": f16_peak,
                         "frac_of_f16_mma_peak": achieved / f16_peak if f16_peak else None,
                         "flops_per_launch": fl,
                         "fp32_ffma_peak_tflops": fp32_peak,
                         "frac_of_fp32_ffma_peak": achieved / fp32_peak if fp32_peak else None,
                         "hbm_stream_gbs": bytes_per_state(cfg) * B / kernel_s / 1e9,
                         "hbm_frac": bytes_per_state(cfg) * B / kernel_s / 1e9 / peaks["hbm_gbs"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "path": e2e_path,
                    "per_rank_ms_per_step": [float(v) for v in per_rank.tolist()]},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
            "step_ms_min_med_max": [min(step_ms), float(np.median(step_ms)), max(step_ms)],
        }
        if clamped_ranks:
            line["value"] = None
            line["rejected"] = (f"{clamped_ranks} rank(s) clamped an fp16 operand inside the timed loop: the timed plans "
                                "are outside the parity contract, no value is reported")
        if shard_parity is not None:
            line["shard_parity"] = shard_parity
        if parity is not None:
            line["parity"] = parity
        line.update(extra)
        if ilqr:  # fp32 CUDA-core kernel: the roofline that bounds it is the measured FFMA peak
            line["roofline"].update(bound="fp32-ffma (not one of the contract's hbm|tensor: this optional line is "
                                          "outside the BASELINE metric)", peak=fp32_peak,
                                    frac=achieved / fp32_peak if fp32_peak else None,
                                    peak_source="measured FP32 FFMA peak (gmpc_measure_fp32_peak)", traffic=None)
            line["ilqr_work_per_step"] = {"tile_outer_iterations": ilqr_work[0] / args.steps,
                                          "tile_rollouts": ilqr_work[1] / args.steps}
        if world == 1 and not args.no_cpu_baseline:
            if ilqr:
                cb, _, _ = time_cpu_ilqr(cfg, params, x0, U0, goal, target_s=args.ref_seconds)
            else:
                cb, _, _ = time_cpu_port(cfg, params, x0, U0, goal, args.lr, target_s=args.ref_seconds)
            line["cpu_baseline"] = cb
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
