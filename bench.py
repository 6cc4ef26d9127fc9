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
                          "data": "synthetic", "config": workload_config(args, cfg, lr, "cpu-oracle-port"),
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
            "config": workload_config(args, cfg, lr, "cpu-oracle-port"),
            "cpu_baseline": cb,
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def workload_config(args, cfg, lr, path):
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
            "path": path, "parallelism": f"dp{args.gpus} (start states sharded, no data-path collective; "
                                         "best plans all-gathered over NCCL)" if args.gpus > 1 else "single GPU"}


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
            h.plan(d_x0, d_U0, d_goal, method="adam", iters=cfg["iters"], lr=args.lr, out=out)
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
    step_ms = [e0.elapsed_time(e1) for e0, e1 in evs]
    total_ms = torch.tensor([sum(step_ms)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_s = float(total_ms.item()) * 1e-3
    clocks = sampler.stop() if rank == 0 else None

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
    barrier()
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * B * e2e_steps / float(e2e_s.item())
    h2d = 4 * (x0.size + h_U0.numel() + goal.size)
    d2h = sum(t.numel() * t.element_size() for t in h_res if t is not None)

    if rank == 0:
        peaks = measured_peaks()
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
            traffic = json.load(open(tpath)).get(f"{args.workload}:{h.last_path}")
        fp32_peak = _lib.measure_fp32_peak(local)
        line = {
            "metric": METRIC_ILQR if ilqr else METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, cfg, args.lr, h.last_path),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"],
                         "unit": "TFLOP/s", "frac": achieved / peaks["bf16_tflops"],
                         "traffic": traffic,
                         "peak_source": f"{peaks['source']} cuBLAS bf16 burst (MEASURED_PEAKS.json)",
                         "flops_per_launch": fl,
                         "fp32_ffma_peak_tflops": fp32_peak,
                         "frac_of_fp32_ffma_peak": achieved / fp32_peak if fp32_peak else None,
                         "hbm_stream_gbs": bytes_per_state(cfg) * B / kernel_s / 1e9,
                         "hbm_frac": bytes_per_state(cfg) * B / kernel_s / 1e9 / peaks["hbm_gbs"]},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": clocks, "wall_s_timed_region": wall,
            "step_ms_min_med_max": [min(step_ms), float(np.median(step_ms)), max(step_ms)],
        }
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
