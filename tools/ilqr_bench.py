"""iLQR mode (gmpc_ilqr, the reference's own planner step) on BASELINE config dims: planned states/s,
iterations, rollouts, and the CPU oracle port beside it on a bounded sample.
    python tools/ilqr_bench.py [--config C2] [--B 4096] [--maxiter 100] [--T 32] [--cpu-states 16]"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gan_mpc_b200 import _lib, synthetic  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--config", default="C2")
    ap.add_argument("--B", type=int, default=None)
    ap.add_argument("--T", type=int, default=None)
    ap.add_argument("--maxiter", type=int, default=100)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--cpu-states", type=int, default=16)
    ap.add_argument("--bilevel", action="store_true",
                    help="time gmpc_bilevel_l2 (iLQR + the bilevel tail, policy/optimizers.py:34-75) instead")
    a = ap.parse_args()
    cfg = dict(synthetic.CONFIGS[a.config], K=1)
    if a.B:
        cfg["B"] = a.B
    if a.T:
        cfg["T"] = a.T
    p = synthetic.planner_params(0, **cfg)
    x0, U0, goal = synthetic.planner_inputs(0, **cfg)
    dev = torch.device("cuda", 0)
    t = lambda x: torch.from_numpy(np.ascontiguousarray(x)).to(dev)
    h = _lib.Handle(cfg["n"], cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"], cfg["cost_layers"],
                    cfg["cost_hidden"], cfg["cost_fout"], device=0)
    h.set_weights([t(w) for w in p["dyn_W"]], [t(b) for b in p["dyn_b"]], [t(w) for w in p["cost_W"]],
                  [t(b) for b in p["cost_b"]], t(p["mpc_weights"]))
    dx0, dU0, dgoal = t(x0), t(U0[:, 0]), t(goal)
    ddes = (dgoal + 0.05 * torch.randn(dgoal.shape, device=dev, generator=torch.Generator(dev).manual_seed(0))).contiguous()
    J0, *_ = h.objective_grad(dx0, dU0, dgoal, want_grad=False, want_X=False)
    times = []
    for rep in range(a.reps + 1):
        h.ilqr_stats()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        if a.bilevel:
            o = h.bilevel_l2(dx0, dU0, dgoal, ddes, maxiter=a.maxiter)
            obj, it = o["obj"], o["iteration"]
        else:
            X, U, obj, g, lam, _, it = h.ilqr(dx0, dU0, dgoal, maxiter=a.maxiter)
        e1.record()
        torch.cuda.synchronize()
        outer, rolls = h.ilqr_stats()
        if rep:
            times.append(e0.elapsed_time(e1))
    ms = float(np.median(times))
    B = cfg["B"]
    out = dict(metric=("bilevel gradients/sec (gmpc_bilevel_l2: iLQR + Hessian + solve + tangent)" if a.bilevel
                       else "planned states/sec (trajax-iLQR mode, gmpc_ilqr)"), value=B / ms * 1e3, unit="states/s",
               ms=ms, config=dict(workload=f"{a.config} dims: B={B}, n={cfg['n']}, m={cfg['m']}, T={cfg['T']}, "
                                  f"maxiter={a.maxiter}, grad_norm_threshold=1e-4, alpha_min=5e-5"),
               iterations=dict(median=float(it.float().median()), max=int(it.max()), min=int(it.min())),
               tile_outer_iterations=outer, tile_rollouts=rolls,
               obj_over_initial_median=float((obj / J0).median()))
    if a.cpu_states > 0 and not a.bilevel:
        # (CPU leg kept for interactive use; the contract-conformant CPU baseline of the iLQR mode is
        # `python bench.py --planner ilqr [--impl reference]`)
        from oracle import ilqr as oilqr
        nb = min(a.cpu_states, B)
        torch.set_num_threads(len(os.sched_getaffinity(0)))
        d = lambda x: torch.from_numpy(x).float()
        op = {k: ([d(w) for w in v] if isinstance(v, list) else d(v)) for k, v in p.items()}
        t0 = time.time()
        oout = oilqr.ilqr(d(x0[:nb]), d(U0[:nb, 0]), d(goal[:nb]), op, maxiter=a.maxiter)
        dt = time.time() - t0
        out["cpu_baseline"] = dict(value=nb / dt, unit="states/s", cores=len(os.sched_getaffinity(0)), kind="port",
                                   sample=f"{nb} states, fp32 torch-CPU oracle port of trajax iLQR, {dt:.1f} s",
                                   obj_over_initial_median=float((oout[2] / J0[:nb].cpu()).median()))
    print(json.dumps(out))


if __name__ == "__main__":
    main()
