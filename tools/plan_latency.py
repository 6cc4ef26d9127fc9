import sys, numpy as np, torch
sys.path.insert(0, '/root/repo')
from gan_mpc_b200 import _lib, synthetic
for name in ("C1", "C2"):
    for B in (1, 32, 64):
        cfg = dict(synthetic.CONFIGS[name], B=B, K=1)
        p = synthetic.planner_params(0, **cfg); x0, U0, goal = synthetic.planner_inputs(0, **cfg)
        g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
        h = _lib.Handle(cfg["n"], cfg["m"], cfg["T"], cfg["dyn_layers"], cfg["dyn_hidden"], cfg["cost_layers"], cfg["cost_hidden"], cfg["cost_fout"])
        h.set_weights([g(w) for w in p["dyn_W"]], [g(b) for b in p["dyn_b"]], [g(w) for w in p["cost_W"]], [g(b) for b in p["cost_b"]], g(p["mpc_weights"]))
        res = {}
        for path in ("auto", "ffma", "tc16"):
            try:
                h.set_path(path)
            except Exception as e:
                res[path] = "n/a"; continue
            ts = []
            for i in range(6):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); h.plan(g(x0), g(U0), g(goal), iters=20); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            res[path] = (round(float(np.median(ts[1:])), 2), h.last_path)
        print(name, "B", B, res)
