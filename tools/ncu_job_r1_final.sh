# ncu captures of the round-1 fp32 kernels (one GPU; each program has already run clean without ncu)
cat > /tmp/dynfit_run.py <<'PY'
import sys, numpy as np, torch
sys.path.insert(0, '.')
from gan_mpc_b200 import synthetic
from tests import util
cfg = dict(synthetic.CONFIGS["C2"], K=1)
p = synthetic.planner_params(0, **cfg)
h = util.make_handle(cfg, p)
B, S = 4096, 8
g = torch.Generator(device="cuda").manual_seed(0)
xs = torch.randn(B, S, cfg["n"], device="cuda", generator=g); us = torch.randn(B, S, cfg["m"], device="cuda", generator=g)
dims = [p["dyn_W"][0].shape[0]] + [w.shape[1] for w in p["dyn_W"]]
for _ in range(2):
    h.dynamics_fit(xs, us, xs + 0.1, 0.9, False, dims)
torch.cuda.synchronize()
PY
python /tmp/dynfit_run.py && ncu --set full --clock-control none --import-source on -k regex:dynfit_kernel -c 1 -o gpurun_out/prof_dynfit_v1 -f python /tmp/dynfit_run.py > gpurun_out/ncu_f_dynfit.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ilqr_kernel -c 1 -o gpurun_out/prof_ilqr_v2 -f python tools/ilqr_bench.py --config C2 --B 4096 --maxiter 2 --cpu-states 0 --reps 0 > gpurun_out/ncu_f_ilqr2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_benchpy_ilqr.csv python bench.py --planner ilqr --workload C1 --batch 4096 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_l_benchpy_ilqr.log 2>&1
tail -2 gpurun_out/ncu_f_dynfit.log gpurun_out/ncu_f_ilqr2.log; tail -4 gpurun_out/launches_benchpy_ilqr.csv | cut -c1-200
