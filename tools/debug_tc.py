import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import util
from oracle import planner as oracle
def dev(a): return torch.from_numpy(np.ascontiguousarray(a)).cuda()
cfg = dict(util.MID, T=8)
for B in (64, 65, 70, 96, 97, 200):
    p, x0, U0, goal = util.case(cfg, 21, B=B)
    op = util.to_oracle(p)
    oX, oJ, odU, olam = oracle.objective_grad(util.tt(x0), util.tt(U0[:, 0]), util.tt(goal), op)
    h = util.make_handle(cfg, p); h.set_path("tc")
    for rep in range(2):
        J, dU, X, lam = h.objective_grad(dev(x0), dev(U0[:, 0]), dev(goal), want_lam=True)
        el = (lam.double().cpu() - olam).abs()         # [B,T+1,n]
        bad = (el.amax(dim=(1, 2)) > 1e-3).nonzero().flatten().tolist()
        print(f"B={B} rep{rep} lam err per t:", [f"{v:.0e}" for v in el.amax(dim=(0, 2)).tolist()], "bad rows:", bad[:40], len(bad))
