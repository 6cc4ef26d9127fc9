"""L2 state-tracking policy (reference norm/l2_policy.py:11-18)."""

import torch

from gan_mpc_b200.policy import base


class L2MPC(base.BaseMPC):
    _loss_is_l2 = True

    def loss(self, xcseq, useq, params, desired_xseq):
        """sum_j mean_t (X[t,j] - desired[t,j])^2; carry columns sliced off.  Unbatched
        ([T+1,n]) or batched ([B,T+1,n])."""
        del useq, params
        n = desired_xseq.shape[-1]
        batched = xcseq.dim() == 3
        X = (xcseq if batched else xcseq[None])[..., :n].to(self.device, torch.float32).contiguous()
        D = (desired_xseq if batched else desired_xseq[None]).to(self.device, torch.float32).contiguous()
        h = self._any_handle(n)
        out = h.l2_loss(X, D)
        return out if batched else out[0]

    def _any_handle(self, n):
        for (hn, _), h in self._handles.items():
            if hn == n:
                return h
        raise RuntimeError("L2MPC.loss: plan once (or call policy._handle(n, m)) before computing losses")
