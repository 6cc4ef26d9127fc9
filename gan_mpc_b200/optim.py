"""optax.multi_transform({"tx": chain(clip_by_global_norm(100), adam(lr)), "zero": set_to_zero()})
as used by the reference (norm/runner.py:46-58, gan/runner.py:51-63), over flat device vectors;
the arithmetic is libgmpc's gmpc_clip_adam_step."""

import torch


class MaskedClipAdam:
    def __init__(self, lr, labels, max_norm=100.0, b1=0.9, b2=0.999, eps=1e-8):
        self.lr, self.labels, self.max_norm = lr, dict(labels), max_norm
        self.b1, self.b2, self.eps = b1, b2, eps

    @property
    def trained(self):
        return [k for k, v in self.labels.items() if v == "tx"]

    def init(self, params):
        """opt_state: update count + first/second moments per trained top-level entry (allocated
        on first use, in the flat layout of the kernel that trains that entry)."""
        return {"count": 0, "mu": {}, "nu": {}}

    def moments(self, opt_state, key, like):
        if key not in opt_state["mu"]:
            opt_state["mu"][key] = torch.zeros_like(like)
            opt_state["nu"][key] = torch.zeros_like(like)
        return opt_state["mu"][key], opt_state["nu"][key]

    def step_flat(self, handle, opt_state, key, flat_params, flat_grad, grad_scale=1.0):
        """one clipped-Adam update of the flat vector `flat_params` in place (count was already
        advanced by the caller for this update)."""
        mu, nu = self.moments(opt_state, key, flat_params)
        handle.clip_adam_step(flat_params, flat_grad, mu, nu, step=opt_state["count"], lr=self.lr,
                              max_norm=self.max_norm, grad_scale=grad_scale, b1=self.b1,
                              b2=self.b2, eps=self.eps)
