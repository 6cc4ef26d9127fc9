"""Jensen-Shannon GAN policy (reference gan/js_policy.py:11-74): critic BCE loss / gradient and
generator loss on the LSTM discriminator, evaluated by libgmpc."""

import torch

from gan_mpc_b200.policy import base, eval


def _zeros_like_tree(tree):
    if isinstance(tree, dict):
        return {k: _zeros_like_tree(v) for k, v in tree.items()}
    return torch.zeros_like(tree) if isinstance(tree, torch.Tensor) else tree


class JS_MPC(base.BaseMPC):
    def __init__(self, config, cost_model, dynamics_model, expert_model, critic_model,
                 loss_vmap=(0,), trajax_ilqr_kwargs=eval.TRAJAX_iLQR_KWARGS, planner_kwargs=None,
                 device=None):
        super().__init__(config, cost_model, dynamics_model, expert_model, loss_vmap,
                         trajax_ilqr_kwargs, planner_kwargs, device)
        self.critic_model = critic_model

    def init(self, mpc_weights, cost_args, dynamics_args, expert_args, critic_args):
        params = super().init(mpc_weights, cost_args, dynamics_args, expert_args)
        params["critic_params"] = self.critic_model.init(*critic_args, device=self.device)
        return params

    # ------------------------------------------------------------------ flat plumbing
    def critic_handle(self, n):
        for (hn, _), h in self._handles.items():
            if hn == n:
                return h
        return self._handle(n, 1)

    def critic_flat(self, params):
        return self.critic_model.model.flatten(params["critic_params"])

    def _xl(self, xseq, label=None):
        batched = xseq.dim() == 3
        x = (xseq if batched else xseq[None]).to(self.device, torch.float32).contiguous()
        if label is None:
            return batched, x, None
        lab = torch.as_tensor(label, dtype=torch.float32, device=self.device).reshape(-1).contiguous()
        return batched, x, lab

    # ------------------------------------------------------------------ reference interface
    def critic_logits(self, xseq, params):
        """CriticModel.predict (critic/critic_model.py:15-16) for one or many trajectories."""
        batched, x, _ = self._xl(xseq)
        s = self.critic_handle(x.shape[-1]).critic_forward(x, self.critic_flat(params))
        return s if batched else s[0]

    def critic_loss(self, xseq, label, params):
        """gan/js_policy.py:41-46: -log(where(label > 0, p, 1 - p)), p = sigmoid(score).  For a
        batch the mean over the batch is returned (the reference vmaps and takes the mean)."""
        _, x, lab = self._xl(xseq, label)
        loss, _ = self.critic_handle(x.shape[-1]).critic_loss_grad(x, lab, self.critic_flat(params),
                                                                   want_grad=False)
        return loss[0]

    def critic_loss_and_grad(self, batch_xseq, batch_label, params):
        """gan/js_policy.py:48-58: value_and_grad of the batch-mean loss w.r.t. the WHOLE params
        pytree (non-critic leaves get zeros, as in the reference)."""
        _, x, lab = self._xl(batch_xseq, batch_label)
        loss, g = self.critic_handle(x.shape[-1]).critic_loss_grad(x, lab, self.critic_flat(params))
        grads = _zeros_like_tree({k: v for k, v in params.items() if k != "critic_params"})
        grads["critic_params"] = self.critic_model.model.unflatten(g, x.shape[-1])
        return loss[0], grads

    def generator_loss(self, xcseq, useq, params, actual_xseq):
        """gan/js_policy.py:60-68: mean(-log p + log(1 - p)) = -score (stable form; the reference's
        log form saturates to +-inf in fp32 for |score| > ~17 and to -score elsewhere)."""
        del useq
        n = actual_xseq.shape[-1]
        return -self.critic_logits(xcseq[..., :n], params)

    def _loss_state_grad(self, X, params, desired):
        """generator loss per trajectory and its gradient w.r.t. the planned states (the seed of
        loss_grad_wrt_control, policy/optimizers.py:78-83): loss = -score, dL/dX = -d score/dX."""
        n = desired.shape[-1]
        xs = X[..., :n].contiguous()
        score, dx = self.critic_handle(n).critic_input_grad(xs, self.critic_flat(params))
        return -score, -dx

    def generator_loss_and_grad(self, batch_xseq, params, batch_loss_args):
        return self.loss_and_grad(batch_xseq, params, batch_loss_args)

    def loss(self, xcseq, useq, params, desired_xseq):
        return self.generator_loss(xcseq, useq, params, desired_xseq)
