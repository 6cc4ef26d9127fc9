"""BaseMPC with the reference's interface (policy/base.py:12-128)."""

import torch

from gan_mpc_b200.policy import eval


class BaseMPC(eval.EvalMPC):
    def __init__(self, config, cost_model, dynamics_model, expert_model, loss_vmap=(0,),
                 trajax_ilqr_kwargs=eval.TRAJAX_iLQR_KWARGS, planner_kwargs=None, device=None):
        super().__init__(config=config, cost_model=cost_model, dynamics_model=dynamics_model,
                         expert_model=expert_model, trajax_ilqr_kwargs=trajax_ilqr_kwargs,
                         planner_kwargs=planner_kwargs, device=device)
        self.loss_vmap = loss_vmap

    def get_dynamics_carry(self, history_x, *args):
        """policy/base.py:31-38 -- ignores history, zero carry (empty for the MLP)."""
        return self.dynamics_model.get_zero_carry(history_x[..., :-1, :].reshape(-1, history_x.shape[-1]))

    def get_optimal_values(self, params, history_x, *args):
        return super().get_optimal_values(params, history_x, None)

    def get_optimal_action(self, params, history_x, *args):
        _, useq, *_ = self.get_optimal_values(params, history_x, *args)
        return useq[..., 0, :]

    def loss(self, xcseq, useq, params, *args):
        raise NotImplementedError

    def _bilevel(self, x0, init_U, params, goal, desired, ilqr_kwargs=None, **kw):
        """The bilevel solve on normalised inputs; returns (batched?, result dict).
        L2MPC: one launch (gmpc_bilevel_l2).  Other losses (JS_MPC.generator_loss) supply
        `_loss_state_grad(X, params, desired) -> (loss [B], dL/dX [B,T+1,n])` and run as
        gmpc_ilqr -> loss gradient -> gmpc_bilevel_tail at the planned U."""
        batched, x0b, Ub, gb = self._prep(x0, init_U, goal)
        if Ub.shape[1] != 1:
            raise ValueError("bilevel_optimization plans one action sequence per state")
        db = (desired if desired.dim() == 3 else desired[None]).to(self.device, torch.float32).contiguous()
        h = self._handle(x0b.shape[1], Ub.shape[3])
        self._stage(h, params)
        ik = dict(self.trajax_ilqr_kwargs if ilqr_kwargs is None else ilqr_kwargs)
        if self._loss_is_l2:
            return batched, h.bilevel_l2(x0b, Ub[:, 0].contiguous(), gb, db, **ik, **kw)
        if not hasattr(self, "_loss_state_grad"):
            raise NotImplementedError("bilevel gradient: this policy's loss has no state-gradient kernel")
        if kw:
            raise NotImplementedError("cost_hessian_wrt_control / cost_vjp go through an L2MPC policy")
        X, U, obj, low, _, _, it = h.ilqr(x0b, Ub[:, 0].contiguous(), gb, **ik)
        loss, dLdX = self._loss_state_grad(X, params, db)
        out = h.bilevel_tail(x0b, U, gb, dLdX)
        out.update(X=X, U=U, obj=obj, low_level_grad=low, iteration=it, loss=loss)
        return batched, out

    def loss_and_grad(self, history_X, params, batch_loss_args):
        """policy/base.py:87-128 -- vmap of bilevel_optimization over the batch, mean loss and
        leaf-wise mean of the per-sample gradient pytrees (here reduced without materialising them).
        history_X [B,h+1,n], batch_loss_args = (batch_y [B,T+1,n],).
        With torch.distributed initialised the batch is sharded over the ranks (contiguous blocks,
        parallel.shard_range) and the sums are all-reduced once: every rank returns the global mean."""
        from gan_mpc_b200 import parallel
        from gan_mpc_b200.policy import bilevel
        (batch_y,) = batch_loss_args
        B = history_X.shape[0]
        rank, world = parallel.rank_world()
        lo, hi = parallel.shard_range(B, rank, world)
        if hi > lo:
            hx, by = history_X[lo:hi], batch_y[lo:hi]
            goal, init_u = self.get_goal_states_init_actions(hx, params)
            _, out = self._bilevel(hx[..., -1, :], init_u, params, goal, by)
            self.last_bilevel = out
            loss_sum = out["loss"].sum()
            grads = bilevel.high_level_grad_tree(params, out, reduce_mean="sum",
                                                 handle=self._handle(hx.shape[-1], init_u.shape[-1]))
        else:  # more ranks than samples: this rank contributes zeros
            loss_sum = torch.zeros((), device=self.device)
            grads = bilevel.zeros_like_tree(params)
        return bilevel.allreduce_mean_tree_(loss_sum, grads, B)
