"""BaseMPC with the reference's interface (policy/base.py:12-128)."""

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

    def loss_and_grad(self, history_X, params, batch_loss_args):
        """policy/base.py:87-128 -- vmap of the bilevel gradient: next scope row."""
        from gan_mpc_b200.policy import optimizers as opt
        opt.bilevel_optimization()
