"""EvalMPC with the reference's interface (policy/eval.py:25-128) over libgmpc."""

import torch

from gan_mpc_b200 import _lib
from gan_mpc_b200.dynamics.nn import dense_stack_lists
from gan_mpc_b200.policy import optimizers as opt

# policy/eval.py:10-20 -- honoured by planner method "ilqr" (gmpc_ilqr), the default: the policy then plans
# exactly like the reference (trajax iLQR), which is also what BaseMPC's bilevel gradient assumes (a
# stationary U*).  The first-order planner of the north star (methods "adam" / "grad") is an explicit opt-in
# (planner_kwargs / YAML mpc.planner.method); it ignores these options and says so once.
TRAJAX_iLQR_KWARGS = {
    "maxiter": 100, "grad_norm_threshold": 1e-4, "relative_grad_norm_threshold": 0.0,
    "obj_step_threshold": 0.0, "inputs_step_threshold": 0.0, "make_psd": False, "psd_delta": 0.0,
    "alpha_0": 1.0, "alpha_min": 0.00005,
}
PLANNER_KWARGS = {"method": "ilqr", "iters": 20, "learning_rate": 1e-2, "num_candidates": 1,
                  "b1": 0.9, "b2": 0.999, "eps": 1e-8, "path": "auto", "return_gradient": True}
COST_ARGS_NAME = ("goal_state",)


def _leaves(tree):
    if isinstance(tree, dict):
        for v in tree.values():
            yield from _leaves(v)
    elif isinstance(tree, torch.Tensor):
        yield tree


class EvalMPC:
    _loss_is_l2 = False

    def __init__(self, config, cost_model, dynamics_model, expert_model,
                 trajax_ilqr_kwargs=TRAJAX_iLQR_KWARGS, planner_kwargs=None, device=None):
        self.config = config
        self.cost_model = cost_model
        self.dynamics_model = dynamics_model
        self.expert_model = expert_model
        self.trajax_ilqr_kwargs = trajax_ilqr_kwargs
        pk = dict(PLANNER_KWARGS)
        cfg_pk = getattr(config.mpc, "planner", None)
        if cfg_pk is not None:
            pk.update(cfg_pk.to_dict())
        pk.update(planner_kwargs or {})
        self.planner_kwargs = pk
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        if hasattr(expert_model, "bind"):
            expert_model.bind(self)     # ExpertModel runs on this policy's libgmpc handles
        self.critic_model = None
        self._handles = {}
        self._staged = {}
        self.solver = self.create_mpc_solver()

    # ------------------------------------------------------------------ kernel plumbing
    def _handle(self, n, m):
        key = (n, m)
        if key not in self._handles:
            dyn, cost = self.dynamics_model.model, self.cost_model.model
            kw = {}
            if self.critic_model is not None:
                c = self.critic_model.model
                kw = dict(critic_features=c.lstm_features, critic_layers=c.num_layers,
                          critic_hidden=c.num_hidden_units)
            h = _lib.Handle(n, m, self.config.mpc.horizon, dyn.num_layers, dyn.num_hidden_units,
                            cost.num_layers, cost.num_hidden_units, cost.fout,
                            device=self.device.index, **kw)
            h.set_path(self.planner_kwargs["path"])
            self._handles[key] = h
        return self._handles[key]

    def _stage(self, h, params):
        """(re)stage the weights when any of the tensors was replaced or modified in place.

        The staged tensors are kept alive and compared by identity (`is`) plus their in-place version
        counter: a freshly built pytree (another checkpoint, a Polyak average, a tree_map result) can never
        be mistaken for the staged one, even when the caching allocator hands it the same addresses."""
        dW, db = dense_stack_lists(params["dynamics_params"])
        cW, cb = dense_stack_lists(params["cost_params"])
        ts = dW + db + cW + cb + [params["mpc_weights"]]
        prev = self._staged.get(id(h))
        same = (prev is not None and len(prev) == len(ts)
                and all(a is t and v == t._version for (a, v), t in zip(prev, ts)))
        if not same:
            h.set_weights([t.contiguous() for t in dW], [t.contiguous() for t in db],
                          [t.contiguous() for t in cW], [t.contiguous() for t in cb],
                          params["mpc_weights"].contiguous())
            self._staged[id(h)] = [(t, t._version) for t in ts]

    def _prep(self, x0, U, goal):
        """normalise (unbatched | batched) inputs to x0 [B,n], U [B,K,T,m], goal [B,T+1,n]."""
        batched = x0.dim() == 2
        x0b = x0 if batched else x0[None]
        if U.dim() == 2:
            Ub = U[None, None]
        elif U.dim() == 3:
            Ub = U[:, None]
        else:
            Ub = U
        gb = goal if goal.dim() == 3 else goal[None]
        f = lambda t: t.to(self.device, torch.float32).contiguous()
        return batched, f(x0b), f(Ub), f(gb)

    def _plan(self, x0, U, params, goal):
        batched, x0b, Ub, gb = self._prep(x0, U, goal)
        h = self._handle(x0b.shape[1], Ub.shape[3])
        self._stage(h, params)
        pk = self.planner_kwargs
        if pk["method"] == "ilqr":
            # the reference's own step: trajax iLQR with TRAJAX_iLQR_KWARGS (policy/optimizers.py:19-21)
            if Ub.shape[1] != 1:
                raise ValueError("method 'ilqr' plans one action sequence per state (no candidates)")
            out = h.ilqr(x0b, Ub[:, 0].contiguous(), gb, want_lqr=pk.get("return_lqr", False),
                         **self.trajax_ilqr_kwargs)
            self.last_plan_info = {"idx": None, "J_all": None, "path": "ilqr"}
            if not batched:
                out = tuple(None if o is None else (tuple(a[0] for a in o) if isinstance(o, tuple) else o[0])
                            for o in out)
            return out
        if not getattr(self, "_warned_first_order", False):
            import warnings
            warnings.warn(f"planner method {pk['method']!r}: the first-order planner of the north star, not the "
                          "reference's trajax iLQR (trajax_ilqr_kwargs are ignored; loss_and_grad still "
                          "differentiates through iLQR plans)", stacklevel=3)
            self._warned_first_order = True
        # Handle.plan reads the fp16 operand-range counter after the call and re-plans on the fp32 kernel (path
        # 'auto') or raises (forced tensor-core path): a clamped plan never leaves this function
        Ubest, X, J, idx, J_all = h.plan(x0b, Ub, gb, method=pk["method"], iters=pk["iters"],
                                         lr=pk["learning_rate"], b1=pk["b1"], b2=pk["b2"],
                                         eps=pk["eps"], check_range=True)
        grad = lam = None
        if pk["return_gradient"]:
            _, grad, _, lam = h.objective_grad(x0b, Ubest, gb, want_X=False, want_lam=True)
        self.last_plan_info = {"idx": idx, "J_all": J_all, "path": h.last_path}
        it = torch.full((x0b.shape[0],), pk["iters"], dtype=torch.int32, device=self.device)
        out = (X, Ubest, J, grad, lam, None, it)
        if not batched:
            out = tuple(None if o is None else o[0] for o in out)
        return out

    def _objective(self, x0, U, params, goal, grad=True):
        batched, x0b, Ub, gb = self._prep(x0, U, goal)
        h = self._handle(x0b.shape[1], Ub.shape[3])
        self._stage(h, params)
        J, dU, X, lam = h.objective_grad(x0b, Ub[:, 0].contiguous(), gb, want_grad=grad,
                                         want_lam=grad)
        out = (J, dU, X, lam)
        return out if batched else tuple(None if o is None else o[0] for o in out)

    def _rollout(self, x0, U, params):
        batched = x0.dim() == 2
        f = lambda t: t.to(self.device, torch.float32).contiguous()
        x0b, Ub = f(x0 if batched else x0[None]), f(U if batched else U[None])
        h = self._handle(x0b.shape[1], Ub.shape[2])
        self._stage(h, params)
        X = h.rollout(x0b, Ub)
        return X if batched else X[0]

    def _l2_loss_grad(self, x0, U, params, desired):
        batched, x0b, Ub, db = self._prep(x0, U, desired)
        h = self._handle(x0b.shape[1], Ub.shape[3])
        self._stage(h, params)
        out = h.l2_loss_grad(x0b, Ub[:, 0].contiguous(), db)
        return out if batched else tuple(o[0] for o in out)

    # ------------------------------------------------------------------ reference interface
    def create_mpc_solver(self):
        def func(xc, useq, params, cost_args, dynamics_args):
            return opt.ilqr_solve(self.cost, self.dynamics, xc, useq, params, cost_args,
                                  dynamics_args, self.trajax_ilqr_kwargs)
        return func

    def init(self, mpc_weights, cost_args, dynamics_args, expert_args):
        params = {}
        params["mpc_weights"] = torch.tensor(mpc_weights, dtype=torch.float32, device=self.device)
        params["cost_params"] = self.cost_model.init(*cost_args, device=self.device)
        params["dynamics_params"] = self.dynamics_model.init(*dynamics_args, device=self.device)
        params["expert_params"] = (self.expert_model.init(*expert_args)
                                   if self.expert_model is not None else {})
        return params

    def cost(self, xc, u, t, params, *args):
        """structured closure (policy/eval.py:64-69): consumed by policy.optimizers, not callable."""
        return self.cost_model.get_cost(xc, u, t, params["cost_params"], params["mpc_weights"], *args)

    def dynamics(self, xc, u, t, params, *args):
        """structured closure (policy/eval.py:71-73)."""
        return self.dynamics_model.predict(xc, u, t, params["dynamics_params"], *args)

    def get_dynamics_carry(self, history_x, history_u, params):
        return self.dynamics_model.get_history_carry(history_x[..., :-1, :], history_u,
                                                     params["dynamics_params"])

    def get_goal_states_init_actions(self, histroy_x, params):
        """policy/eval.py:87-107 -- the expert proposes goal_xseq [T+1,n] and init_useq [T,m]."""
        if self.expert_model is None:
            raise ValueError("no expert model: pass goal states / initial actions explicitly to "
                             "policy.solver(xc, useq, params, (goal_xseq,), ())")
        expert_params = params["expert_params"]
        if hasattr(self.expert_model, "propose"):   # ExpertModel: history carry + proposal, one launch
            return self.expert_model.propose(self, histroy_x, expert_params)
        x = histroy_x[..., -1, :]
        T = self.config.mpc.horizon
        xseq = torch.cat([x[..., None, :], torch.zeros(*x.shape[:-1], T - 1, x.shape[-1],
                                                       device=x.device, dtype=x.dtype)], dim=-2)
        carry = self.expert_model.get_history_carry(histroy_x, xseq, expert_params)
        _, (goal_xseq, init_useq) = self.expert_model.get_carry_next_state_and_action_seq(
            carry, xseq, expert_params)
        return goal_xseq, init_useq

    def get_optimal_values(self, params, history_x, history_u):
        """policy/eval.py:109-124.  history_x [h+1,n] (or [B,h+1,n])."""
        goal_xseq, init_useq = self.get_goal_states_init_actions(history_x, params)
        init_carry = self.get_dynamics_carry(history_x, history_u, params)
        x = history_x[..., -1, :]
        xc = torch.cat([x, init_carry.expand(*x.shape[:-1], init_carry.shape[-1])], dim=-1)
        return self.solver(xc, init_useq, params, (goal_xseq,), ())

    def get_optimal_action(self, params, history_x, history_u):
        _, useq, *_ = self.get_optimal_values(params, history_x, history_u)
        return useq[..., 0, :]
