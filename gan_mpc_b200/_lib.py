"""ctypes binding of libgmpc.so (include/gmpc.h) -- the only way Python reaches the kernels.

PyTorch is used for device memory and streams only: every call passes `tensor.data_ptr()` and
`torch.cuda.current_stream().cuda_stream`.  There is no CPU fallback: a missing library or a
non-zero status raises.
"""

import ctypes as C
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GMPC_LIB_PATH", os.path.join(_HERE, "libgmpc.so"))  # override: experiments only

METHOD_GRAD, METHOD_ADAM = 0, 1
PATH_AUTO, PATH_FFMA, PATH_TC, PATH_TC16, PATH_TC16S, PATH_T128 = 0, 1, 2, 3, 4, 5
METHODS = {"grad": METHOD_GRAD, "adam": METHOD_ADAM}
PATHS = {"auto": PATH_AUTO, "ffma": PATH_FFMA, "tc16": PATH_TC16, "tc16s": PATH_TC16S, "t128": PATH_T128}


class GmpcError(RuntimeError):
    pass


class Config(C.Structure):
    _fields_ = [(k, C.c_int32) for k in (
        "n", "m", "T", "dyn_layers", "dyn_hidden", "cost_layers", "cost_hidden", "cost_fout",
        "critic_features", "critic_layers", "critic_hidden", "device")]


class IlqrOptions(C.Structure):
    """gmpc_ilqr_options: the trajax options the reference sets (policy/eval.py:10-20)."""
    _fields_ = [("maxiter", C.c_int32), ("grad_norm_threshold", C.c_float),
                ("alpha_0", C.c_float), ("alpha_min", C.c_float), ("gradient_lag", C.c_int32)]


_f = C.c_void_p  # device/host float*
_SIGS = {
    "gmpc_create": (C.c_int, [C.POINTER(Config), C.POINTER(C.c_void_p)]),
    "gmpc_destroy": (C.c_int, [C.c_void_p]),
    "gmpc_last_error": (C.c_char_p, []),
    "gmpc_critic_param_count": (C.c_int64, [C.c_void_p]),
    "gmpc_set_path": (C.c_int, [C.c_void_p, C.c_int]),
    "gmpc_last_path": (C.c_int, [C.c_void_p]),
    "gmpc_launch_count": (C.c_int64, [C.c_void_p]),
    "gmpc_range_overflow": (C.c_int, [C.c_void_p, C.POINTER(C.c_int32), C.c_void_p]),
    "gmpc_set_weights": (C.c_int, [C.c_void_p, C.POINTER(_f), C.POINTER(_f), C.POINTER(_f),
                                   C.POINTER(_f), _f, C.c_void_p]),
    "gmpc_rollout": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, C.c_void_p]),
    "gmpc_objective_grad": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, _f, _f, _f, _f,
                                      C.c_void_p]),
    "gmpc_l2_loss_grad": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, _f, _f, _f, C.c_void_p]),
    "gmpc_plan": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, C.c_int32, C.c_int32,
                            C.c_float, C.c_float, C.c_float, C.c_float, _f, _f, _f, _f, _f,
                            C.c_void_p]),
    "gmpc_plan_host": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, C.c_int32,
                                 C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, _f, _f,
                                 _f, _f, _f, C.c_void_p]),
    "gmpc_ilqr": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, C.POINTER(IlqrOptions), _f, _f, _f,
                            _f, _f, _f, _f, _f, C.c_void_p]),
    "gmpc_ilqr_host": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, C.POINTER(IlqrOptions)] + [_f] * 6
                       + [C.c_void_p]),
    "gmpc_bilevel_l2": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, _f, C.POINTER(IlqrOptions)] + [_f] * 12
                        + [C.c_void_p]),
    "gmpc_bilevel_tail": (C.c_int, [C.c_void_p, C.c_int64] + [_f] * 9 + [C.c_void_p]),
    "gmpc_critic_input_grad": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, _f, C.c_void_p]),
    "gmpc_ilqr_stats": (C.c_int, [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64),
                                  C.c_void_p]),
    "gmpc_dynamics_fit": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, C.c_float, C.c_int32, _f,
                                    C.POINTER(_f), C.POINTER(_f), C.c_void_p]),
    "gmpc_dynamics_fit_columns": (C.c_int64, [C.c_void_p, C.c_int64, C.c_int32]),
    "gmpc_gemm_nt": (C.c_int, [C.c_void_p, C.c_int32, C.c_int32, C.c_int64, _f, _f, C.c_float, C.c_int32, _f, _f,
                               C.c_void_p]),
    "gmpc_cost_mixed_vjp": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, C.c_float, C.POINTER(_f), C.POINTER(_f),
                                      C.c_void_p]),
    "gmpc_expert_propose": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, C.c_int32, C.c_int32,
                                      C.c_int32, _f, _f, C.c_void_p]),
    "gmpc_expert_param_count": (C.c_int64, [C.c_int32] * 5),
    "gmpc_critic_forward": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, C.c_void_p]),
    "gmpc_critic_loss_grad": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, C.c_float,
                                        _f, _f, C.c_void_p]),
    "gmpc_critic_loss_grad_gather": (C.c_int, [C.c_void_p, C.c_int64, C.c_int32, _f, _f, _f, _f,
                                               C.c_float, _f, _f, C.c_void_p]),
    "gmpc_clip_adam_step": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, _f, C.c_int32, C.c_float,
                                      C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                      C.c_void_p]),
    "gmpc_critic_train_scan": (C.c_int, [C.c_void_p, C.c_int32, C.c_int64, C.c_int32, _f, _f, _f, _f, _f, _f,
                                         C.c_int32, C.c_float, C.c_float, C.c_float, C.c_float, C.c_float,
                                         _f, _f, C.c_void_p]),
    "gmpc_l2_loss": (C.c_int, [C.c_void_p, C.c_int64, _f, _f, _f, C.c_void_p]),
    "gmpc_measure_fp32_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "gmpc_measure_f16_mma_peak": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def load():
    """Load libgmpc.so (built in-tree by __graft_entry__.build()).  Raises if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GmpcError(
                f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; "
                "g.build()'` (there is no CPU fallback)")
        lib = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _lib = lib
    return _lib


def _check(rc):
    if rc != 0:
        raise GmpcError(f"libgmpc status {rc}: {load().gmpc_last_error().decode()}")


def _ptr(t, dtype=torch.float32, device=None, name="tensor"):
    if t is None:
        return None
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name}: expected a torch.Tensor, got {type(t).__name__}")
    if t.dtype != dtype:
        raise TypeError(f"{name}: expected dtype {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise ValueError(f"{name}: must be contiguous")
    if device is not None and t.device != device:
        raise ValueError(f"{name}: expected device {device}, got {t.device}")
    return t.data_ptr()


def _stream(device):
    return torch.cuda.current_stream(device).cuda_stream


class Handle:
    """One planner/critic instance bound to one GPU (include/gmpc.h gmpc_handle)."""

    def __init__(self, n, m, T, dyn_layers, dyn_hidden, cost_layers, cost_hidden, cost_fout,
                 critic_features=0, critic_layers=1, critic_hidden=1, device=0):
        self.lib = load()
        self.device = torch.device("cuda", int(device))
        self.cfg = Config(n, m, T, dyn_layers, dyn_hidden, cost_layers, cost_hidden, cost_fout,
                          critic_features, critic_layers, critic_hidden, int(device))
        self._h = C.c_void_p()
        _check(self.lib.gmpc_create(C.byref(self.cfg), C.byref(self._h)))
        self.n, self.m, self.T = n, m, T
        self._path = PATH_AUTO

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self.lib.gmpc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---------------------------------------------------------------- configuration
    def set_path(self, path):
        code = PATHS[path] if isinstance(path, str) else int(path)
        _check(self.lib.gmpc_set_path(self._h, code))
        self._path = code

    @property
    def last_path(self):
        return {PATH_FFMA: "ffma", PATH_TC16: "tc16", PATH_TC16S: "tc16s", PATH_T128: "t128"}[self.lib.gmpc_last_path(self._h)]

    def range_overflow(self):
        """CTAs of the fp16-split kernel that clamped an operand since the last call (synchronises)."""
        cnt = C.c_int32(0)
        _check(self.lib.gmpc_range_overflow(self._h, C.byref(cnt), _stream(self.device)))
        return int(cnt.value)

    @property
    def launch_count(self):
        return int(self.lib.gmpc_launch_count(self._h))

    @property
    def critic_param_count(self):
        return int(self.lib.gmpc_critic_param_count(self._h))

    def set_weights(self, dyn_W, dyn_b, cost_W, cost_b, mpc_weights):
        dev = self.device
        if len(dyn_W) != self.cfg.dyn_layers or len(cost_W) != self.cfg.cost_layers:
            raise ValueError("set_weights: layer count does not match the handle's config")

        def arr(ts, nm):
            a = (_f * len(ts))()
            for i, t in enumerate(ts):
                a[i] = _ptr(t, device=dev, name=f"{nm}[{i}]")
            return a

        self._keep = (dyn_W, dyn_b, cost_W, cost_b, mpc_weights)
        _check(self.lib.gmpc_set_weights(
            self._h, arr(dyn_W, "dyn_W"), arr(dyn_b, "dyn_b"), arr(cost_W, "cost_W"),
            arr(cost_b, "cost_b"), _ptr(mpc_weights, device=dev, name="mpc_weights"),
            _stream(dev)))

    # ---------------------------------------------------------------- planner
    def rollout(self, x0, U):
        B = x0.shape[0]
        X = torch.empty(B, self.T + 1, self.n, device=self.device, dtype=torch.float32)
        _check(self.lib.gmpc_rollout(self._h, B, _ptr(x0, device=self.device, name="x0"),
                                     _ptr(U, device=self.device, name="U"), _ptr(X),
                                     _stream(self.device)))
        return X

    def objective_grad(self, x0, U, goal, want_grad=True, want_X=True, want_lam=False):
        B, dev = x0.shape[0], self.device
        J = torch.empty(B, device=dev, dtype=torch.float32)
        dU = torch.empty(B, self.T, self.m, device=dev, dtype=torch.float32) if want_grad else None
        X = torch.empty(B, self.T + 1, self.n, device=dev, dtype=torch.float32) if want_X else None
        lam = torch.empty(B, self.T + 1, self.n, device=dev, dtype=torch.float32) if want_lam else None
        _check(self.lib.gmpc_objective_grad(
            self._h, B, _ptr(x0, device=dev, name="x0"), _ptr(U, device=dev, name="U"),
            _ptr(goal, device=dev, name="goal"), _ptr(J), _ptr(dU), _ptr(X), _ptr(lam),
            _stream(dev)))
        return J, dU, X, lam

    def l2_loss_grad(self, x0, U, desired, want_grad=True, want_X=True):
        B, dev = x0.shape[0], self.device
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        dU = torch.empty(B, self.T, self.m, device=dev, dtype=torch.float32) if want_grad else None
        X = torch.empty(B, self.T + 1, self.n, device=dev, dtype=torch.float32) if want_X else None
        _check(self.lib.gmpc_l2_loss_grad(
            self._h, B, _ptr(x0, device=dev, name="x0"), _ptr(U, device=dev, name="U"),
            _ptr(desired, device=dev, name="desired"), _ptr(loss), _ptr(dU), _ptr(X),
            _stream(dev)))
        return loss, dU, X

    def plan(self, x0, U0, goal, method="adam", iters=20, lr=1e-2, b1=0.9, b2=0.999, eps=1e-8,
             want_J_all=True, out=None, check_range=True):
        """x0 [B,n], U0 [B,K,T,m], goal [B,T+1,n] -> (U_best, X_best, J_best, idx, J_all).

        gmpc_plan itself never synchronises.  With check_range (default) the fp16 operand-range counter of
        the tensor-core kernels is read after the call (one 4-byte D2H copy, synchronises the stream): if
        an operand was clamped the plan is outside the parity contract and is re-done on the fp32 CUDA-core
        kernel when the path is AUTO, or GmpcError is raised when a tensor-core path was forced.  A caller
        that keeps the stream asynchronous passes check_range=False and calls range_overflow() itself."""
        dev = self.device
        B, K = U0.shape[0], U0.shape[1]
        if out is None:
            out = self.alloc_plan_outputs(B, K, want_J_all)
        U_best, X_best, J_best, idx, J_all = out

        def launch():
            _check(self.lib.gmpc_plan(
                self._h, B, K, _ptr(x0, device=dev, name="x0"), _ptr(U0, device=dev, name="U0"),
                _ptr(goal, device=dev, name="goal"), METHODS[method], iters, lr, b1, b2, eps,
                _ptr(U_best), _ptr(X_best), _ptr(J_best), _ptr(idx, dtype=torch.int32), _ptr(J_all),
                _stream(dev)))

        launch()
        if check_range and B > 0 and self.last_path != "ffma" and self.range_overflow() > 0:
            if self._path != PATH_AUTO:
                raise GmpcError(f"plan: an operand left the fp16 hi/lo range on the forced path "
                                f"{self.last_path!r}; use path 'auto' or 'ffma'")
            self.set_path(PATH_FFMA)
            try:
                launch()
            finally:
                self.set_path(PATH_AUTO)
        return out

    def alloc_plan_outputs(self, B, K, want_J_all=True, device=None, pin=False):
        dev = self.device if device is None else device
        kw = dict(device=dev, pin_memory=pin) if pin else dict(device=dev)
        return (torch.empty(B, self.T, self.m, dtype=torch.float32, **kw),
                torch.empty(B, self.T + 1, self.n, dtype=torch.float32, **kw),
                torch.empty(B, dtype=torch.float32, **kw),
                torch.empty(B, dtype=torch.int32, **kw),
                torch.empty(B, K, dtype=torch.float32, **kw) if want_J_all else None)

    def plan_host(self, x0, U0, goal, method="adam", iters=20, lr=1e-2, b1=0.9, b2=0.999,
                  eps=1e-8, out=None):
        """Same as plan() with HOST (cpu, ideally pinned) tensors in and out; synchronous."""
        B, K = U0.shape[0], U0.shape[1]
        cpu = torch.device("cpu")
        if out is None:
            out = self.alloc_plan_outputs(B, K, True, device=cpu, pin=True)
        U_best, X_best, J_best, idx, J_all = out
        _check(self.lib.gmpc_plan_host(
            self._h, B, K, _ptr(x0, device=cpu, name="x0"), _ptr(U0, device=cpu, name="U0"),
            _ptr(goal, device=cpu, name="goal"), METHODS[method], iters, lr, b1, b2, eps,
            _ptr(U_best, device=cpu), _ptr(X_best, device=cpu), _ptr(J_best, device=cpu),
            _ptr(idx, dtype=torch.int32, device=cpu), _ptr(J_all, device=cpu),
            _stream(self.device)))
        return out

    def ilqr(self, x0, U0, goal, maxiter=100, grad_norm_threshold=1e-4, alpha_0=1.0,
             alpha_min=0.00005, want_lqr=False, gradient_lag=False, **unsupported):
        """trajax iLQR (the reference's ilqr_solve, policy/optimizers.py:10-21), batched:
        x0 [B,n], U0 [B,T,m], goal [B,T+1,n] ->
        (X, U, obj, gradient, adjoints, (A, B) or None, iteration).  Keyword names are those of
        TRAJAX_iLQR_KWARGS (policy/eval.py:10-20); the thresholds the reference leaves at 0.0 and
        make_psd=False must keep those values."""
        for k, v in unsupported.items():
            if k not in ("relative_grad_norm_threshold", "obj_step_threshold",
                         "inputs_step_threshold", "make_psd", "psd_delta"):
                raise TypeError(f"ilqr: unknown option {k!r}")
            if v not in (0.0, False):
                raise NotImplementedError(f"ilqr: {k}={v!r} (the reference uses 0.0 / False)")
        dev = self.device
        B = x0.shape[0]
        f = dict(device=dev, dtype=torch.float32)
        X = torch.empty(B, self.T + 1, self.n, **f)
        U = torch.empty(B, self.T, self.m, **f)
        obj = torch.empty(B, **f)
        grad = torch.empty(B, self.T, self.m, **f)
        lam = torch.empty(B, self.T + 1, self.n, **f)
        it = torch.empty(B, device=dev, dtype=torch.int32)
        A = torch.empty(B, self.T, self.n, self.n, **f) if want_lqr else None
        Bm = torch.empty(B, self.T, self.n, self.m, **f) if want_lqr else None
        opt = IlqrOptions(int(maxiter), float(grad_norm_threshold), float(alpha_0), float(alpha_min),
                          int(bool(gradient_lag)))
        _check(self.lib.gmpc_ilqr(
            self._h, B, _ptr(x0, device=dev, name="x0"), _ptr(U0, device=dev, name="U0"),
            _ptr(goal, device=dev, name="goal"), C.byref(opt), _ptr(X), _ptr(U), _ptr(obj),
            _ptr(grad), _ptr(lam), _ptr(it, dtype=torch.int32), _ptr(A), _ptr(Bm), _stream(dev)))
        return X, U, obj, grad, lam, ((A, Bm) if want_lqr else None), it

    def ilqr_host(self, x0, U0, goal, maxiter=100, grad_norm_threshold=1e-4, alpha_0=1.0,
                  alpha_min=0.00005, gradient_lag=False):
        """gmpc_ilqr_host: HOST (cpu) tensors in and out, synchronous: (X, U, obj, gradient, adjoints,
        None, iteration)."""
        cpu = torch.device("cpu")
        B = x0.shape[0]
        f = dict(dtype=torch.float32)
        X = torch.empty(B, self.T + 1, self.n, **f)
        U = torch.empty(B, self.T, self.m, **f)
        obj = torch.empty(B, **f)
        grad = torch.empty(B, self.T, self.m, **f)
        lam = torch.empty(B, self.T + 1, self.n, **f)
        it = torch.empty(B, dtype=torch.int32)
        opt = IlqrOptions(int(maxiter), float(grad_norm_threshold), float(alpha_0), float(alpha_min),
                          int(bool(gradient_lag)))
        _check(self.lib.gmpc_ilqr_host(
            self._h, B, _ptr(x0, device=cpu, name="x0"), _ptr(U0, device=cpu, name="U0"),
            _ptr(goal, device=cpu, name="goal"), C.byref(opt), _ptr(X), _ptr(U), _ptr(obj), _ptr(grad),
            _ptr(lam), _ptr(it, dtype=torch.int32), _stream(self.device)))
        return X, U, obj, grad, lam, None, it

    def bilevel_l2(self, x0, U0, goal, desired, maxiter=100, grad_norm_threshold=1e-4, alpha_0=1.0,
                   alpha_min=0.00005, want_hessian=False, V=None, gradient_lag=False, **unsupported):
        """bilevel_optimization (policy/optimizers.py:34-75) for the L2 loss, batched; see
        include/gmpc.h.  Returns a dict of X, U, obj, low_level_grad, iteration, loss, B, hessian,
        H, dxT, grad_mpc_weights."""
        for k, v in unsupported.items():
            if k not in ("relative_grad_norm_threshold", "obj_step_threshold",
                         "inputs_step_threshold", "make_psd", "psd_delta"):
                raise TypeError(f"bilevel_l2: unknown option {k!r}")
            if v not in (0.0, False):
                raise NotImplementedError(f"bilevel_l2: {k}={v!r} (the reference uses 0.0 / False)")
        dev = self.device
        B, T, n, m = x0.shape[0], self.T, self.n, self.m
        f = dict(device=dev, dtype=torch.float32)
        o = dict(X=torch.empty(B, T + 1, n, **f), U=torch.empty(B, T, m, **f), obj=torch.empty(B, **f),
                 low_level_grad=torch.empty(B, T, m, **f),
                 iteration=torch.empty(B, device=dev, dtype=torch.int32), loss=torch.empty(B, **f),
                 B=torch.empty(B, T, m, **f),
                 hessian=torch.empty(B, T * m, T * m, **f) if want_hessian else None,
                 H=torch.empty(B, T, m, **f), dxT=torch.empty(B, n, **f),
                 grad_mpc_weights=torch.empty(B, 3, **f))
        opt = IlqrOptions(int(maxiter), float(grad_norm_threshold), float(alpha_0), float(alpha_min),
                          int(bool(gradient_lag)))
        _check(self.lib.gmpc_bilevel_l2(
            self._h, B, _ptr(x0, device=dev, name="x0"), _ptr(U0, device=dev, name="U0"),
            _ptr(goal, device=dev, name="goal"), _ptr(desired, device=dev, name="desired"), C.byref(opt),
            _ptr(o["X"]), _ptr(o["U"]), _ptr(o["obj"]), _ptr(o["low_level_grad"]),
            _ptr(o["iteration"], dtype=torch.int32), _ptr(o["loss"]), _ptr(o["B"]), _ptr(o["hessian"]),
            _ptr(o["H"]), _ptr(o["dxT"]), _ptr(o["grad_mpc_weights"]),
            _ptr(V, device=dev, name="V"), _stream(dev)))
        return o

    def bilevel_tail(self, x0, U, goal, dLdX, want_hessian=False):
        """gmpc_bilevel_tail: the bilevel tail at U for a loss given by its state gradient dLdX."""
        dev = self.device
        B, T, n, m = x0.shape[0], self.T, self.n, self.m
        f = dict(device=dev, dtype=torch.float32)
        o = dict(B=torch.empty(B, T, m, **f),
                 hessian=torch.empty(B, T * m, T * m, **f) if want_hessian else None,
                 H=torch.empty(B, T, m, **f), dxT=torch.empty(B, n, **f),
                 grad_mpc_weights=torch.empty(B, 3, **f))
        _check(self.lib.gmpc_bilevel_tail(
            self._h, B, _ptr(x0, device=dev, name="x0"), _ptr(U, device=dev, name="U"),
            _ptr(goal, device=dev, name="goal"), _ptr(dLdX, device=dev, name="dLdX"), _ptr(o["B"]),
            _ptr(o["hessian"]), _ptr(o["H"]), _ptr(o["dxT"]), _ptr(o["grad_mpc_weights"]), _stream(dev)))
        return o

    def critic_input_grad(self, xseq, params_flat):
        """(logit [Bc], d logit / d xseq [Bc,T1,n])."""
        Bc, T1, n = xseq.shape
        logit = torch.empty(Bc, device=self.device, dtype=torch.float32)
        dx = torch.empty(Bc, T1, n, device=self.device, dtype=torch.float32)
        _check(self.lib.gmpc_critic_input_grad(
            self._h, Bc, T1, _ptr(xseq, device=self.device, name="xseq"),
            _ptr(params_flat, device=self.device, name="params_flat"), _ptr(logit), _ptr(dx),
            _stream(self.device)))
        return logit, dx

    def ilqr_stats(self):
        """(tile-level outer iterations, rollouts) of the ilqr calls since the last query."""
        a, b = C.c_int64(0), C.c_int64(0)
        _check(self.lib.gmpc_ilqr_stats(self._h, C.byref(a), C.byref(b), _stream(self.device)))
        return int(a.value), int(b.value)

    def dynamics_fit(self, xseq, useq, next_xseq, discount_factor, teacher_forcing, dims):
        """gmpc_dynamics_fit: (loss [B], act [L x [K_l,R]], cot [L x [N_l,R]]); dims = the dynamics MLP's
        layer widths [n+m, H, ..., n]."""
        dev = self.device
        B, S = xseq.shape[0], xseq.shape[1]
        R = int(self.lib.gmpc_dynamics_fit_columns(self._h, B, S))
        L = len(dims) - 1
        loss = torch.empty(B, device=dev, dtype=torch.float32)
        act = [torch.empty(dims[l], R, device=dev, dtype=torch.float32) for l in range(L)]
        cot = [torch.empty(dims[l + 1], R, device=dev, dtype=torch.float32) for l in range(L)]
        pa, pc = (_f * L)(), (_f * L)()
        for l in range(L):
            pa[l], pc[l] = act[l].data_ptr(), cot[l].data_ptr()
        _check(self.lib.gmpc_dynamics_fit(
            self._h, B, S, _ptr(xseq, device=dev, name="xseq"), _ptr(useq, device=dev, name="useq"),
            _ptr(next_xseq, device=dev, name="next_xseq"), float(discount_factor), int(bool(teacher_forcing)),
            _ptr(loss), pa, pc, _stream(dev)))
        return loss, act, cot

    def gemm_nt(self, A, B, alpha=1.0, want_rowsum=False):
        """gmpc_gemm_nt: A [M,R], B [N,R] -> (alpha A B^T [M,N], alpha rowsum(B) [N] or None)."""
        dev = self.device
        M, R = A.shape
        N = B.shape[0]
        Cm = torch.empty(M, N, device=dev, dtype=torch.float32)
        rs = torch.empty(N, device=dev, dtype=torch.float32) if want_rowsum else None
        _check(self.lib.gmpc_gemm_nt(self._h, M, N, R, _ptr(A, device=dev, name="A"), _ptr(B, device=dev, name="B"),
                                     float(alpha), 0, _ptr(Cm), _ptr(rs), _stream(dev)))
        return Cm, rs

    def cost_mixed_vjp(self, xT, dxT, scale, dims):
        """gmpc_cost_mixed_vjp: x_T [B,n], dx_T [B,n] -> (gW [L x [in_l,out_l]], gb [L x [out_l]]), batch-reduced and
        multiplied by `scale`; dims = the cost MLP's layer widths [n, H, ..., fout]."""
        dev = self.device
        B, L = xT.shape[0], len(dims) - 1
        gW = [torch.empty(dims[l], dims[l + 1], device=dev, dtype=torch.float32) for l in range(L)]
        gb = [torch.empty(dims[l + 1], device=dev, dtype=torch.float32) for l in range(L)]
        pW, pb = (_f * L)(), (_f * L)()
        for l in range(L):
            pW[l], pb[l] = gW[l].data_ptr(), gb[l].data_ptr()
        _check(self.lib.gmpc_cost_mixed_vjp(self._h, B, _ptr(xT, device=dev, name="xT"), _ptr(dxT, device=dev, name="dxT"),
                                            float(scale), pW, pb, _stream(dev)))
        return gW, gb

    def expert_propose(self, history_x, params_flat, lstm_features, num_layers, num_hidden_units):
        """gmpc_expert_propose: history_x [B,hist+1,n] -> (goal_xseq [B,T+1,n], init_useq [B,T,m])."""
        dev = self.device
        B, rows = history_x.shape[0], history_x.shape[1]
        goal = torch.empty(B, self.T + 1, self.n, device=dev, dtype=torch.float32)
        useq = torch.empty(B, self.T, self.m, device=dev, dtype=torch.float32)
        want = self.lib.gmpc_expert_param_count(self.n, self.m, lstm_features, num_layers, num_hidden_units)
        if params_flat.numel() != want:
            raise ValueError(f"expert_propose: params_flat has {params_flat.numel()} floats, the network needs {want}")
        _check(self.lib.gmpc_expert_propose(
            self._h, B, rows - 1, _ptr(history_x, device=dev, name="history_x"),
            _ptr(params_flat, device=dev, name="params_flat"), int(lstm_features), int(num_layers),
            int(num_hidden_units), _ptr(goal), _ptr(useq), _stream(dev)))
        return goal, useq

    # ---------------------------------------------------------------- critic / losses
    def critic_forward(self, xseq, params_flat):
        Bc, T1 = xseq.shape[0], xseq.shape[1]
        logit = torch.empty(Bc, device=self.device, dtype=torch.float32)
        _check(self.lib.gmpc_critic_forward(
            self._h, Bc, T1, _ptr(xseq, device=self.device, name="xseq"),
            _ptr(params_flat, device=self.device, name="params_flat"), _ptr(logit),
            _stream(self.device)))
        return logit

    def critic_loss_grad(self, xseq, label, params_flat, inv_count=None, want_grad=True,
                         perm=None, batch=None):
        """BCE loss (and flat gradient).  With `perm` (int32 [Bc]) xseq/label are the whole
        dataset and the minibatch is gathered inside the kernel."""
        dev = self.device
        T1 = xseq.shape[1]
        Bc = xseq.shape[0] if perm is None else perm.shape[0]
        if inv_count is None:
            inv_count = 1.0 / Bc
        loss = torch.empty(1, device=dev, dtype=torch.float32)
        grad = torch.empty(self.critic_param_count, device=dev, dtype=torch.float32) if want_grad else None
        if perm is None:
            _check(self.lib.gmpc_critic_loss_grad(
                self._h, Bc, T1, _ptr(xseq, device=dev, name="xseq"),
                _ptr(label, device=dev, name="label"),
                _ptr(params_flat, device=dev, name="params_flat"), inv_count, _ptr(loss),
                _ptr(grad), _stream(dev)))
        else:
            _check(self.lib.gmpc_critic_loss_grad_gather(
                self._h, Bc, T1, _ptr(xseq, device=dev, name="xseq"),
                _ptr(label, device=dev, name="label"),
                _ptr(perm, dtype=torch.int32, device=dev, name="perm"),
                _ptr(params_flat, device=dev, name="params_flat"), inv_count, _ptr(loss),
                _ptr(grad), _stream(dev)))
        return loss, grad

    def clip_adam_step(self, params_flat, grad_flat, mom, vel, step, lr, max_norm=100.0,
                       grad_scale=1.0, b1=0.9, b2=0.999, eps=1e-8):
        dev = self.device
        _check(self.lib.gmpc_clip_adam_step(
            self._h, params_flat.numel(), _ptr(params_flat, device=dev, name="params_flat"),
            _ptr(grad_flat, device=dev, name="grad_flat"), _ptr(mom, device=dev, name="mom"),
            _ptr(vel, device=dev, name="vel"), int(step), lr, max_norm, grad_scale, b1, b2, eps,
            _stream(dev)))

    def critic_train_scan(self, xseq, label, perm, params_flat, mom, vel, step0, lr, max_norm=100.0,
                          b1=0.9, b2=0.999, eps=1e-8):
        """The whole minibatch scan of one critic update in one call (single GPU): perm is int32
        [steps, Bc]; params/mom/vel are updated in place; returns the minibatch losses [steps]."""
        dev = self.device
        steps, Bc = perm.shape
        losses = torch.empty(steps, device=dev, dtype=torch.float32)
        scratch = torch.empty(self.critic_param_count, device=dev, dtype=torch.float32)
        _check(self.lib.gmpc_critic_train_scan(
            self._h, steps, Bc, xseq.shape[1], _ptr(xseq, device=dev, name="xseq"),
            _ptr(label, device=dev, name="label"), _ptr(perm, dtype=torch.int32, device=dev, name="perm"),
            _ptr(params_flat, device=dev, name="params_flat"), _ptr(mom, device=dev, name="mom"),
            _ptr(vel, device=dev, name="vel"), int(step0), lr, max_norm, b1, b2, eps, _ptr(losses),
            _ptr(scratch), _stream(dev)))
        return losses

    def l2_loss(self, X, desired):
        B = X.shape[0]
        loss = torch.empty(B, device=self.device, dtype=torch.float32)
        _check(self.lib.gmpc_l2_loss(self._h, B, _ptr(X, device=self.device, name="X"),
                                     _ptr(desired, device=self.device, name="desired"),
                                     _ptr(loss), _stream(self.device)))
        return loss


def measure_fp32_peak(device=0):
    """Measured FP32 FFMA TFLOP/s of `device` (microbenchmark inside libgmpc)."""
    out = C.c_float(0.0)
    _check(load().gmpc_measure_fp32_peak(int(device), C.byref(out)))
    return float(out.value)


def measure_f16_mma_peak(device=0):
    """Measured dense kind::f16 tcgen05 TFLOP/s of `device` (the pipe the planner kernels use)."""
    out = C.c_float(0.0)
    _check(load().gmpc_measure_f16_mma_peak(int(device), C.byref(out)))
    return float(out.value)
