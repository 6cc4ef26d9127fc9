"""gan_mpc_b200 -- B200-native planning hot path of returaj/gan_mpc.

Host-side mirror of the reference's Python API for the planner path (policy/optimizers,
EvalMPC/BaseMPC/L2MPC/JS_MPC, critic_trainer) over libgmpc.so (include/gmpc.h).  Hand-written
CUDA for sm_100a; PyTorch is plumbing (device memory, streams, torch.distributed) only.
"""

__all__ = ["_lib", "synthetic"]
