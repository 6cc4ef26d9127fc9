"""One process per GPU over torch.distributed (NCCL on the box, gloo in CPU tests).

Planning is embarrassingly parallel over start states (the reference's batch axis is a plain
vmap, policy/base.py:122-125): contiguous blocks of states per rank, weights replicated, no
data-path collective.  The two exchanges the north star names are here: the gather of the best
plans and the sum all-reduce of the flat critic gradient."""

import torch
import torch.distributed as dist


def rank_world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_range(B, rank=None, world=None):
    """contiguous block [lo, hi) of rank `rank` (all K candidates of a state stay on one rank);
    block sizes differ by at most one."""
    if rank is None or world is None:
        rank, world = rank_world()
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_rows(local, B):
    """all-gather row blocks produced by shard_range back into the full [B, ...] tensor."""
    rank, world = rank_world()
    if world == 1:
        return local
    cap = (B + world - 1) // world
    pad = torch.zeros((cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = torch.empty((world * cap,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_range(B, r, world)
        parts.append(out[r * cap:r * cap + (hi - lo)])
    return torch.cat(parts, dim=0)


def gather_best_plans(U_best, J_best, idx_best, B):
    """north star: 'gather of the best plans' -- (U*[B,T,m], J*[B], idx[B]) on every rank."""
    return gather_rows(U_best, B), gather_rows(J_best, B), gather_rows(idx_best, B)


def allreduce_sum_(t):
    """in-place sum all-reduce (flat critic gradient / loss)."""
    _, world = rank_world()
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def plan_sharded(handle, x0, U0, goal, **plan_kwargs):
    """Plan this rank's block of the global batch (host/CPU or device tensors holding the FULL
    batch on every rank) and gather the best plans.  Returns full-batch (U*, J*, idx)."""
    B = x0.shape[0]
    lo, hi = shard_range(B)
    dev = handle.device
    f = lambda t: t[lo:hi].to(dev).contiguous()
    U, X, J, idx, _ = handle.plan(f(x0), f(U0), f(goal), **plan_kwargs)
    return gather_best_plans(U, J, idx, B)


def ilqr_sharded(handle, x0, U0, goal, **ilqr_kwargs):
    """trajax-iLQR planning (gmpc_ilqr) of this rank's block of the global batch, plans gathered on
    every rank: full-batch (U [B,T,m], obj [B], iteration [B])."""
    B = x0.shape[0]
    lo, hi = shard_range(B)
    dev = handle.device
    f = lambda t: t[lo:hi].to(dev).contiguous()
    _, U, obj, _, _, _, it = handle.ilqr(f(x0), f(U0), f(goal), **ilqr_kwargs)
    return gather_rows(U, B), gather_rows(obj, B), gather_rows(it, B)
