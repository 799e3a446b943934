"""Batch sharding of a solve over the GPUs of one box (one process per GPU, torch.distributed).

Trajectories are independent along the batch axis (every op of the path is row-wise: reference
base_sde.py:98-102, all tableaus, brownian_interval.py:247-248), and the Brownian rows are Philox
streams keyed by the GLOBAL row index, so a sharded solve reproduces the unsharded one bit for bit
(tests/test_gpu_solver.py::test_batch_sharding_is_invisible).  There is no collective on the data
path.  The optional collectives are: one all-gather of the output series (or of the terminal
states), and — for training — one all-reduce of the adjoint's parameter gradients.
"""
import torch
import torch.distributed as dist


def row_range(n_rows, rank=None, world=None):
    """Contiguous rows [lo, hi) owned by `rank`; sizes differ by at most one."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world is None:
        world = dist.get_world_size() if dist.is_initialized() else 1
    base, rem = divmod(n_rows, world)
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def shard_rows(y0, bm_factory, rank=None, world=None):
    """Return (local y0, local Brownian motion).  `bm_factory(n_local_rows)` must build the
    BrownianInterval for the local rows with the SAME entropy on every rank."""
    lo, hi = row_range(y0.shape[0], rank, world)
    bm = bm_factory(hi - lo)
    bm.shard_rows(lo)
    return y0[lo:hi].contiguous(), bm


def all_gather_rows(x_local, n_rows, dim=0):
    """Gather row shards (possibly of unequal size) back into the full tensor on every rank."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x_local
    world = dist.get_world_size()
    sizes = [row_range(n_rows, r, world) for r in range(world)]
    max_rows = max(hi - lo for lo, hi in sizes)
    x = x_local.movedim(dim, 0).contiguous()
    pad = torch.zeros((max_rows, *x.shape[1:]), dtype=x.dtype, device=x.device)
    pad[:x.shape[0]] = x
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    full = torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)
    return full.movedim(0, dim)


def all_reduce_grads(params):
    """Sum parameter gradients over the batch shards (needed for sdeint_adjoint with a sharded batch)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    for p in params:
        if p.grad is not None:
            dist.all_reduce(p.grad, op=dist.ReduceOp.SUM)
