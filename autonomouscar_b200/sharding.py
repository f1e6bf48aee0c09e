"""Host-side logic of the multi-GPU path (one process per GPU, torch.distributed).

The hot path shards naturally: poses of a batch and plan queries are independent
(SURVEY.md section 8(e)).  Every rank owns a contiguous block of the input, works on it
with NO data-path collective, and the results are gathered with one all-gather (NCCL over
NVLink on the GPU box; the same code runs on gloo/CPU tensors in the tests).
"""
import torch
import torch.distributed as dist


def shard_bounds(n, rank, world):
    """Contiguous block [lo, hi) of `n` items owned by `rank`; blocks differ by at most one
    item and concatenating them in rank order restores the input order."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def padded_block(n, world):
    """Items per rank for an equal-count all-gather (the tail of the last ranks is padding)."""
    return (n + world - 1) // world


def gather_blocks(local, n_total, group=None):
    """All-gather per-rank blocks of possibly unequal length along dim 0 and return the
    `n_total` rows in input order on every rank.  `local` is this rank's block, produced
    from shard_bounds()."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return local
    per = padded_block(n_total, world)
    pad_shape = (per,) + tuple(local.shape[1:])
    send = torch.zeros(pad_shape, dtype=local.dtype, device=local.device)
    send[: local.shape[0]] = local
    out = torch.empty((per * world,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out, send, group=group)
    pieces = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        pieces.append(out[r * per: r * per + (hi - lo)])
    return torch.cat(pieces, dim=0)


def max_over_ranks(value, device, group=None):
    """Timing rule of the bench contract: the slowest rank defines the step."""
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())


def plan_records(results, device):
    """Fixed-stride records of a batch of plan results, ready for gather_blocks():
    status, n_expansions, n_nodes, n_path, digest (low 63 bits)."""
    rows = [[r.status, r.n_expansions, r.n_nodes, r.n_path, int(r.tree_digest) & 0x7FFFFFFFFFFFFFFF]
            for r in results]
    return torch.tensor(rows, dtype=torch.int64, device=device).reshape(-1, 5)


def rollout_sharded(starts, plan_batch, ticks, device, route=None, masks_of=None, group=None):
    """Scenario sweep over the ranks (row N4 at N > 1): the cars are independent, so every rank drives its own
    contiguous block of them (rollout.RouteRollout around ITS planner) with no data-path collective, and the
    car states after every tick are gathered once at the end.  Returns the [ticks, n_cars, 4] trace in input
    order on every rank.  `plan_batch` as in rollout.RouteRollout.tick; `masks_of(rollout)` (optional) returns the
    region masks of the block's costmaps for the coming tick."""
    from . import rollout
    starts = [tuple(s) for s in starts]
    n = len(starts)
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_bounds(n, rank, world)
    ro = rollout.RouteRollout(starts[lo:hi], route=route)
    for _ in range(ticks):
        ro.tick(plan_batch, masks_of(ro) if masks_of is not None else None)
    local = torch.tensor(ro.trace, dtype=torch.float64, device=device).reshape(ticks, hi - lo, 4)
    # gather_blocks concatenates along dim 0: make the car index the leading dimension
    return gather_blocks(local.permute(1, 0, 2).contiguous(), n, group=group).permute(1, 0, 2).contiguous()
