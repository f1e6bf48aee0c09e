"""Host-side logic of the multi-GPU path (one process per GPU, torch.distributed).

The hot path shards naturally: poses of a batch and plan queries are independent
(SURVEY.md section 8(e)).  Every rank owns a contiguous block of the input, works on it
with NO data-path collective, and the results are gathered with one all-gather (NCCL over
NVLink on the GPU box; the same code runs on gloo/CPU tensors in the tests) or, for the packed
collision flags, pulled from the peers by the copy engines (ResultGather).
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


class ResultGather:
    """All-gather of equal-sized per-rank result blocks (the packed collision flags) that does not need an SM.

    transport "symm": the send buffers live in symmetric peer memory (torch.distributed._symmetric_memory: every
    rank's buffer mapped into every rank's address space).  A gather is a signal-pad barrier (all ranks' blocks are
    in place), one plain device-to-device copy per peer -- copy engines over NVLink, pulled -- and a second barrier
    (everyone has pulled: the send buffer may be rewritten), all on a side stream.  It therefore runs beside a
    persistent kernel that occupies every SM, where a collective's own kernel would have to wait for it.
    transport "nccl" (also the fallback when the platform refuses the peer mappings): one all_gather_into_tensor.

    `slots` send / receive buffer pairs allow the gather of one step to overlap the work of the next:
        g = ResultGather(n_words, torch.int32, device); kernel writes g.send[i]; ev = g.gather(i, after=event_of_kernel)
    `ev` completes when g.all[i] (shape [world, n]) holds every rank's block and g.send[i] may be rewritten."""

    def __init__(self, n, dtype, device, group=None, transport="symm", slots=2):
        self.group = group if group is not None else dist.group.WORLD
        self.world, self.rank = dist.get_world_size(self.group), dist.get_rank(self.group)
        self.n, self.dtype, self.device = int(n), dtype, device
        self.transport, self.handles, self.stream = ("nccl" if device.type == "cuda" else "collective"), None, None
        if transport == "symm" and device.type == "cuda":
            try:
                import torch.distributed._symmetric_memory as symm
                send = [symm.empty(self.n, dtype=dtype, device=device) for _ in range(slots)]
                self.handles = [symm.rendezvous(b, self.group) for b in send]
                self.send = send
                self.stream = torch.cuda.Stream(device=device, priority=-1)
                self.transport = "symm"
            except Exception as exc:   # noqa: BLE001 -- any failure of the optional transport means NCCL
                self.why_not_symm = f"{type(exc).__name__}: {exc}"
                self.handles = None
        if self.handles is None:
            self.send = [torch.empty(self.n, dtype=dtype, device=device) for _ in range(slots)]
        self.all = [torch.empty((self.world, self.n), dtype=dtype, device=device) for _ in range(slots)]

    def gather(self, slot, after=None):
        """Enqueue the gather of send[slot] into all[slot]; `after` = CUDA event the producer recorded.  Returns the
        event that marks its completion (None for CPU tensors: the collective has completed on return)."""
        if self.device.type != "cuda":
            dist.all_gather_into_tensor(self.all[slot].view(-1), self.send[slot], group=self.group)
            return None
        done = torch.cuda.Event()
        if self.handles is not None:
            if after is not None:
                self.stream.wait_event(after)
            else:
                self.stream.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(self.stream):
                hdl, rows = self.handles[slot], self.all[slot]
                hdl.barrier()
                for shift in range(self.world):
                    r = (self.rank - shift) % self.world
                    rows[r].copy_(hdl.get_buffer(r, (self.n,), self.dtype))
                hdl.barrier()
                done.record(self.stream)
        else:
            cur = torch.cuda.current_stream(self.device)
            if after is not None:
                cur.wait_event(after)
            dist.all_gather_into_tensor(self.all[slot].view(-1), self.send[slot], group=self.group)
            done.record(cur)
        return done


def plan_batch_gather(planner, queries, n_total, cap_path=128, group=None, buffers=None):
    """BASELINE configs[4]: this rank plans ITS contiguous block of the n_total queries (`queries` = that block,
    from shard_bounds) with pp_plan_batch_dev -- the fixed-stride records (summary + path + profile of every plan,
    local_planner.cpp:503-583) stay on the device -- and ONE all-gather moves them over NVLink, so that every rank
    holds all n_total records in input order as a uint8 tensor [n_total (padded to world * per), stride].  Nothing
    is copied to the host here; a rank that wants the results reads them once (results.PlanRecords).
    `buffers` (optional dict) keeps the send / receive tensors across calls."""
    import torch
    from .capi import pp_query
    from .results import record_stride
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    per = padded_block(n_total, world)
    stride = record_stride(cap_path)
    dev = planner.torch_device
    buffers = buffers if buffers is not None else {}
    key = (per, stride, world)
    if buffers.get("key") != key:
        buffers["key"] = key
        buffers["send"] = torch.zeros((per, stride), dtype=torch.uint8, device=dev)
        buffers["all"] = torch.empty((per * world, stride), dtype=torch.uint8, device=dev) if world > 1 else None
    send = buffers["send"]
    m = len(queries)
    assert m <= per
    q_arr = queries if not isinstance(queries, (list, tuple)) else (pp_query * m)(*queries)
    if m:
        planner.plan_batch_dev(q_arr, m, cap_path, out=send)
    if world == 1:
        return send
    if dev.type == "cuda":
        ev = torch.cuda.Event()
        ev.record(planner.stream)
        torch.cuda.current_stream(dev).wait_event(ev)      # the gather is ordered after the plan kernel, on the device
    dist.all_gather_into_tensor(buffers["all"], send, group=group)
    return buffers["all"]


def gathered_rows(n_total, world):
    """Row indices of the n_total real records inside a gathered [world * per, ...] block (the last ranks' blocks
    end in padding when n_total is not a multiple of world)."""
    import numpy as np
    per = padded_block(n_total, world)
    idx = []
    for r in range(world):
        lo, hi = shard_bounds(n_total, r, world)
        idx.append(np.arange(r * per, r * per + (hi - lo)))
    return np.concatenate(idx) if idx else np.zeros(0, dtype=np.int64)


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
