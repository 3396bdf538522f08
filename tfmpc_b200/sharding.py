"""Sharding of a problem batch over the GPUs of one node.

Problems are independent, so the batch is cut into contiguous blocks, one per rank; nothing is exchanged while
solving, and the per-problem summaries (total cost, iteration count, status) are all-gathered once at the end
over whatever backend the process group uses (NCCL on GPUs, gloo in the CPU tests).
"""
import torch
import torch.distributed as dist


def shard_range(batch, rank, world):
    """Contiguous [lo, hi) block of `batch` problems for `rank` of `world`; sizes differ by at most one."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(int(batch), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard(tensor, rank, world):
    lo, hi = shard_range(tensor.shape[0], rank, world)
    return tensor[lo:hi]


def gather_summaries(total_cost, iterations, status, batch, group=None):
    """All-gather per-problem summaries of ragged shards into full-batch tensors (every rank gets them).

    total_cost [b_r] float, iterations [b_r] int32, status [b_r] int32 for this rank's shard of `batch`."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return total_cost, iterations, status
    rank = dist.get_rank(group)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    assert total_cost.shape[0] == sizes[rank]
    pad = max(sizes)

    def padded(t):
        out = torch.zeros(pad, dtype=t.dtype, device=t.device)
        out[: t.shape[0]] = t
        return out

    outs = []
    for t in (total_cost, iterations, status):
        bufs = [torch.empty(pad, dtype=t.dtype, device=t.device) for _ in range(world)]
        dist.all_gather(bufs, padded(t), group=group)
        outs.append(torch.cat([b[:s] for b, s in zip(bufs, sizes)]))
    return tuple(outs)


def gather_results(out, batch, group=None):
    """All-gather the FULL results of ragged shards -- states [b_r,T+1,n], actions [b_r,T,m], costs [b_r,T+1], stats [b_r,4]
    -- into full-batch tensors on every rank: one collective per buffer (~1 KB per C3 problem).
    Shards are padded to the largest block so that a single all_gather_into_tensor moves each buffer."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if world == 1:
        return dict(out)
    sizes = [shard_range(batch, r, world)[1] - shard_range(batch, r, world)[0] for r in range(world)]
    pad = max(sizes)
    full = {}
    for key in ("states", "actions", "costs", "stats"):
        t = out[key]
        if t.shape[0] != pad:
            tp = torch.zeros((pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
            tp[: t.shape[0]] = t
            t = tp
        buf = torch.empty((world * pad,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
        dist.all_gather_into_tensor(buf, t.contiguous(), group=group)
        if all(sz == pad for sz in sizes):
            full[key] = buf
        else:
            full[key] = torch.cat([buf[r * pad: r * pad + sizes[r]] for r in range(world)])
    return full


def solve_sharded(solver, x0, T, u_init, group=None, gather="summaries"):
    """iLQR solve of the global batch (x0 [B,n], u_init [B,T,m], identical on every rank): each rank solves its
    block on its own GPU, then the summaries are gathered.  Returns (local result dict, (cost, iterations, status)); with
    gather="full" the second item is the dict of full-batch states / actions / costs / stats instead."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x0.shape[0]
    lo, hi = shard_range(B, rank, world)
    out = solver.solve_device(x0[lo:hi], T, u_init=u_init[lo:hi])
    if gather == "full":      # the gather SURVEY section 8(e) describes: costs, states, actions and iteration counts of every problem
        return out, gather_results(out, B, group)
    summary = gather_summaries(out["costs"].sum(1), out["stats"][:, 0].contiguous(), out["stats"][:, 3].contiguous(), B, group)
    return out, summary
