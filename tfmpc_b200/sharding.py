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


def balancing_permutation(batch, seed, device=None):
    """The same pseudo-random order of `batch` problems on every rank (seeded host generator).  Iteration counts differ by up
    to 10x between problems; when neighbouring problems are alike (a grid of initial states, sorted scenarios) contiguous
    blocks give the ranks unequal work, and cutting the blocks from a shuffled order evens it out."""
    perm = torch.randperm(int(batch), generator=torch.Generator().manual_seed(int(seed)))
    return perm if device is None else perm.to(device)


def unpermute(tensor, perm):
    """Inverse of `tensor[perm]` along the leading axis."""
    out = torch.empty_like(tensor)
    out[perm] = tensor
    return out


def solve_sharded(solver, x0, T, u_init, group=None, gather="summaries", permute_seed=None):
    """iLQR solve of the global batch (x0 [B,n], u_init [B,T,m], identical on every rank): each rank solves its
    block on its own GPU, then the summaries are gathered.  Returns (local result dict, (cost, iterations, status)); with
    gather="full" the second item is the dict of full-batch states / actions / costs / stats instead.

    permute_seed: cut the blocks from a seeded shuffle of the problems instead of their given order (load balance across
    ranks, see balancing_permutation); gathered results come back in the ORIGINAL order, the local dict holds this rank's
    problems `perm[lo:hi]` in shuffled order, and the permutation is appended to the return value."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = x0.shape[0]
    lo, hi = shard_range(B, rank, world)
    perm = None
    if permute_seed is not None:
        perm = balancing_permutation(B, permute_seed, x0.device)
        mine = perm[lo:hi]
        out = solver.solve_device(x0[mine].contiguous(), T, u_init=u_init[mine].contiguous())
    else:
        out = solver.solve_device(x0[lo:hi], T, u_init=u_init[lo:hi])
    if gather == "full":      # the gather SURVEY section 8(e) describes: costs, states, actions and iteration counts of every problem
        full = gather_results(out, B, group)
        if perm is None:
            return out, full
        return out, {k: unpermute(v, perm) for k, v in full.items()}, perm
    summary = gather_summaries(out["costs"].sum(1), out["stats"][:, 0].contiguous(), out["stats"][:, 3].contiguous(), B, group)
    if perm is None:
        return out, summary
    return out, tuple(unpermute(t, perm) for t in summary), perm
