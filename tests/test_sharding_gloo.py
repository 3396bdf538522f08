"""World-size-2 gloo tests (CPU) of the host-side multi-GPU logic: contiguous sharding and the single end-of-solve
all-gather.  The per-rank compute is stubbed by the CPU oracle here -- these tests cover
the plumbing, the GPU tests cover the kernels."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def test_shard_range_partitions_the_batch():
    from tfmpc_b200.sharding import shard_range
    for B in (0, 1, 7, 8, 65536, 65537):
        for world in (1, 2, 3, 8):
            blocks = [shard_range(B, r, world) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == B
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import oracle
        from tfmpc_b200.envs import synthetic
        from tfmpc_b200.sharding import gather_summaries, shard_range
        cfg = synthetic.navlqr_config([5.5, -9.0], 5.0, -1.0, 1.0)
        rng = np.random.RandomState(0)          # the same global batch on every rank
        T = 10
        x0 = synthetic.sample_x0(cfg, B, rng)
        u0 = synthetic.sample_u_init([-1, -1], [1, 1], B, T, rng)
        lo, hi = shard_range(B, rank, world)
        o = oracle.Oracle("f32")
        r = o.ilqr_solve(o.make_env(cfg), x0[lo:hi], u0[lo:hi], nthreads=1)     # stand-in for the per-GPU solve
        cost, its, st = gather_summaries(torch.from_numpy(r["costs"].sum(1)), torch.from_numpy(r["iterations"]),
                                         torch.from_numpy(r["status"]), B)
        full = o.ilqr_solve(o.make_env(cfg), x0, u0, nthreads=1)
        ok = (cost.shape[0] == B and np.allclose(cost.numpy(), full["costs"].sum(1)) and (its.numpy() == full["iterations"]).all()
              and (st.numpy() == full["status"]).all())
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 11])
def test_sharded_solve_and_gather_world2(B):
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, B, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


def _worker_full(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tfmpc_b200.sharding import gather_results, shard_range
        T, n, m = 5, 2, 2
        g = torch.Generator().manual_seed(0)        # the same global results on every rank; each rank contributes its block
        glob = {"states": torch.rand(B, T + 1, n, generator=g), "actions": torch.rand(B, T, m, generator=g),
                "costs": torch.rand(B, T + 1, generator=g), "stats": torch.randint(0, 100, (B, 4), generator=g, dtype=torch.int32)}
        lo, hi = shard_range(B, rank, world)
        full = gather_results({k: v[lo:hi].clone() for k, v in glob.items()}, B)
        ret[rank] = all(full[k].shape == glob[k].shape and torch.equal(full[k], glob[k]) for k in glob)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 11, 2])
def test_full_result_gather_world2(B):
    """sharding.gather_results: one all_gather_into_tensor per buffer (states, actions, costs, stats) of RAGGED shards (padded to
    the largest block) reassembles the full batch on every rank -- the gather SURVEY section 8(e) describes."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_full, args=(world, port, B, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}


class _StandInSolver:
    """solve_device of the right shapes whose outputs are a per-problem function of the inputs, so that a shuffled and
    re-assembled batch can be compared row by row with the unsharded call."""
    def solve_device(self, x0, T, u_init):
        b = x0.shape[0]
        states = x0[:, None, :] + torch.arange(T + 1, dtype=x0.dtype)[None, :, None]
        costs = (states ** 2).sum(-1) + torch.cat([u_init.abs().sum(-1), torch.zeros(b, 1, dtype=x0.dtype)], 1)
        stats = torch.stack([(x0.abs().sum(1) * 7).to(torch.int32) % 100, torch.zeros(b, dtype=torch.int32),
                             torch.zeros(b, dtype=torch.int32), (x0[:, 0] > 0).to(torch.int32)], 1)
        return {"states": states, "actions": u_init.clone(), "costs": costs, "stats": stats}


def _worker_permuted(rank, world, port, B, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tfmpc_b200.sharding import shard_range, solve_sharded
        T = 4
        g = torch.Generator().manual_seed(1)
        x0 = torch.randn(B, 2, generator=g)
        u0 = torch.randn(B, T, 2, generator=g)
        ref = _StandInSolver().solve_device(x0, T, u0)
        local, full, perm = solve_sharded(_StandInSolver(), x0, T, u0, gather="full", permute_seed=5)
        lo, hi = shard_range(B, rank, world)
        ok = sorted(perm.tolist()) == list(range(B)) and all(torch.equal(full[k], ref[k]) for k in ref)
        ok = ok and all(torch.equal(local[k], ref[k][perm[lo:hi]]) for k in ref)
        _, (cost, its, st), perm2 = solve_sharded(_StandInSolver(), x0, T, u0, permute_seed=5)
        ok = ok and torch.equal(perm, perm2) and torch.allclose(cost, ref["costs"].sum(1)) and torch.equal(its, ref["stats"][:, 0]) \
            and torch.equal(st, ref["stats"][:, 3])
        _, plain = solve_sharded(_StandInSolver(), x0, T, u0, gather="full")
        ok = ok and all(torch.equal(plain[k], ref[k]) for k in ref)
        ret[rank] = bool(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("B", [9, 16])
def test_solve_sharded_with_balancing_permutation_world2(B):
    """solve_sharded(permute_seed=...): every rank draws the same shuffle, solves its block of the shuffled order, and the
    gathered results (full or summaries) come back in the original problem order (a random permutation
    before sharding evens out the ranks)."""
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker_permuted, args=(world, port, B, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
