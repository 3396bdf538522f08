"""Multi-GPU tests of the sharded solve over NCCL (one process per GPU, torchrun): need >= 2 GPUs, skipped otherwise.
`gpurun --gpus 2 -- python -m pytest tests/test_gpu_multi.py -m gpu` runs them; the CPU counterpart (gloo, world size 2) is
tests/test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

_RANK_SCRIPT = r"""
import os, sys
sys.path.insert(0, {root!r})
import numpy as np, torch, torch.distributed as dist
from tfmpc_b200 import envs, sharding
from tfmpc_b200.envs import synthetic
from tfmpc_b200.solvers.ilqr import iLQR
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
cfg = synthetic.navigation_config()
rng = np.random.RandomState(0)                       # the same global batch on every rank
B, T = 3001, 50                                      # ragged: 1501 + 1500
x0 = synthetic.sample_x0(cfg, B, rng).astype(np.float32)
u0 = synthetic.sample_u_init([-1, -1], [1, 1], B, T, rng).astype(np.float32)
ok = True
for dtype in (torch.float64, torch.float32):
    solver = iLQR(envs.make_env(cfg), dtype=dtype)
    dx0, du0 = torch.from_numpy(x0).cuda().to(dtype), torch.from_numpy(u0).cuda().to(dtype)
    local_out, full = sharding.solve_sharded(solver, dx0, T, du0, gather="full")
    _, summ = sharding.solve_sharded(solver, dx0, T, du0)
    whole = solver.solve_device(dx0, T, u_init=du0)   # the same batch on this GPU alone: problems are independent
    torch.cuda.synchronize()
    lo, hi = sharding.shard_range(B, rank, world)
    ok = ok and all(full[k].shape == whole[k].shape for k in whole)
    ok = ok and all(torch.equal(local_out[k], full[k][lo:hi]) for k in whole)          # the gather put every block where it belongs
    ok = ok and torch.equal(summ[1], full["stats"][:, 0]) and torch.allclose(summ[0], full["costs"].sum(1))
    _, fullp, perm = sharding.solve_sharded(solver, dx0, T, du0, gather="full", permute_seed=3)   # blocks cut from a seeded shuffle
    ok = ok and sorted(perm.tolist()) == list(range(B)) and all(torch.equal(fullp[k], full[k]) for k in whole)   # schedule-independent results
    # a problem's result does not depend on the schedule (which GPU, which warp, pop sizes, solo engine or not), in both builds:
    # the fp64 build is compiled without FMA contraction, the fp32 build spells its FMAs out where the code shapes differ
    ok = ok and all(torch.equal(full[k], whole[k]) for k in whole)
flag = torch.tensor([1 if ok else 0], device="cuda")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    print("SHARDED_OK" if int(flag) == 1 else "SHARDED_MISMATCH", flush=True)
dist.destroy_process_group()
"""


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_solve_sharded_under_nccl_two_gpus(tmp_path):
    """sharding.solve_sharded on 2 GPUs over NCCL: each rank solves its contiguous (ragged) block of a 3,001-problem C3-type batch on
    its own GPU; the full-result gather (one all_gather_into_tensor per buffer) and the summary gather must both reproduce the solve of
    the whole batch on one GPU bit for bit, in both builds; cutting the blocks from a seeded
    shuffle (permute_seed) returns bit-identical results in the original order."""
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT.format(root=ROOT))
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29517", str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-3000:]
    assert "SHARDED_OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
