"""GPU box helper: uncontended per-iteration cost of the solo engine against the lane-per-problem path -- a batch of B
lone problems (one warp each), phase times from the scheduling trace."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tfmpc_b200 import _native, envs, ops
from tfmpc_b200.solvers.ilqr import iLQR

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=148)
a = ap.parse_args()
T = 50
cfg = bench.workload_cfg("c3")
env = envs.make_env(cfg)
nat = env.native()
opts = iLQR(env)._opts()
x0, u0 = bench.make_inputs(cfg, a.batch, T, seed=1000)
x0, u0 = torch.from_numpy(x0).cuda(), torch.from_numpy(u0).cuda()
dev = torch.device("cuda", 0)
ops.set_option("queue_trace", 1)
ops.set_option("queue_w_target", 1 << 20)
for solo in (0, 1):
    ops.set_option("queue_solo_max", solo)
    out = ops.ilqr_solve(nat, x0, u0, opts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = ops.ilqr_solve(nat, x0, u0, opts, out); e1.record()
    torch.cuda.synchronize()
    t = ops.queue_trace(nat, a.batch, T, _native._WS_CACHE[(dev, torch.cuda.current_stream().cuda_stream)])
    its = float((out["stats"][:, 0] + 1).sum())
    if solo:
        iters = t[:, 7].sum()
        print(f"solo: batch {e0.elapsed_time(e1):.3f} ms, {len(t)} visits, {iters} iterations; per iteration: total {t[:, 2].sum() / iters / 1e3:.2f} us = "
              f"linearise {t[:, 4].sum() / iters / 1e3:.2f} + backward {t[:, 5].sum() / iters / 1e3:.2f} + search {t[:, 6].sum() / iters / 1e3:.2f} us (+ rest)")
    else:
        print(f"lane-per-problem: batch {e0.elapsed_time(e1):.3f} ms, {len(t)} warp iterations; per iteration: total {t[:, 2].mean() / 1e3:.2f} us = "
              f"set-up {t[:, 4].mean() / 1e3:.2f} + backward {t[:, 5].mean() / 1e3:.2f} + search {t[:, 6].mean() / 1e3:.2f} + store {t[:, 7].mean() / 1e3:.2f} us")
