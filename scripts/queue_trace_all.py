"""GPU box helper: scheduling traces of EVERY solve of a pipelined run (S streams x R rounds, one workspace per solve), to
reconstruct how many warps of how many kernels are busy at each moment.  Writes gpurun_out/<tag>_traceall.npz."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tfmpc_b200 import envs, ops
from tfmpc_b200.solvers.ilqr import iLQR

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="qta")
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--rounds", type=int, default=4)
a = ap.parse_args()
desc, B, T = bench.WORKLOADS["c3"]
cfg = bench.workload_cfg("c3")
env = envs.make_env(cfg)
nat = env.native()
opts = iLQR(env)._opts()
x0, u0 = bench.make_inputs(cfg, B, T, seed=1000)
x0, u0 = torch.from_numpy(x0).cuda(), torch.from_numpy(u0).cuda()
ops.set_option("queue_trace", 1)
S, R = a.streams, a.rounds
streams = [torch.cuda.Stream() for _ in range(S)]
proto = ops.ilqr_solve(nat, x0, u0, opts)
torch.cuda.synchronize()
outs = [{k: torch.empty_like(v) for k, v in proto.items()} for _ in range(S)]
works = [[ops.ilqr_workspace(nat, B, T) for _ in range(R)] for _ in range(S)]
done = [[torch.cuda.Event() for _ in range(R)] for _ in range(S)]
for i, st in enumerate(streams):          # warm
    with torch.cuda.stream(st):
        ops.ilqr_solve_async(nat, x0, u0, outs[i], works[i][0], done[i][0], opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
main = torch.cuda.current_stream()
e0.record()
for st in streams:
    st.wait_event(e0)
for r in range(R):
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            ops.ilqr_solve_async(nat, x0, u0, outs[i], works[i][r], done[i][r], opts)
for st in streams:
    ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
e1.record()
torch.cuda.synchronize()
res = {"ms_per_batch": np.array(e0.elapsed_time(e1) / (S * R))}
for i in range(S):
    for r in range(R):
        t = ops.queue_trace(nat, B, T, works[i][r])
        res[f"s{i}_r{r}"] = np.stack([t[:, 0], t[:, 1], t[:, 2], t[:, 3] & 0xff, (t[:, 3] >> 8) & 0xff, t[:, 3] >> 16, t[:, 4], t[:, 5], t[:, 6], t[:, 7]], axis=1)
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/{a.tag}_traceall.npz", **res)
print(a.tag, "pipelined ms/batch", float(res["ms_per_batch"]))
