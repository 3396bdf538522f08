"""GPU box helper: scheduling traces of the queue solver (option queue_trace): one lone C3 batch, then S streams x R batches
in flight.  Writes gpurun_out/<tag>_trace.npz with, per solve, records [acquire start ns (32-bit), wait ns, work ns, lanes,
rounds, warp slot, set-up ns, backward ns, search ns, store-pass ns]."""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from tfmpc_b200 import _native, envs, ops
from tfmpc_b200.solvers.ilqr import iLQR

ap = argparse.ArgumentParser()
ap.add_argument("--tag", default="qt")
ap.add_argument("--streams", type=int, default=8)
ap.add_argument("--rounds", type=int, default=3)
a = ap.parse_args()
desc, B, T = bench.WORKLOADS["c3"]
cfg = bench.workload_cfg("c3")
env = envs.make_env(cfg)
nat = env.native()
opts = iLQR(env)._opts()
x0, u0 = bench.make_inputs(cfg, B, T, seed=1000)
x0, u0 = torch.from_numpy(x0).cuda(), torch.from_numpy(u0).cuda()
dev = torch.device("cuda", 0)
ops.set_option("queue_trace", 1)


def unpack(t):
    return np.stack([t[:, 0], t[:, 1], t[:, 2], t[:, 3] & 0xff, (t[:, 3] >> 8) & 0xff, t[:, 3] >> 16, t[:, 4], t[:, 5], t[:, 6], t[:, 7]], axis=1)


out = ops.ilqr_solve(nat, x0, u0, opts)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = ops.ilqr_solve(nat, x0, u0, opts, out); e1.record()
torch.cuda.synchronize()
main = torch.cuda.current_stream()
lone = unpack(ops.queue_trace(nat, B, T, _native._WS_CACHE[(dev, main.cuda_stream)]))
res = {"lone": lone, "lone_ms": np.array(e0.elapsed_time(e1))}
streams = [torch.cuda.Stream() for _ in range(a.streams)]
outs = [{k: torch.empty_like(v) for k, v in out.items()} for _ in streams]
for i, st in enumerate(streams):          # warm the per-stream workspaces
    with torch.cuda.stream(st):
        ops.ilqr_solve(nat, x0, u0, opts, outs[i])
torch.cuda.synchronize()
e0.record()
for st in streams:
    st.wait_event(e0)
for r in range(a.rounds):
    for i, st in enumerate(streams):
        with torch.cuda.stream(st):
            ops.ilqr_solve(nat, x0, u0, opts, outs[i])
for st in streams:
    ev = torch.cuda.Event(); ev.record(st); main.wait_event(ev)
e1.record()
torch.cuda.synchronize()
res["pipe_ms_per_batch"] = np.array(e0.elapsed_time(e1) / (a.rounds * a.streams))
for i, st in enumerate(streams):
    res[f"pipe{i}"] = unpack(ops.queue_trace(nat, B, T, _native._WS_CACHE[(dev, st.cuda_stream)]))
os.makedirs("gpurun_out", exist_ok=True)
np.savez_compressed(f"gpurun_out/{a.tag}_trace.npz", **res)
print(a.tag, "lone ms", float(res["lone_ms"]), "pipelined ms/batch", float(res["pipe_ms_per_batch"]), "records", len(lone))
