"""Diagnostic: host time to enqueue back-to-back tfmpc_ilqr_solve_async calls, before/after torch.distributed (NCCL) is set up."""
import os, sys, time, threading
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from tfmpc_b200 import envs, ops
from tfmpc_b200.solvers.ilqr import iLQR

rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
cfg = bench.workload_cfg("c3")
env = envs.make_env(cfg)
nat, opts = env.native(), iLQR(env)._opts()
B, T, S, K = 65536, 50, 8, 16
x0, u0 = (torch.from_numpy(a).to(dev) for a in bench.make_inputs(cfg, B, T, seed=1000 + rank))
outs = [ops.ilqr_solve(nat, x0, u0, opts) for _ in range(S)]
works = [ops.ilqr_workspace(nat, B, T, dev) for _ in range(S)]
done = [torch.cuda.Event() for _ in range(S)]
main = torch.cuda.current_stream()


def run(tag):
    torch.cuda.synchronize()
    per = []
    t0 = time.perf_counter()
    for k in range(K):
        s = k % S
        if k >= S:
            main.wait_event(done[s])
        a = time.perf_counter()
        ops.ilqr_solve_async(nat, x0, u0, outs[s], works[s], done[s], opts)
        per.append(1e3 * (time.perf_counter() - a))
    host = 1e3 * (time.perf_counter() - t0)
    torch.cuda.synchronize()
    tot = 1e3 * (time.perf_counter() - t0)
    print(f"[rank {rank}] {tag}: host enqueue {host / K:.2f} ms/step (per call min {min(per):.2f} max {max(per):.2f}), total {tot / K:.2f} ms/step, "
          f"threads {threading.active_count()}, affinity {len(os.sched_getaffinity(0))}, OMP {os.environ.get('OMP_NUM_THREADS')}", flush=True)


run("warm"); run("before dist")
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
    run("after init_process_group")
    dist.barrier(); torch.cuda.synchronize()
    run("after barrier")
    t = torch.zeros(1 << 20, device=dev)
    dist.all_gather([torch.empty_like(t) for _ in range(world)], t); torch.cuda.synchronize()
    run("after all_gather")
    dist.destroy_process_group()
    run("after destroy")
