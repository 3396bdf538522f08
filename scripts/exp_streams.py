import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from tfmpc_b200 import envs, ops
from tfmpc_b200.solvers.ilqr import iLQR
desc, B, T = bench.WORKLOADS["c3"]
B = int(os.environ.get("B", B))
cfg = bench.workload_cfg("c3"); env = envs.make_env(cfg); solver = iLQR(env)
x0, u0 = bench.make_inputs(cfg, B, T, 1000); x0, u0 = torch.from_numpy(x0).cuda(), torch.from_numpy(u0).cuda()
nat, opts = env.native(), solver._opts()
for S in (1, 2, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(S)]
    outs = [None] * S
    for i in range(S):
        with torch.cuda.stream(streams[i]):
            outs[i] = ops.ilqr_solve(nat, x0, u0, opts)
    torch.cuda.synchronize()
    K = 2 * S
    t0 = time.perf_counter()
    for k in range(K):
        with torch.cuda.stream(streams[k % S]):
            ops.ilqr_solve(nat, x0, u0, opts, outs[k % S])
    t_enq = time.perf_counter() - t0
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"S={S} K={K} enqueue={1e3*t_enq:.1f} ms total={1e3*dt:.1f} ms per-solve={1e3*dt/K:.2f} ms")
