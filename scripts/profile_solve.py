"""GPU box helper for ncu: one warm solve, then ONE solve inside cudaProfilerStart/Stop.
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv \
        python scripts/profile_solve.py --workload c3
"""
import argparse, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from tfmpc_b200 import envs, ops
from tfmpc_b200.solvers.ilqr import iLQR

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="c3")
ap.add_argument("--batch", type=int, default=0)
ap.add_argument("--max-iterations", type=int, default=100)
a = ap.parse_args()
desc, B, T = bench.WORKLOADS[a.workload]
B = a.batch or B
cfg = bench.workload_cfg(a.workload)
env = envs.make_env(cfg)
solver = iLQR(env, max_iterations=a.max_iterations)
x0, u0 = bench.make_inputs(cfg, B, T, seed=1000)
x0, u0 = torch.from_numpy(x0).cuda(), torch.from_numpy(u0).cuda()
out = ops.ilqr_solve(env.native(), x0, u0, solver._opts())
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = ops.ilqr_solve(env.native(), x0, u0, solver._opts(), out)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
st = out["stats"].cpu().numpy()
print("problem-iterations", int((st[:, 0] + 1).sum()), "status", [int((st[:, 3] == i).sum()) for i in range(5)])
