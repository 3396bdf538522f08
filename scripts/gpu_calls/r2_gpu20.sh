#!/bin/bash
# Round-2 GPU call 20: compute-sanitizer (memcheck, racecheck, synccheck) on the queue kernel: lane-per-problem rounds, solo engine
# (sticky and one-iteration visits), plant noise and start RNG kernels; then smoke() and the fixed tests.
mkdir -p gpurun_out
O=gpurun_out
cat > /tmp/sani.py <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np, torch
import bench
from tfmpc_b200 import envs, ops
from tfmpc_b200.solvers.ilqr import iLQR
cfg = bench.workload_cfg("c3")
env = envs.make_env(cfg)
solver = iLQR(env, max_iterations=12)
for B, opts in ((700, {"queue_mode": 1}), (300, {"queue_mode": 2}), (96, {"queue_solo_max": 4, "queue_w_target": 32, "queue_w_solo": 1})):
    for k, v in opts.items():
        ops.set_option(k, v)
    x0, u0 = bench.make_inputs(cfg, B, 50, seed=3)
    out = solver.solve_device(x0, 50, u_init=u0)
    torch.cuda.synchronize()
    print("B", B, opts, "status", np.bincount(out["stats"][:, 3].cpu().numpy(), minlength=6).tolist())
nat = env.native()
x = torch.rand(1000, 2, device="cuda"); u = torch.rand(1000, 2, device="cuda")
ops.env_step_noisy(nat, x, u, 1, 2)
solver.initial_actions(100, 7, seed=1)
res = envs.make_env(bench.workload_cfg("c4"))
ops.env_step_noisy(res.native(), torch.rand(100, 20, device="cuda") * 50 + 20, torch.rand(100, 20, device="cuda"), 1, 2)
torch.cuda.synchronize()
print("done")
PY
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --print-limit 5 python /tmp/sani.py > $O/g20_sanitizer_$tool.log 2>&1
  echo "== $tool rc=$?" | tee -a $O/g20_summary.txt
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard|Invalid|done|status" $O/g20_sanitizer_$tool.log | head -12 | tee -a $O/g20_summary.txt
done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -n 2 | tee -a $O/g20_summary.txt
timeout 600 python -m pytest tests/test_gpu_parity.py -q -k "launchers or cli" 2>&1 | tail -n 3 | tee -a $O/g20_summary.txt
