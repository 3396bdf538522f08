#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched iLQR hot path (BASELINE.json: "batched iLQR
problem-iterations/sec at 1/2/4/8 B200 vs TF CPU ref").

    python bench.py --gpus 1 --steps 5 --warmup 3                      # this repo's CUDA path
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
           bench.py --gpus N --steps K --warmup W                      # N GPUs, one rank per GPU
    python bench.py --impl reference --steps 2 --warmup 1              # the CPU arm (oracle port, all host cores)

A "step" is one pass of the hot path over one batch: iLQR.solve of B independent problems to convergence
(reference tfmpc/solvers/ilqr.py:214-283 per problem).  The metric counts problem-iterations = passes of
the reference's outer loop ilqr.py:227-277 (linearise + >= 1 backward pass + line search), summed over
problems.  Workload (default): BASELINE config C3 -- nonlinear 2-D navigation with two deceleration zones,
H = 50, B = 65,536 problems PER GPU (weak scaling: the batch is sharded, no data-path collective; with
N > 1 ranks the per-problem costs and iteration counts resident at the end are all-gathered over NCCL once, inside
the timed region).
Prints ONE JSON line (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (description, batch per GPU, horizon)
    "c2": ("C2 navlin linear-navigation LQR, beta=5.0, H=10, random (x0, goal) pairs", 65536, 10),
    "c3": ("C3 iLQR nonlinear navigation (nav.config.json, 2 zones, actions in [-1,1]) H=50", 65536, 50),
    "c4": ("C4 iLQR reservoir control, 20 reservoirs, H=40", 16384, 40),
    "c5s": ("C5 (single solve) iLQR HVAC 32 rooms 4x8 grid, H=48", 16384, 48),
    "c5": ("C5 iLQR HVAC 32 rooms 4x8 grid, receding-horizon MPC loop: 48 plant steps, horizon 48-t", 16384, 48),
}


def workload_cfg(name):
    from tfmpc_b200.envs import synthetic
    return {"c3": synthetic.navigation_config, "c4": lambda: synthetic.reservoir_config(20),
            "c5s": lambda: synthetic.hvac_grid_config(4, 8), "c5": lambda: synthetic.hvac_grid_config(4, 8)}[name]()


def make_inputs(cfg, B, T, seed):
    from tfmpc_b200.envs import synthetic
    rng = np.random.RandomState(seed)
    x0 = synthetic.sample_x0(cfg, B, rng).astype(np.float32)
    if cfg["cls_name"] == "Navigation":
        lo, hi = np.ravel(cfg["config"]["low"]), np.ravel(cfg["config"]["high"])
    else:
        lo, hi = np.zeros(x0.shape[1]), np.ones(x0.shape[1])
    u0 = synthetic.sample_u_init(lo, hi, B, T, rng).astype(np.float32)
    return np.ascontiguousarray(x0), np.ascontiguousarray(u0)


# algorithmic work per problem-iteration, SURVEY.md section 8(d) / DESIGN.md "roofline":
#   FLOPs = H * (lin_env * iterations + bwd * backward passes + fwd * reference-semantics rollouts)
#   bytes = 4 * [(H(n+m)+n) read nominal + ((H+1)n + Hm + H+1) write candidate]  per problem-iteration (fused)
def algorithmic_work(name, T, stats):
    its = stats[:, 0].astype(np.float64) + 1.0
    nb, nf = stats[:, 1].astype(np.float64), stats[:, 2].astype(np.float64)
    if name == "c3":
        n = m = 2
        lin, bwd, fwd = 80.0, 329.0 + 100.0, 2 * m * n + 8 * m + n + 1 + 32.0
    elif name == "c4":
        n = m = 20
        lin, bwd, fwd = 12.0 * n, 2.0 * n * n + 2 * n * m + 5 * m, 8 * m + n + 1 + 2.0 * n * n + 21 * n
    else:
        n = m = 32
        lin, bwd, fwd = 6.0 * n, 2.0 * n * n + 2 * n * m + 5 * m, 8 * m + n + 1 + 3.0 * n * n + 25 * n
    flops = T * (lin * its + bwd * nb + fwd * nf)
    byts = 4.0 * ((T * (n + m) + n) + ((T + 1) * n + T * m + T + 1)) * its
    return float(flops.sum()), float(byts.sum())


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return None
        time.sleep(0.15)
        self.proc.terminate()
        rows = [r for ts, r in self.rows if t0 - 0.05 <= ts <= t1 + 0.15] or [r for _, r in self.rows]
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            f = [x.strip() for x in r.split(",")]
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except (ValueError, IndexError):
                continue
            for nme, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        if not sm:
            return None
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            return json.load(fh), "measured (MEASURED_PEAKS.json)"
    return {"hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def cpu_baseline(name, T, sample, threads=0, repeats=1):
    """The CPU arm: the oracle port (C, OpenMP over problems) on the host cores, on a bounded sample of the
    same workload.  Returns (problem-iterations/s, cores, problem-iterations, seconds)."""
    from oracle import oracle
    o = oracle.Oracle("f32")
    cfg = workload_cfg(name)
    x0, u0 = make_inputs(cfg, sample, T, seed=12345)
    env = o.make_env(cfg)
    cores = max(o.max_threads(), len(os.sched_getaffinity(0))) if threads <= 0 else threads   # every host core the process may use
    best = None
    for _ in range(repeats):
        t0 = time.perf_counter()
        r = o.ilqr_solve(env, x0, u0, nthreads=cores)
        dt = time.perf_counter() - t0
        pi = float((r["iterations"] + 1).sum())
        if best is None or dt < best[1]:
            best = (pi, dt)
    return best[0] / best[1], cores, best[0], best[1]


def reference_under_shim(name, T, problems=3):
    """Extra data point for the CPU arm: the UNMODIFIED reference sources (baseline/_ref, installed with pip --no-deps)
    executed under the torch-backed TensorFlow API shim (oracle/tf_shim), one process, a handful of problems.  This is the
    reference's own Python control flow (py_function callback per timestep and all) with torch CPU kernels in place of
    TensorFlow's; None when baseline/_ref is absent."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref, "tfmpc")):
        return None
    import tempfile
    cfg = workload_cfg(name)
    x0, u0 = make_inputs(cfg, problems, T, seed=12345)
    with tempfile.TemporaryDirectory() as tmp:
        json.dump(cfg, open(os.path.join(tmp, "env.json"), "w"))
        np.savez(os.path.join(tmp, "in.npz"), x0=x0, u0=u0)
        env = dict(os.environ, PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "tf_shim"), ref]), OMP_NUM_THREADS="1")
        try:
            r = subprocess.run([sys.executable, os.path.join(ROOT, "oracle", "run_reference_shim.py"), os.path.join(tmp, "env.json"),
                                os.path.join(tmp, "in.npz")], env=env, capture_output=True, text=True, timeout=600)
            d = json.loads(r.stdout.strip().splitlines()[-1])
        except Exception as exc:  # noqa: BLE001
            return {"error": str(exc)[:200]}
    return {"value_per_core": d["problem_iterations"] / d["seconds"], "unit": "problem-iterations/s", "cores": 1,
            "kind": "reference sources under a torch-backed TensorFlow API shim", "sample": f"{d['problems']} problems, {d['seconds']:.1f} s",
            "iterations": d["iterations"]}


def c2_inputs(B, seed):
    rng = np.random.RandomState(seed)
    return rng.uniform(-10, 10, size=(B, 2)).astype(np.float32), rng.normal(size=(B, 2)).astype(np.float32)


def c2_cpu(B, T, threads=0):
    """CPU arm of C2: the oracle's LQR (C, OpenMP over problems).  Returns (problems/s, cores, seconds)."""
    from oracle import oracle
    o = oracle.Oracle("f32")
    goal, x0 = c2_inputs(B, 12345)
    F = np.concatenate([np.eye(2), np.eye(2)], axis=1)
    c = np.concatenate([-2 * goal, np.zeros_like(goal)], axis=1)
    cores = max(o.max_threads(), len(os.sched_getaffinity(0))) if threads <= 0 else threads   # every host core the process may use
    t0 = time.perf_counter()
    o.lqr_solve(F, np.zeros(2), np.diag([2.0, 2.0, 10.0, 10.0]), c, x0, T, nthreads=cores)
    dt = time.perf_counter() - t0
    return B / dt, cores, dt


def run_c2(args):
    """BASELINE config C2: one LQR solve (backward Riccati sweep + forward rollout, policy and value function returned, as
    reference LQR.backward + LQR.forward do) per problem; a step = one batch of B problems.  The kernel is HBM-bound:
    24 bytes in, 760 bytes out per problem."""
    import torch
    import torch.distributed as dist

    from tfmpc_b200 import _native, envs, ops

    rank, local, world = (int(os.environ.get(k, d)) for k, d in (("RANK", "0"), ("LOCAL_RANK", "0"), ("WORLD_SIZE", "1")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc, B, T = WORKLOADS["c2"]
    B = args.batch or B
    goal_h, x0_h = c2_inputs(B, 1000 + rank)
    solver = envs.make_lqr_linear_navigation(goal_h, 5.0)
    x0_pin = torch.from_numpy(x0_h).pin_memory()
    x0 = x0_pin.to(dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    Fd, fd, Cd, cd = solver._device_params()
    out = ops.lqr_solve(Fd, fd, Cd, cd, x0, T)

    def step():
        return ops.lqr_solve(Fd, fd, Cd, cd, x0, T, out=out)   # trajectory + policy + value function, results reused in place

    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)
    for _ in range(max(3, args.warmup)):
        out = step()
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    t_wall0 = time.time()
    launches0 = _native.kernel_launch_count("f32")
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for a, b in ev:
        flush.zero_()
        a.record()
        out = step()
        b.record()
    barrier()
    launches = _native.kernel_launch_count("f32") - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    t = torch.tensor([sum(step_ms)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms = float(t[0])
    value = B * world * args.steps / (total_ms * 1e-3)
    # end to end: host buffers in, host buffers out (trajectory only, what LQR.solve returns)
    F, f, Cm, c = (t.cpu() for t in (Fd, fd, Cd, cd))
    c_pin = c.pin_memory() if c.dim() == 2 else c
    ho = {k: v.pin_memory() for k, v in ops.lqr_solve_host(F, f, Cm, c_pin, x0_pin, T).items()}
    ops.lqr_solve_host(F, f, Cm, c_pin, x0_pin, T, out=ho)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        ops.lqr_solve_host(F, f, Cm, c_pin, x0_pin, T, out=ho)
    barrier()
    te = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    clocks = sampler.stop(t_wall0, time.time()) if sampler else None
    if rank == 0:
        peaks, peak_src = measured_peaks()
        byts = B * (24 + 4 * ((T + 1) * 2 + T * 2 + (T + 1) + 1 + T * (4 + 2 + 4 + 2 + 1)))
        ach = byts / (total_ms / args.steps * 1e-3) / 1e9
        traffic, traffic_src = None, None
        tpath = os.path.join(ROOT, "profiles", "r01_traffic_c2.json")
        if os.path.exists(tpath):      # DRAM bytes per problem measured once with ncu (provenance inside the file)
            tr = json.load(open(tpath))
            traffic, traffic_src = tr["dram_bytes_per_problem"] * B, tr["source"]
        line = {"metric": "batched LQR problems/sec", "value": value, "unit": "problems/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
                "data": "synthetic",
                "config": {"workload": desc, "batch_per_gpu": B, "global_batch": B * world, "horizon": T, "state_dim": 2, "action_dim": 2,
                           "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                           "l2": "256 MB buffer written between timed iterations (untimed L2 flush)",
                           "status_nonzero": int((out["status"] != 0).sum())},
                "step_ms": step_ms,
                "roofline": {"bound": "hbm", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach / peaks["hbm_gbs"],
                             "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                             "kernel": "k_lqr_small_staged<2,2> (thread per problem, backward + forward fused, TMA bulk stores)",
                             "algorithmic_bytes_per_problem": byts // B},
                "e2e": {"value": B * world * args.steps / float(te[0]), "unit": "problems/s", "h2d_bytes_per_step": int(x0_pin.numel() * 4 + c.numel() * 4),
                        "d2h_bytes_per_step": int(sum(v.numel() * 4 for v in ho.values())), "api": "tfmpc_lqr_solve_host (pinned host buffers in and out, synchronous), one call per step"},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = c2_cpu(args.cpu_sample or B, T)
            line["cpu_baseline"] = {"value": v, "unit": "problems/s", "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_sample or B} problems in {dt:.2f} s, C/OpenMP restatement of reference LQR (oracle/)"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    desc, _, T = WORKLOADS[args.workload]
    if args.workload == "c2":
        B = args.cpu_sample or args.batch or WORKLOADS["c2"][1]
        for _ in range(args.warmup):
            c2_cpu(B, T)
        secs = 0.0
        for _ in range(args.steps):
            v, cores, dt = c2_cpu(B, T)
            secs += dt
        value = B * args.steps / secs
        print(json.dumps({"impl": "reference", "metric": "batched LQR problems/sec", "value": value, "unit": "problems/s", "n_gpus": args.gpus,
                          "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps, "higher_is_better": True,
                          "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                          "config": {"workload": desc, "batch_per_step": B, "horizon": T},
                          "cpu_baseline": {"value": value, "unit": "problems/s", "cores": cores, "kind": "port",
                                           "sample": f"{B} problems per step, {args.steps} steps"},
                          "e2e": {"value": value, "unit": "problems/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}), flush=True)
        return
    sample = args.cpu_sample or {"c3": 65536, "c4": 512, "c5s": 128, "c5": 128}[args.workload]
    for _ in range(args.warmup):
        cpu_baseline(args.workload, T, max(64, sample // 8))
    vals, secs, pis = [], 0.0, 0.0
    for _ in range(args.steps):
        v, cores, pi, dt = cpu_baseline(args.workload, T, sample)
        vals.append(v); secs += dt; pis += pi
    value = pis / secs
    line = {"impl": "reference", "metric": "batched iLQR problem-iterations/sec", "value": value, "unit": "problem-iterations/s",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "batch_per_step": sample, "horizon": T,
                       "note": "CPU arm = this repo's C/OpenMP restatement of the reference algorithm (oracle/); the TensorFlow "
                               "reference itself is not installable offline"},
            "cpu_baseline": {"value": value, "unit": "problem-iterations/s", "cores": cores, "kind": "port",
                             "sample": f"{sample} problems of the workload per step, {args.steps} steps"},
            "e2e": {"value": value, "unit": "problem-iterations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    if args.workload == "c3":
        line["reference_under_shim"] = reference_under_shim(args.workload, T)
    print(json.dumps(line), flush=True)


def run_ours(args):
    import torch
    import torch.distributed as dist

    from tfmpc_b200 import _native, envs, ops
    from tfmpc_b200.solvers.ilqr import iLQR

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    desc, B, T = WORKLOADS[args.workload]
    if args.batch:
        B = args.batch
    cfg = workload_cfg(args.workload)
    env = envs.make_env(cfg)
    solver = iLQR(env)
    n, m = env.state_size, env.action_size
    x0_h, u0_h = make_inputs(cfg, B, T, seed=1000 + rank)
    x0_pin, u0_pin = torch.from_numpy(x0_h).pin_memory(), torch.from_numpy(u0_h).pin_memory()
    x0, u0 = x0_pin.to(dev), u0_pin.to(dev)
    S = max(1, min(args.streams, args.steps))
    nat, opts = env.native(), solver._opts()

    def new_out(host=False):
        kw = {} if host else {"device": dev}
        o = {"states": torch.empty(B, T + 1, n, **kw), "actions": torch.empty(B, T, m, **kw), "costs": torch.empty(B, T + 1, **kw),
             "stats": torch.empty(B, 4, dtype=torch.int32, **kw)}
        return {k: v.pin_memory() for k, v in o.items()} if host else o

    outs = [new_out() for _ in range(S)]
    flush = torch.empty(256 * 1024 * 1024 // 4, device=dev)   # 256 MB > 126 MB L2
    streams = [torch.cuda.Stream(device=dev) for _ in range(S)]
    main = torch.cuda.current_stream()
    use_async = args.pipeline == "async" and args.workload in ("c3",) and S > 1
    if use_async:
        works = [ops.ilqr_workspace(nat, B, T, dev) for _ in range(S)]
        done = [torch.cuda.Event() for _ in range(S)]

    # Sharded solve: no inter-GPU traffic while solving, and nothing but the solves is enqueued while they run (a gather --
    # or any other device operation -- per step on the compute streams was measured first: -25 % at 2 GPUs).  After the
    # last batch the per-problem total costs and iteration counts of the S batches resident in the output ring are
    # reduced to [S, B] and ONE all-gather per buffer, inside the timed region, brings them to every rank.
    totals_all = torch.empty(S, B, device=dev) if world > 1 else None
    iters_all = torch.empty(S, B, dtype=torch.int32, device=dev) if world > 1 else None
    gathered = ([torch.empty_like(totals_all) for _ in range(world)], [torch.empty_like(iters_all) for _ in range(world)]) if world > 1 else None

    extra = os.environ.get("TFMPC_BENCH_EXTRA", "")   # diagnostics: what an op between two solves of a stream costs
    xbuf = torch.empty(B, device=dev)

    def step(slot, k=None):   # k: index of the timed step (unused by the solve itself)
        ops.ilqr_solve(nat, x0, u0, opts, outs[slot])
        if extra == "ops":
            torch.sum(outs[slot]["costs"], dim=1, out=xbuf)
        elif extra == "event":
            torch.cuda.Event().record()
        elif extra == "sleep":
            time.sleep(1e-4)
        elif extra == "memset":
            xbuf.zero_()
        elif extra == "own":
            ops.env_final_cost(nat, x0[:1024])
        elif extra == "memcpy":
            xbuf[:1024].copy_(xbuf[1024:2048])

    def final_gather():
        if world > 1:
            dist.all_gather(gathered[0], totals_all)
            dist.all_gather(gathered[1], iters_all)

    if args.workload == "c5":
        # BASELINE config 5: the shrinking-horizon MPC loop of reference agents/mpc.py:10-15 + runners/__init__.py:14-43 for B
        # plants at once: at plant step t re-solve over the remaining horizon T - t from fresh initial actions, apply the
        # first action to the (deterministic) plant.  One bench "step" = the whole 48-step loop; stats are summed on the device.
        mpc_stats = [torch.zeros(B, 4, dtype=torch.int32, device=dev) for _ in range(S)]
        u_inits = [torch.from_numpy(make_inputs(cfg, B, T - t, seed=2000 + rank + t)[1]).to(dev) for t in range(T)]

        def step(slot):  # noqa: F811
            state = x0
            mpc_stats[slot].zero_()
            for t in range(T):
                o = ops.ilqr_solve(nat, state, u_inits[t], opts)
                mpc_stats[slot][:, :3] += o["stats"][:, :3] + torch.tensor([1, 0, 0], dtype=torch.int32, device=dev)
                state, _ = ops.env_step(nat, state, o["actions"][:, 0].contiguous(), want_cost=False)
            outs[slot]["stats"].copy_(mpc_stats[slot])
            outs[slot]["stats"][:, 0] -= 1      # keep the "iteration index" convention: problem-iterations = stats[:,0] + 1

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3) if args.workload != "c5" else 1):
        step(0)
    for i in range(S):             # warm every stream's workspace (the caching allocator keeps one pool per stream)
        with torch.cuda.stream(streams[i]):
            step(i)
        if use_async:
            ops.ilqr_solve_async(nat, x0, u0, outs[i], works[i], done[i], opts)
    final_gather()                 # warm NCCL's all-gather path (connection set-up happens on the first call)
    barrier()
    sampler = ClockSampler(local) if rank == 0 and not args.no_clock_sampler else None
    t_wall0 = time.time()

    # ---- (1) sequential: one batch at a time, L2 flushed between steps -> per-batch latency
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    barrier()
    for s in range(args.steps):
        flush.zero_()                      # evict L2 between timed iterations (not timed)
        ev[s][0].record()
        step(0)
        ev[s][1].record()
    barrier()
    step_ms = [a.elapsed_time(b) for a, b in ev]
    seq_ms = float(sum(step_ms))
    queue_counters = None
    try:    # scheduling counters of the last sequential solve (queue solver only; the tick path leaves other data there)
        ws_main = _native._WS_CACHE.get((dev, main.cuda_stream))
        if ws_main is not None and args.workload == "c3" and os.environ.get("TFMPC_SOLVER", "queue") != "ticks":
            queue_counters = ops.queue_counters(ws_main)
    except Exception as exc:  # noqa: BLE001
        queue_counters = {"error": str(exc)[:100]}

    # ---- (2) pipelined: the same K independent batches issued round-robin on S streams, so the latency-bound tail of
    #      one batch (few unconverged problems) overlaps the throughput-bound head of the next
    launches0 = _native.kernel_launch_count("f32")
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record(main)
    for st in streams:
        st.wait_event(e0)
    tl = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_issue0 = time.perf_counter()
    if use_async:
        # ONE stream, K back-to-back tfmpc_ilqr_solve_async calls over a ring of S output/workspace slots: the heads run one
        # after another on `main`, each batch's stragglers on the library's priority streams; `done[slot]` guards slot reuse.
        for k in range(args.steps):
            slot = k % S
            if k >= S:
                main.wait_event(done[slot])       # outputs and workspace of this slot are free again
            ops.ilqr_solve_async(nat, x0, u0, outs[slot], works[slot], done[slot], opts)
        for k in range(max(0, args.steps - S), args.steps):
            main.wait_event(done[k % S])
    elif args.issue == "threads" and S > 1:      # one issuing host thread per stream (ctypes drops the GIL inside the C call)
        def issue(i):
            torch.cuda.set_device(local)
            with torch.cuda.stream(streams[i]):
                for k in range(i, args.steps, S):
                    if args.timeline:
                        tl[k][0].record()
                    step(i, k)
                    if args.timeline:
                        tl[k][1].record()
        workers = [threading.Thread(target=issue, args=(i,)) for i in range(S)]
        for th in workers:
            th.start()
        for th in workers:
            th.join()
    else:
        for s in range(args.steps):
            with torch.cuda.stream(streams[s % S]):
                if args.timeline:
                    tl[s][0].record()
                step(s % S, s)
                if args.timeline:
                    tl[s][1].record()
    for st in streams:
        fin = torch.cuda.Event()
        fin.record(st)
        main.wait_event(fin)
    issue_ms = 1e3 * (time.perf_counter() - t_issue0)      # host time spent enqueueing the K steps (no synchronisation inside)
    if world > 1:       # the S batches resident in the output ring (every batch solves the same inputs) -> [S, B] totals
        for i in range(S):
            torch.sum(outs[i]["costs"], dim=1, out=totals_all[i])
            iters_all[i].copy_(outs[i]["stats"][:, 0])
    e_mid = torch.cuda.Event(enable_timing=True)
    e_mid.record(main)
    final_gather()      # on `main`, behind every stream's last batch, inside the timed region
    e1.record(main)
    barrier()
    pipe_ms = float(e0.elapsed_time(e1))
    gather_ms = float(e_mid.elapsed_time(e1))
    timeline = [[k % S, round(e0.elapsed_time(a), 3), round(e0.elapsed_time(b), 3)] for k, (a, b) in enumerate(tl)] if args.timeline else None
    launches = _native.kernel_launch_count("f32") - launches0
    t_wall1 = time.time()
    clocks = sampler.stop(t_wall0, t_wall1) if sampler else None

    stats = outs[0]["stats"].cpu().numpy()
    pi_local = float((stats[:, 0] + 1).sum())
    t = torch.tensor([pipe_ms, seq_ms, pi_local, gather_ms, issue_ms / args.steps], dtype=torch.float64, device=dev)
    per_rank = None
    if world > 1:
        allt = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(allt, t)
        per_rank = [[round(float(a[0]) / args.steps, 4), round(float(a[1]) / args.steps, 4), float(a[2]), round(float(a[3]), 3), round(float(a[4]), 3)] for a in allt]
        tmax = t.clone(); dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        tsum = t.clone(); dist.all_reduce(tsum, op=dist.ReduceOp.SUM)
        pipe_ms, seq_ms, pi_all = float(tmax[0]), float(tmax[1]), float(tsum[2])
    else:
        pi_all = pi_local
    total_ms = pipe_ms if S > 1 else seq_ms
    value = pi_all * args.steps / (total_ms * 1e-3)

    # ---- end to end through the public host-buffer API: pinned host inputs -> results in host memory, every step.
    #      S host threads, each with its own env handle and stream (the C ABI is re-entrant across streams).
    e2e_steps = max(S, min(args.steps, 8 * S))
    nats = [envs.make_env(cfg).native() for _ in range(S)]
    houts = [new_out(host=True) for _ in range(S)]

    if args.workload == "c5":
        u_pins = [u.cpu().pin_memory() for u in u_inits]
        applied = [torch.empty(B, T, m).pin_memory() for _ in range(S)]

    def e2e_worker(i, count):
        torch.cuda.set_device(local)
        with torch.cuda.stream(streams[i]):
            for _ in range(count):
                if args.workload == "c5":   # closed loop: inputs from pinned host memory, applied actions back to the host
                    state = x0_pin.to(dev, non_blocking=True)
                    acts = []
                    for t in range(T):
                        o = ops.ilqr_solve(nats[i], state, u_pins[t].to(dev, non_blocking=True), opts)
                        acts.append(o["actions"][:, 0])
                        state, _ = ops.env_step(nats[i], state, o["actions"][:, 0].contiguous(), want_cost=False)
                    applied[i].copy_(torch.stack(acts, dim=1), non_blocking=True)
                    torch.cuda.current_stream().synchronize()
                else:
                    ops.ilqr_solve_host(nats[i], x0_pin, u0_pin, opts, houts[i])

    for i in range(S):
        e2e_worker(i, 1)                   # warm (allocates each handle's cached device scratch)
    barrier()
    counts = [e2e_steps // S + (1 if i < e2e_steps % S else 0) for i in range(S)]
    threads = [threading.Thread(target=e2e_worker, args=(i, counts[i])) for i in range(S)]
    t0 = time.perf_counter()
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    barrier()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = pi_all * e2e_steps / float(te[0])
    h2d = x0_pin.numel() * 4 + u0_pin.numel() * 4
    d2h = sum(v.numel() * 4 for v in houts[0].values())
    if args.workload == "c5":
        h2d = x0_pin.numel() * 4 + sum(u.numel() * 4 for u in u_pins)
        d2h = applied[0].numel() * 4

    if rank == 0:
        peaks, peak_src = measured_peaks()
        lib = _native.load("f32")
        tf, kms = ctypes.c_double(), ctypes.c_double()
        lib.tfmpc_measure_fp32_peak(ctypes.byref(tf), ctypes.byref(kms))
        flops, byts = algorithmic_work(args.workload, T if args.workload != "c5" else (T + 1) / 2.0, stats)
        solve_s = total_ms * 1e-3 / args.steps             # device time per solve (launch sequence), pipelined if S > 1
        ach_tf, ach_gb = flops / solve_s / 1e12, byts / solve_s / 1e9
        fp32 = {"bound": "fp32", "achieved": ach_tf, "peak": tf.value, "unit": "TFLOP/s", "frac": ach_tf / tf.value if tf.value else None,
                "traffic": None, "peak_source": "FP32 FMA microbenchmark run in this process (tfmpc_measure_fp32_peak); nominal 148 SM x 128 lanes x 2 x clock"}
        hbm = {"bound": "hbm", "achieved": ach_gb, "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": ach_gb / peaks["hbm_gbs"], "traffic": None,
               "peak_source": peak_src}
        tpath = os.path.join(ROOT, "profiles", f"r01_traffic_{args.workload}.json")
        if os.path.exists(tpath) and not args.batch:   # DRAM bytes of one solve, measured once with ncu (provenance inside the file)
            tr = json.load(open(tpath))
            hbm["traffic"] = fp32["traffic"] = tr["dram_bytes_per_solve"]
            hbm["traffic_source"] = fp32["traffic_source"] = tr["source"]
        primary = fp32 if (fp32["frac"] or 0) >= hbm["frac"] else hbm
        roofline = dict(primary)
        roofline["kernel"] = ("launch sequence of one solve: k_tick_backward + k_tick_linesearch per tick (thread-per-problem)"
                              if args.workload == "c3" else "kw_solve (lane-per-state, persistent)" + (" x 48 plant steps" if args.workload == "c5" else ""))
        roofline["algorithmic_flops_per_solve"] = flops
        roofline["algorithmic_bytes_per_solve"] = byts
        roofline["other"] = hbm if primary is fp32 else fp32
        line = {"metric": "batched iLQR problem-iterations/sec", "value": value, "unit": "problem-iterations/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": total_ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": desc, "batch_per_gpu": B, "global_batch": B * world, "horizon": T, "state_dim": n, "action_dim": m,
                           "parallelism": f"batch sharded over {world} GPU(s), no data-path collective; one NCCL all-gather of the resident per-problem costs and iteration counts at the end of the timed region"
                                          + ("; one NCCL all-gather of per-problem costs per step" if world > 1 else ""),
                           "streams": S,
                           "pipelining": ((f"the {args.steps} steps are independent batches issued back to back on ONE stream with tfmpc_ilqr_solve_async over a "
                                           f"ring of {S} output/workspace slots (straggler ticks on the library's priority streams)") if use_async else
                                          f"the {args.steps} steps are independent batches issued round-robin on {S} CUDA streams" if S > 1
                                          else "one batch at a time"),
                           "solver": "reference defaults atol=5e-3 max_iterations=100 mu_min=1e-6 delta_0=2 c1=0 alpha_min=1e-3",
                           "l2": ("pipelined: aggregate working set of the concurrent batches (S x ~420 MB) >> 126 MB L2; "
                                  "sequential: 256 MB buffer written between timed iterations (untimed L2 flush)"),
                           "mean_iterations_per_solve": float((stats[:, 0] + 1).mean()),
                           "problems_per_s": B * world * args.steps / (total_ms * 1e-3),
                           "status_histogram": np.bincount(stats[:, 3], minlength=5).tolist()},
                "sequential": {"value": pi_all * args.steps / (seq_ms * 1e-3), "unit": "problem-iterations/s",
                               "latency_ms_per_batch": seq_ms / args.steps, "step_ms": step_ms},
                "per_rank": ({"columns": ["pipelined ms/step", "sequential ms/batch", "problem-iterations per batch", "final all-gather ms (incl. waiting for the slowest rank)", "host enqueue ms/step"], "ranks": per_rank}
                             if per_rank else None),
                "host_enqueue_ms_per_step": issue_ms / args.steps,
                "queue_counters": queue_counters,
                "pipeline_timeline_ms": {"columns": ["stream", "start", "end"], "steps": timeline},
                "roofline": roofline,
                "e2e": {"value": e2e_value, "unit": "problem-iterations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "steps": e2e_steps, "host_threads": S,
                        "api": ("tfmpc_ilqr_solve_host (pinned host buffers in, host buffers out, synchronous), one call per step" if args.workload != "c5"
                                else "MPC loop: x0 and every step's initial actions copied from pinned host memory, applied actions copied back")},
                "gpu_launches": int(launches), "clocks": clocks}
        if world == 1 and not args.no_cpu_baseline:
            sample = args.cpu_sample or {"c3": 65536, "c4": 512, "c5s": 128, "c5": 128}[args.workload]
            v, cores, pi, dt = cpu_baseline(args.workload, T, sample)
            line["cpu_baseline"] = {"value": v, "unit": "problem-iterations/s", "cores": cores, "kind": "port",
                                    "sample": f"{sample} problems of the same workload ({pi:.0f} problem-iterations in {dt:.1f} s), "
                                              "C/OpenMP restatement of the reference algorithm (oracle/), one problem per thread"}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=0, help="timed steps (default: 64 for the GPU arm -- the pipeline needs a few waves of 8 "
                    "batches to reach its steady state --, 5 for the CPU reference arm)")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="c3", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="problems per GPU (default: the workload's BASELINE batch)")
    ap.add_argument("--cpu-sample", type=int, default=0, help="problems in the CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--timeline", action="store_true", help="record start/end events around every pipelined step")
    ap.add_argument("--no-clock-sampler", action="store_true", help="diagnostics only: a line without `clocks` is not a valid bench line")
    ap.add_argument("--pipeline", default="streams", choices=["async", "streams"],
                    help="how the K timed batches are kept in flight: async = one stream + tfmpc_ilqr_solve_async, streams = S streams")
    ap.add_argument("--issue", default="single", choices=["threads", "single"], help="host threads issuing the pipelined steps")
    ap.add_argument("--streams", type=int, default=8, help="CUDA streams the independent steps are pipelined over (1 = strictly sequential)")
    args = ap.parse_args()
    if args.steps <= 0:
        args.steps = 5 if (args.impl == "reference" or args.workload == "c5") else 64
    if args.impl == "reference":
        run_reference(args)
    elif args.workload == "c2":
        run_c2(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
