#!/usr/bin/env python
"""Generate the golden fixtures in this directory by running the UNMODIFIED
reference (thiagopbueno/tf-mpc v0.7.0, /root/reference) under the torch-backed
TensorFlow API shim in oracle/tf_shim (TensorFlow itself is not installable in
the build container).  Run from the repo root, in the BUILD container only:

    python tests/golden/make_golden.py            # float32 + float64 fixtures

The script re-executes itself once per precision (the shim's float type is fixed
at import).  Nothing here is imported by the product package, and the GPU box
never runs it (it has no /root/reference): tests read only the committed .npz.

What each fixture holds (all arrays; `env_json` is the reference's env JSON):
  lqr_*.npz      F f C c x0 T -> states actions costs K k V v const   (lqr.py:59-166)
  boxqp.npz      H q low high x0 -> x free clamped Hfree(padded)      (optimization.py:6-127)
  env_*.npz      x u -> f f_x f_u l l_x l_u l_xx l_uu l_ux l_xu, final l l_x l_xx (diffenv.py:13-101)
  stage_*.npz    x0 u_init mu alpha -> start rollout, backward K k J dV1 dV2, forward x u c J residual
  solve_*.npz    x0 u_init -> states actions costs iterations + per-call trace
  retry_*.npz    the same for problems whose first Cholesky fails (iLQR._backward's retry wrapper, ilqr.py:285-315): the trace
                 also records every FAILED backward call (kind 2) with the regularisation it was tried at
"""
import json
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"


def _reexec():
    for prec in ("float32", "float64"):
        env = dict(os.environ, TF_SHIM_FLOAT=prec, TFMPC_GOLDEN_CHILD="1",
                   PYTHONPATH=os.pathsep.join([os.path.join(ROOT, "oracle", "tf_shim"), REF, ROOT]))
        subprocess.check_call([sys.executable, os.path.abspath(__file__)] + sys.argv[1:], env=env)


if os.environ.get("TFMPC_GOLDEN_CHILD") != "1":
    _reexec()
    sys.exit(0)

import numpy as np  # noqa: E402
import tensorflow as tf  # noqa: E402  (the shim)

from tfmpc import envs as ref_envs  # noqa: E402
from tfmpc.envs.lqr import navigation as ref_navlqr  # noqa: E402
from tfmpc.solvers import ilqr as ref_ilqr  # noqa: E402
from tfmpc.solvers.lqr import LQR as RefLQR  # noqa: E402
from tfmpc.utils import optimization as ref_opt  # noqa: E402

from tfmpc_b200.envs import synthetic  # noqa: E402  (pure-python config generators, no CUDA)

PREC = os.environ["TF_SHIM_FLOAT"]
SUF = "f32" if PREC == "float32" else "f64"
NP = np.float32 if PREC == "float32" else np.float64


def npy(t):
    return np.asarray(t.numpy() if hasattr(t, "numpy") else t)


def save(name, **arrays):
    path = os.path.join(HERE, f"{name}_{SUF}.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", os.path.relpath(path, ROOT), f"({len(arrays)} arrays)")


# ------------------------------------------------------------------ LQR
def run_lqr(solver, x0, T):
    policy, value_fn = solver.backward(T)
    states, actions, costs = solver.forward(policy, tf.constant(x0), T)
    return dict(
        F=npy(solver.F), f=npy(solver.f), C=npy(solver.C), c=npy(solver.c), x0=x0, T=np.int32(T),
        states=npy(states), actions=npy(actions), costs=npy(costs).reshape(-1),
        K=np.stack([npy(K) for K, _ in policy]), k=np.stack([npy(k) for _, k in policy]),
        V=np.stack([npy(V) for V, _, _ in value_fn]), v=np.stack([npy(v) for _, v, _ in value_fn]),
        const=np.array([float(c) for _, _, c in value_fn], dtype=NP))


def gen_lqr():
    np.random.seed(0)
    solver = ref_envs.make_lqr(3, 2)                       # BASELINE config C1 (README example shape)
    save("lqr_c1", **run_lqr(solver, np.array([[-1.0], [0.5], [3.6]], dtype=NP), 10))
    rng = np.random.RandomState(1)
    for i, (n, m, T) in enumerate([(2, 2, 10), (4, 3, 7), (9, 9, 12), (5, 1, 3), (1, 4, 5), (8, 2, 20)]):
        np.random.seed(100 + i)
        solver = ref_envs.make_lqr(n, m)
        save(f"lqr_rand{i}", **run_lqr(solver, rng.normal(size=(n, 1)).astype(NP), T))
    # navlin: README.md:74-92 command line  `tfmpc navlin -b 5.0 -hr 10 -- "0.0 0.0" "8.0 -9.0"`
    g = np.array([[8.0], [-9.0]], dtype=np.float32)
    solver = ref_envs.make_lqr_linear_navigation(g, 5.0)
    save("lqr_navlin", **run_lqr(solver, np.zeros((2, 1), dtype=NP), 10))


# ------------------------------------------------------------------ box-QP
def gen_boxqp():
    rng = np.random.RandomState(7)
    cases = []
    kats = [([0.0, 0.0], [-1.0, 0.5], [1.0, 1.0]), ([0.0, 0.0], [0.5, -1.0], [1.0, 1.0]),
            ([1.0, 1.0], [0.0, 1.5], [2.0, 2.0]), ([1.0, 1.0], [1.5, 0.0], [2.0, 2.0]),
            ([0.0, 0.0, 0.0], [-1.0, 0.5, -1.0], [1.0, 1.0, 1.0]),
            ([0.0, 0.0, 0.0], [-1.0, 0.5, 0.30], [1.0, 1.0, 1.0])]     # tests/test_utils_optimization.py:7-15
    for goal, low, high in kats:
        d = len(goal)
        H = 2 * np.eye(d)
        q = -2 * np.array(goal).reshape(d, 1)
        lo, hi = np.array(low).reshape(d, 1), np.array(high).reshape(d, 1)
        starts = [(lo + hi) / 2, lo.copy(), hi.copy()] + [lo + rng.uniform(size=(d, 1)) * (hi - lo) for _ in range(3)]
        for x0 in starts:
            cases.append((H, q, lo, hi, x0))
    for d in (1, 2, 2, 2, 3, 5, 8, 8):
        for _ in range(6):
            A = rng.normal(size=(d, d))
            H = A @ A.T + 0.5 * np.eye(d)
            q = rng.normal(size=(d, 1)) * 3
            lo = -rng.uniform(0.1, 1.5, size=(d, 1))
            hi = rng.uniform(0.1, 1.5, size=(d, 1))
            cases.append((H, q, lo, hi, (lo + hi) / 2))
    M = 8
    out = {k: [] for k in ("dim", "H", "q", "low", "high", "x0", "x", "free", "clamped", "Hfree", "nfree")}
    for H, q, lo, hi, x0 in cases:
        d = H.shape[0]
        x, Hfree, free, clamped = ref_opt.projected_newton_qp(
            tf.constant(H.astype(NP)), tf.constant(q.astype(NP)), tf.constant(lo.astype(NP)),
            tf.constant(hi.astype(NP)), tf.constant(x0.astype(NP)))

        def pad(a, shape):
            r = np.zeros(shape, dtype=NP)
            a = np.asarray(a, dtype=NP)
            r[tuple(slice(0, s) for s in a.shape)] = a
            return r
        out["dim"].append(d)
        out["H"].append(pad(H, (M, M))); out["q"].append(pad(q, (M, 1)))
        out["low"].append(pad(lo, (M, 1))); out["high"].append(pad(hi, (M, 1))); out["x0"].append(pad(x0, (M, 1)))
        out["x"].append(pad(npy(x), (M, 1)))
        out["free"].append(pad(npy(free).astype(NP), (M, 1))); out["clamped"].append(pad(npy(clamped).astype(NP), (M, 1)))
        hf = npy(Hfree)
        out["nfree"].append(hf.shape[0]); out["Hfree"].append(pad(hf, (M, M)))
    save("boxqp", **{k: np.array(v) for k, v in out.items()})


# ------------------------------------------------------------------ envs
def load_json(rel):
    with open(os.path.join(REF, rel)) as fh:
        return json.load(fh)


def env_configs():
    cfgs = {
        "nav": load_json("tfmpc/envs/navigation/nav.config.json"),
        "res4": load_json("tfmpc/envs/reservoir/res4.config.json"),
        "hvac6": load_json("tfmpc/envs/hvac/hvac6.config.json"),
        "res20": synthetic.reservoir_config(20),
        "hvac32": synthetic.hvac_grid_config(4, 8),
        "nav1": {"module": "navigation", "cls_name": "Navigation",
                 "config": {"goal": [[3.0], [-2.0]],
                            "deceleration": {"center": [[[1.0], [-0.5]]], "decay": [2.5]},
                            "low": [[-0.5], [-1.0]], "high": [[1.0], [0.75]]},
                 "initial_state": [[0.0], [0.0]]},
    }
    return cfgs


def make_ref_env(cfg):
    cfg = json.loads(json.dumps(cfg))          # HVAC.load mutates its argument
    if cfg["cls_name"] == "NavigationLQR":     # navlin.config.json names a module that does not exist
        return ref_navlqr.NavigationLQR.load(cfg["config"])
    return ref_envs.make_env(cfg)


def navlqr_cfg(goal, beta, low=None, high=None):
    c = {"goal": [[float(g)] for g in goal], "beta": float(beta)}
    if low is not None:
        c["low"], c["high"] = float(low), float(high)
    return {"module": "lqr.navigation", "cls_name": "NavigationLQR", "config": c,
            "initial_state": [[0.0]] * len(goal)}


def sample_xu(name, env, rng, T):
    n, m = env.state_size, env.action_size
    if name.startswith("res"):
        lb, ub = npy(env.lower_bound), npy(env.upper_bound)
        x = lb + rng.uniform(-0.3, 1.3, size=(T, n, 1)) * (ub - lb)
        u = rng.uniform(0, 1, size=(T, m, 1))
    elif name.startswith("hvac"):
        x = rng.normal(loc=18.0, scale=5.0, size=(T, n, 1))
        u = rng.uniform(0, 1, size=(T, m, 1))
    else:
        x = rng.uniform(-3, 9, size=(T, n, 1))
        u = rng.uniform(-1, 1, size=(T, m, 1))
    return x.astype(NP), u.astype(NP)


def gen_envs():
    rng = np.random.RandomState(11)
    cfgs = env_configs()
    cfgs["navlqr"] = navlqr_cfg([5.5, -9.0], 5.0, -1.0, 1.0)
    cfgs["navlqr3"] = navlqr_cfg([1.0, -2.0, 3.0], 0.5)
    for name, cfg in cfgs.items():
        env = make_ref_env(cfg)
        T = 6
        x, u = sample_xu(name, env, rng, T)
        xt, ut = tf.constant(x), tf.constant(u)
        tm = env.get_linear_transition(xt, ut, batch=True)
        cm = env.get_quadratic_cost(tf.constant(x), tf.constant(u), batch=True)
        fm = env.get_quadratic_final_cost(tf.constant(x[-1]))
        nxt = env.transition(tf.constant(x), tf.constant(u), batch=True)
        cost = env.cost(tf.constant(x), tf.constant(u), batch=True)
        fcost = np.array([float(env.final_cost(tf.constant(x[t]))) for t in range(T)], dtype=NP)
        save(f"env_{name}", env_json=json.dumps(cfg), x=x, u=u, next=npy(nxt), cost=npy(cost).reshape(-1), final_cost=fcost,
             f=npy(tm.f), f_x=npy(tm.f_x), f_u=npy(tm.f_u),
             l=npy(cm.l).reshape(-1), l_x=npy(cm.l_x), l_u=npy(cm.l_u), l_xx=npy(cm.l_xx), l_uu=npy(cm.l_uu),
             l_ux=npy(cm.l_ux), l_xu=npy(cm.l_xu),
             fl=npy(fm.l).reshape(()), fl_x=npy(fm.l_x), fl_xx=npy(fm.l_xx))


# ------------------------------------------------------------------ iLQR stages and solves
def rollout(env, x0, u_init):
    """What iLQR.start (ilqr.py:53-82) does once its random actions are fixed to u_init."""
    T = u_init.shape[0]
    state = tf.constant(x0)
    states, costs = [state], []
    for t in range(T):
        action = tf.constant(u_init[t])
        costs.append(tf.reshape(env.cost(state, action), []))
        state = env.transition(state, action)
        states.append(state)
    costs.append(tf.reshape(env.final_cost(state), []))
    return tf.stack(states), tf.constant(u_init), tf.stack(costs)


def make_u_init(env, rng, T):
    low, high = env.action_space.low, env.action_space.high
    lo = np.where(np.isinf(low), -1.0, low).reshape(-1, 1)
    hi = np.where(np.isinf(high), 1.0, high).reshape(-1, 1)
    r = rng.uniform(size=(T, 1, 1))                         # one scalar per step (ilqr.py:70)
    return (lo + r * (hi - lo)).astype(NP)


def stage_cases():
    cfgs = env_configs()
    return [
        ("navlqr_free", navlqr_cfg([5.5, -9.0], 5.0), [0.0, 0.0], 10),
        ("navlqr_box", navlqr_cfg([5.5, -9.0], 5.0, -1.0, 1.0), [0.0, 0.0], 10),
        ("navlqr_box_b0", navlqr_cfg([5.5, -9.0], 0.0, -1.0, 1.0), [0.0, 0.0], 10),
        ("nav", cfgs["nav"], [0.0, 0.0], 50),
        ("nav1", cfgs["nav1"], [0.3, 0.2], 12),
        ("res4", cfgs["res4"], [75.0, 50.0, 50.0, 50.0], 40),
        ("hvac6", cfgs["hvac6"], [10.0] * 6, 48),
    ]


def gen_stages():
    rng = np.random.RandomState(21)
    for name, cfg, x0, T in stage_cases():
        env = make_ref_env(cfg)
        solver = ref_ilqr.iLQR(env)
        x0 = np.array(x0, dtype=NP).reshape(-1, 1)
        u_init = make_u_init(env, rng, T)
        xh, uh, ch = rollout(env, x0, u_init)
        tm, cm, fm = solver.derivatives(xh, uh)
        out = dict(env_json=json.dumps(cfg), x0=x0, u_init=u_init, T=np.int32(T),
                   states0=npy(xh), costs0=npy(ch))
        for tag, mu in (("mu0", 0.0), ("mu1", 1.0), ("mu3", 1e-3)):
            K, k, J, dV1, dV2 = solver.backward(T, uh, tm, cm, fm, tf.constant(mu, dtype=tf.float32))
            out.update({f"K_{tag}": npy(K), f"k_{tag}": npy(k), f"J_{tag}": np.array(float(J), dtype=NP),
                        f"dV1_{tag}": np.array(float(dV1), dtype=NP), f"dV2_{tag}": np.array(float(dV2), dtype=NP)})
            if tag == "mu0":
                K0, k0 = K, k
        for i, alpha in enumerate(np.geomspace(1.0, 1e-3, 11)[[0, 3, 10]]):
            xs, us, cs, J, res = solver.forward(xh, uh, K0, k0, tf.constant(alpha, dtype=tf.float32))
            out.update({f"alpha_{i}": np.array(alpha, dtype=NP), f"fx_{i}": npy(xs), f"fu_{i}": npy(us), f"fc_{i}": npy(cs),
                        f"fJ_{i}": np.array(float(J), dtype=NP), f"fres_{i}": np.array(float(res), dtype=NP)})
        save(f"stage_{name}", **out)


def traced_solve(env, x0, u_init, **kw):
    solver = ref_ilqr.iLQR(env, **kw)
    T = u_init.shape[0]
    start = rollout(env, x0, u_init)
    solver.start = lambda x0_, T_: start                    # bypass the unseeded random init only
    trace = []
    bwd, fwd = solver.backward, solver.forward

    def backward(T_, u, tm, cm, fm, mu):
        try:
            out = bwd(T_, u, tm, cm, fm, mu)
        except tf.errors.InvalidArgumentError:
            trace.append([2.0, float(mu), 0.0, 0.0, 0.0])      # a failed pass: _backward bumps its LOCAL mu / delta and retries (:305-309)
            raise
        trace.append([0.0, float(mu), float(out[2]), float(out[3]), float(out[4])])
        return out

    def forward(x, u, K, k, alpha):
        out = fwd(x, u, K, k, alpha)
        trace.append([1.0, float(alpha), float(out[3]), float(out[4]), 0.0])
        return out
    solver.backward, solver.forward = backward, forward
    traj, it = solver.solve(tf.constant(x0), T, show_progress=False)
    return traj, it, np.array(trace, dtype=np.float64)


def solve_cases():
    cfgs = env_configs()
    cases = []
    for beta in (0.0, 5.0):
        cases.append((f"navlqr_free_b{int(beta)}", navlqr_cfg([5.5, -9.0], beta), [[0.0, 0.0]], 10, 2))
        cases.append((f"navlqr_box_b{int(beta)}", navlqr_cfg([5.5, -9.0], beta, -1.0, 1.0), [[0.0, 0.0]], 10, 2))
    cases.append(("navlqr3_box", navlqr_cfg([1.0, -2.0, 3.0], 0.5, -0.4, 0.6), [[0.0, 0.5, -0.5]], 8, 2))
    nav_x0 = [[0.0, 0.0], [-1.5, 1.0], [1.9, -1.9], [0.7, 1.3], [-2.0, -2.0], [1.0, 0.2]]
    cases.append(("nav_h50", cfgs["nav"], nav_x0, 50, 1))
    cases.append(("nav_h20", cfgs["nav"], nav_x0[:3], 20, 1))
    cases.append(("nav1_h12", cfgs["nav1"], [[0.3, 0.2], [-1.0, 1.0]], 12, 1))
    cases.append(("res4_h40", cfgs["res4"], [[75.0, 50.0, 50.0, 50.0], [30.0, 100.0, 200.0, 300.0]], 40, 1))
    cases.append(("hvac6_h48", cfgs["hvac6"], [[10.0] * 6, [12.0, 9.0, 10.5, 11.0, 8.0, 10.0]], 48, 1))
    cases.append(("res20_h40", cfgs["res20"], [None], 40, 1))
    cases.append(("hvac32_h48", cfgs["hvac32"], [None], 48, 1))
    return cases


def gen_solves():
    rng = np.random.RandomState(31)
    for name, cfg, x0s, T, reps in solve_cases():
        env = make_ref_env(cfg)
        n = env.state_size
        X0, U0, S, A, Cc, IT, TR, TRN = [], [], [], [], [], [], [], []
        for x0 in x0s:
            for _ in range(reps):
                if x0 is None:
                    x0v = synthetic.sample_x0(cfg, 1, rng)[0]
                else:
                    x0v = np.array(x0)
                x0v = x0v.astype(NP).reshape(n, 1)
                u_init = make_u_init(env, rng, T)
                traj, it, trace = traced_solve(env, x0v, u_init)
                X0.append(x0v); U0.append(u_init); S.append(traj.states); A.append(traj.actions)
                Cc.append(traj.costs); IT.append(it); TRN.append(len(trace))
                pad = np.zeros((4096, 5)); pad[:len(trace)] = trace[:4096]; TR.append(pad)
                print(f"  {name}: iterations={it} total={traj.total_cost:.6f} calls={len(trace)}")
        mx = max(TRN)
        save(f"solve_{name}", env_json=json.dumps(cfg), x0=np.array(X0), u_init=np.array(U0), T=np.int32(T),
             states=np.array(S), actions=np.array(A), costs=np.array(Cc), iterations=np.array(IT, dtype=np.int32),
             trace=np.array(TR)[:, :mx], trace_len=np.array(TRN, dtype=np.int32))


def gen_retry():
    """Row a16: unbounded NavigationLQR with beta < -1, so that Q_uu_reg = 2 beta + V_xx + mu is not positive definite at the
    last timestep for mu = 0 (terminal V_xx = 2 I) and the reference takes ilqr.py:305-309 until its local mu exceeds
    -(2 beta + 2).  The objective is unbounded below for beta < -1, so the runs are cut by max_iterations."""
    rng = np.random.RandomState(47)
    for name, beta, T, max_it in (("b1p0005", -1.0005, 6, 4), ("b1p2", -1.2, 5, 3), ("b3", -3.0, 4, 2)):
        cfg = navlqr_cfg([1.5, -0.5], beta)
        env = make_ref_env(cfg)
        X0, U0, S, A, Cc, IT, TR, TRN = [], [], [], [], [], [], [], []
        for x0 in ([0.0, 0.0], [1.0, -1.0], [0.3, 0.8]):
            x0v = np.array(x0).astype(NP).reshape(2, 1)
            u_init = (0.1 * rng.uniform(-1, 1, size=(T, 1, 1)) * np.ones((1, 2, 1))).astype(NP)
            traj, it, trace = traced_solve(env, x0v, u_init, max_iterations=max_it)
            X0.append(x0v); U0.append(u_init); S.append(traj.states); A.append(traj.actions)
            Cc.append(traj.costs); IT.append(it); TRN.append(len(trace))
            pad = np.zeros((4096, 5)); pad[:len(trace)] = trace[:4096]; TR.append(pad)
            print(f"  retry_{name}: iterations={it} total={traj.total_cost:.6g} calls={len(trace)} failed={int((trace[:, 0] == 2).sum())}")
        mx = max(TRN)
        save(f"retry_{name}", env_json=json.dumps(cfg), x0=np.array(X0), u_init=np.array(U0), T=np.int32(T), max_iterations=np.int32(max_it),
             states=np.array(S), actions=np.array(A), costs=np.array(Cc), iterations=np.array(IT, dtype=np.int32),
             trace=np.array(TR)[:, :mx], trace_len=np.array(TRN, dtype=np.int32))


if __name__ == "__main__":
    which = sys.argv[1:] or ["lqr", "boxqp", "envs", "stages", "solves", "retry"]
    for w in which:
        {"lqr": gen_lqr, "boxqp": gen_boxqp, "envs": gen_envs, "stages": gen_stages, "solves": gen_solves, "retry": gen_retry}[w]()
