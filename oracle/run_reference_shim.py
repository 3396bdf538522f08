"""TEST / BENCH INFRASTRUCTURE -- times the UNMODIFIED reference (thiagopbueno/tf-mpc v0.7.0, installed with
`pip install --no-deps --target baseline/_ref`) on a few problems of a bench workload, under the torch-backed
TensorFlow API shim in oracle/tf_shim (TensorFlow itself is not installable offline).  Run as a subprocess by
bench.py --impl reference; prints one JSON line.

    PYTHONPATH=oracle/tf_shim:baseline/_ref python oracle/run_reference_shim.py <env.json> <inputs.npz>
"""
import json
import sys
import time

import numpy as np
import tensorflow as tf  # the shim

from tfmpc import envs as ref_envs
from tfmpc.solvers import ilqr as ref_ilqr


def rollout(env, x0, u_init):   # iLQR.start with its random actions pinned (ilqr.py:53-82)
    state = tf.constant(x0)
    states, costs = [state], []
    for t in range(u_init.shape[0]):
        action = tf.constant(u_init[t])
        costs.append(tf.reshape(env.cost(state, action), []))
        state = env.transition(state, action)
        states.append(state)
    costs.append(tf.reshape(env.final_cost(state), []))
    return tf.stack(states), tf.constant(u_init), tf.stack(costs)


def main():
    cfg = json.load(open(sys.argv[1]))
    d = np.load(sys.argv[2])
    x0, u0 = d["x0"], d["u0"]
    env = ref_envs.make_env(json.loads(json.dumps(cfg)))
    its, costs = [], []
    t0 = time.perf_counter()
    for b in range(x0.shape[0]):
        solver = ref_ilqr.iLQR(env)
        start = rollout(env, x0[b].reshape(-1, 1).astype(np.float32), u0[b][..., None].astype(np.float32))
        solver.start = lambda x0_, T_: start
        traj, it = solver.solve(tf.constant(x0[b].reshape(-1, 1).astype(np.float32)), u0.shape[1], show_progress=False)
        its.append(int(it) + 1)
        costs.append(float(traj.total_cost))
    dt = time.perf_counter() - t0
    print(json.dumps({"problems": int(x0.shape[0]), "problem_iterations": int(sum(its)), "seconds": dt, "costs": costs, "iterations": its}))


if __name__ == "__main__":
    main()
