"""Launchers -- mirror of tfmpc/launchers/__init__.py:12-51: env JSON -> solver -> solve -> data.csv."""
import json
import os

import numpy as np

from .. import agents, envs, runners
from ..solvers import ilqr


def _load(config):
    env_config = config.pop("env")
    if isinstance(env_config, (str, os.PathLike)):
        with open(env_config, "r") as file:
            env_config = json.load(file)
    env = envs.make_env(env_config)
    x0 = np.asarray(env_config["initial_state"], dtype=np.float32)
    T = int(config.pop("horizon"))
    return env, x0, T


def _save(traj, config, num_samples):
    if "logdir" not in config:
        return
    if num_samples is None:
        traj.save(os.path.join(config["logdir"], "data.csv"))
    else:       # one run directory per sample, as the reference's experiment runner lays them out
        for i in range(len(traj)):
            traj[i].save(os.path.join(config["logdir"], f"run{i}", "data.csv"))


def ilqr_run(config, num_samples=None):
    """launchers/__init__.py:12-29.  num_samples = N solves N samples (same x0, N independent random initial action sequences, the
    reference's `--num-samples` runs) as ONE batch on the GPU and returns a BatchTrajectory."""
    config = dict(config)
    env, x0, T = _load(config)
    solver = ilqr.iLQR(env, **config)
    if num_samples is not None:
        x0 = np.repeat(x0.reshape(1, -1), int(num_samples), axis=0)
    trajectory, iterations = solver.solve(x0, T, seed=config.get("seed"))
    _save(trajectory, config, num_samples)
    return env, trajectory


def online_ilqr_run(config, num_samples=None):
    """launchers/__init__.py:32-51.  num_samples = N runs N closed loops (N plants with independent noise and initial actions) as one
    batch: every plant step is one batched solve and one batched plant step."""
    config = dict(config)
    env, x0, T = _load(config)
    solver = ilqr.iLQR(env, **config)
    env.seed(config.get("seed"))       # the plant is the stochastic one (GymEnv.step -> transition(cec=False), gymenv.py:18)
    controller = agents.MPC(solver, T, seed=config.get("seed"))
    runner = runners.Runner(env, controller)
    if num_samples is not None:
        x0 = np.repeat(x0.reshape(1, -1), int(num_samples), axis=0)
    with runner(x0, T) as r:
        trajectory = r.run()
        _save(trajectory, config, num_samples)
    return env, trajectory
