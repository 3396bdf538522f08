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


def ilqr_run(config):
    config = dict(config)
    env, x0, T = _load(config)
    solver = ilqr.iLQR(env, **config)
    trajectory, iterations = solver.solve(x0, T, seed=config.get("seed"))
    if "logdir" in config:
        trajectory.save(os.path.join(config["logdir"], "data.csv"))
    return env, trajectory


def online_ilqr_run(config):
    config = dict(config)
    env, x0, T = _load(config)
    solver = ilqr.iLQR(env, **config)
    env.seed(config.get("seed"))       # the plant is the stochastic one (GymEnv.step -> transition(cec=False), gymenv.py:18)
    controller = agents.MPC(solver, T, seed=config.get("seed"))
    runner = runners.Runner(env, controller)
    with runner(x0, T) as r:
        trajectory = r.run()
        if "logdir" in config:
            trajectory.save(os.path.join(config["logdir"], "data.csv"))
    return env, trajectory
