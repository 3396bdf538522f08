#!/usr/bin/env python
# coding: utf-8
"""`tfmpc` command line -- same commands and flags as the reference's scripts/tfmpc.py:26-215
(`tfmpc lqr`, `tfmpc navlin`, `tfmpc ilqr ENV`).  The reference fans `--num-samples` runs out over tuneconfig worker
processes (scripts/tfmpc.py:203-210); here the samples are a batch axis solved in one launch sequence on the GPU,
so `--num-workers` is accepted and ignored."""
import json
import logging
import os
import sys

import click
import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tfmpc_b200 import envs  # noqa: E402
from tfmpc_b200.launchers import ilqr_run, online_ilqr_run  # noqa: E402


@click.group()
def cli():
    pass


def _verbosity(debug, verbose):
    logging.basicConfig(level=logging.DEBUG if debug else (logging.INFO if verbose else logging.ERROR))


@cli.command()
@click.argument("initial-state")
@click.option("--action-size", "-a", type=click.IntRange(min=1), default=1, help="The number of action variables.")
@click.option("--horizon", "-hr", type=click.IntRange(min=1), default=10, help="The number of timesteps.")
@click.option("--debug", is_flag=True, help="Debug flag.")
@click.option("--verbose", "-v", is_flag=True, help="Verbosity flag.")
def lqr(initial_state, action_size, horizon, debug, verbose):
    """Generate and solve a randomly-created LQR problem.

    Args:

        initial_state: list of floats.
    """
    _verbosity(debug, verbose)
    initial_state = list(map(float, initial_state.split()))
    x0 = np.array(initial_state, dtype=np.float32)[:, np.newaxis]
    solver = envs.make_lqr(len(initial_state), action_size)
    trajectory = solver.solve(x0, horizon)
    print(repr(trajectory))
    print()
    print(str(trajectory))


@cli.command()
@click.argument("initial-state")
@click.argument("goal")
@click.option("--beta", "-b", type=float, default=1.0, help="The weight of the action cost.")
@click.option("--horizon", "-hr", type=click.IntRange(min=1), default=10, help="The number of timesteps.")
@click.option("--debug", is_flag=True, help="Debug flag.")
@click.option("--verbose", "-v", is_flag=True, help="Verbosity flag.")
def navlin(initial_state, goal, beta, horizon, debug, verbose):
    """Generate and solve the linear navigation LQR problem.

    Args:

        initial_state: list of floats.

        goal: list of floats.
    """
    _verbosity(debug, verbose)
    x0 = np.array(list(map(float, initial_state.split())), dtype=np.float32)[:, np.newaxis]
    g = np.array(list(map(float, goal.split())), dtype=np.float32)[:, np.newaxis]
    solver = envs.make_lqr_linear_navigation(g, beta)
    trajectory = solver.solve(x0, horizon)
    print(repr(trajectory))
    print()
    print(str(trajectory))


@cli.command()
@click.argument("env", type=click.Path(exists=True))
@click.option("--online", is_flag=True, help="Online mode flag.", show_default=True)
@click.option("--horizon", "-hr", type=click.IntRange(min=1), default=10, help="The number of timesteps.", show_default=True)
@click.option("--atol", type=click.FloatRange(min=0.0), default=5e-3, help="Absolute tolerance for convergence.", show_default=True)
@click.option("--max-iterations", "-miter", type=click.IntRange(min=1), default=100, help="Maximum number of iterations.", show_default=True)
@click.option("--logdir", type=click.Path(), default="/tmp/ilqr/", help="Directory used for logging results.", show_default=True)
@click.option("--num-samples", "-ns", type=click.IntRange(min=1), default=1, help="Number of runs.", show_default=True)
@click.option("--num-workers", "-nw", type=click.IntRange(min=1), default=1, help="Accepted for compatibility; runs are batched on the GPU.", show_default=True)
@click.option("--seed", type=int, default=None, help="Seed of the initial action sequences (the reference's are unseeded).")
@click.option("--verbose", "-v", count=True, help="Verbosity level flag.")
def ilqr(**kwargs):
    """Run iLQR for a given environment and horizon.

    Args:

        ENV: Path to the environment's config JSON file.
    """
    verbose = kwargs.pop("verbose")
    logging.basicConfig(level={1: logging.INFO, 2: logging.DEBUG}.get(verbose, logging.ERROR))
    online = kwargs.pop("online")
    num_samples = kwargs.pop("num_samples")
    kwargs.pop("num_workers")
    seed = kwargs.pop("seed")
    exec_func = online_ilqr_run if online else ilqr_run
    config = dict(kwargs)
    if seed is not None:
        config["seed"] = seed
    # the reference fans the samples out over worker processes (tuneconfig); here they are ONE batch on the GPU
    _, batch = exec_func(config, num_samples=num_samples)
    for run_id in range(num_samples):
        trajectory = batch[run_id]
        print(repr(trajectory))
        print(str(trajectory))


if __name__ == "__main__":
    cli()
