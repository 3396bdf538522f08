"""Result containers (host side).  Mirrors tfmpc/utils/trajectory.py:11-90 of the reference:
same properties, printer and CSV format; `BatchTrajectory` is the batch-of-problems extension."""
import os
from collections import namedtuple

import numpy as np

Transition = namedtuple("Transition", "state action cost")


def _np(a):
    if hasattr(a, "detach"):
        a = a.detach().cpu().numpy()
    elif hasattr(a, "numpy"):
        a = a.numpy()
    return np.asarray(a)


class Trajectory:
    """states [T+1,n(,1)], actions [T,m(,1)], costs [T+1].

    Reference quirk Q7: LQR.forward hands over costs shaped [T+1,1,1], which
    breaks the reference's own __str__ on current NumPy; costs are flattened to [T+1] here."""

    def __init__(self, states, actions, costs):
        states, actions, costs = _np(states), _np(actions), _np(costs)
        if states.ndim == 3 and states.shape[-1] == 1:
            states = np.squeeze(states, axis=-1)     # trajectory.py:13
        if actions.ndim == 3 and actions.shape[-1] == 1:
            actions = np.squeeze(actions, axis=-1)   # trajectory.py:14
        self.states = states
        self.actions = actions
        self.costs = costs.reshape(-1)

    @property
    def initial_state(self):
        return self.states[0]

    @property
    def final_state(self):
        return self.states[-1]

    @property
    def total_cost(self):
        return np.sum(self.costs)

    @property
    def cumulative_cost(self):
        return np.cumsum(self.costs)

    @property
    def cost_to_go(self):
        return np.cumsum(self.costs[::-1])[::-1]

    def __len__(self):
        return len(self.actions)

    def __getitem__(self, t):
        return Transition(self.states[t + 1], self.actions[t], self.costs[t])

    def __iter__(self):
        return (self[t] for t in range(len(self)))

    def __repr__(self):
        return f"Trajectory(init={self.initial_state}, final={self.final_state}, total={self.total_cost:.4f})"

    def __str__(self):
        rows = [("Steps", "States", "Actions", "Costs")]
        for t, (state, action, cost) in enumerate(self):
            state = "[" + ", ".join(f"{x:8.4f}" for x in state) + "]"
            action = "[" + ", ".join(f"{u:8.4f}" for u in action) + "]"
            rows.append((str(t), state, action, f"{cost:8.4f}"))
        sizes = [max(map(len, col)) for col in zip(*rows)]
        out = " | ".join(h.center(sz) for h, sz in zip(rows[0], sizes)) + "\n"
        out += " | ".join("=" * sz for sz in sizes) + "\n"
        for row in rows[1:]:
            out += " | ".join(col.center(sz) for col, sz in zip(row, sizes)) + "\n"
        return out

    def save(self, filepath):
        """CSV with columns x[1..n], u[1..m], costs and index Timestep (trajectory.py:71-90)."""
        import pandas as pd
        df = pd.DataFrame()
        for i, x_i in enumerate(np.transpose(self.states[1:])):
            df[f"x[{i+1}]"] = x_i
        for i, u_i in enumerate(np.transpose(self.actions)):
            df[f"u[{i+1}]"] = u_i
        df["costs"] = self.costs[:-1]
        dirname = os.path.dirname(filepath)
        if dirname and not os.path.exists(dirname):
            os.makedirs(dirname)
        df.to_csv(filepath, index=True, index_label="Timestep")


class BatchTrajectory:
    """B trajectories: states [B,T+1,n], actions [B,T,m], costs [B,T+1] (+ optional per-problem
    iteration counts and status codes).  Indexing yields the reference's single-problem Trajectory."""

    def __init__(self, states, actions, costs, iterations=None, status=None):
        self.states, self.actions, self.costs = _np(states), _np(actions), _np(costs)
        self.iterations = None if iterations is None else _np(iterations)
        self.status = None if status is None else _np(status)

    def __len__(self):
        return self.states.shape[0]

    def __getitem__(self, b):
        return Trajectory(self.states[b], self.actions[b], self.costs[b])

    @property
    def total_cost(self):
        return self.costs.sum(axis=1)

    def __repr__(self):
        return f"BatchTrajectory(B={len(self)}, T={self.actions.shape[1]}, mean_total={self.total_cost.mean():.4f})"
