"""HVAC: n-room thermal network, actions in [0,1], piecewise-linear cost.
Mirror of tfmpc/envs/hvac/__init__.py:8-195 (constants :10-15)."""
import numpy as np

from .diffenv import Box, DiffEnv
from .gymenv import GymEnv

_KEYS = ["temp_outside", "temp_hall", "temp_lower_bound", "temp_upper_bound", "R_outside", "R_hall", "capacity", "air_max",
         "adj_outside", "adj_hall"]


class HVAC(DiffEnv, GymEnv):  # the reference's HVAC is not a GymEnv, so its --online mode fails (quirk Q8)
    _kind = 3
    CAP_AIR, COST_AIR, TEMP_AIR, TIME_DELTA, PENALTY, SET_POINT_PENALTY = 1.006, 1.0, 40.0, 1.0, 20_000.0, 10.0

    def __init__(self, temp_outside, temp_hall, temp_lower_bound, temp_upper_bound, R_outside, R_hall, R_wall, capacity, air_max,
                 adj, adj_outside, adj_hall):
        GymEnv.__init__(self)
        col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1)  # noqa: E731
        self.temp_outside, self.temp_hall = col(temp_outside), col(temp_hall)
        self.temp_lower_bound, self.temp_upper_bound = col(temp_lower_bound), col(temp_upper_bound)
        self.R_outside, self.R_hall = col(R_outside), col(R_hall)
        n = self.temp_lower_bound.shape[0]
        self.R_wall = np.asarray(R_wall, dtype=np.float64).reshape(n, n)
        self.capacity, self.air_max = col(capacity), col(air_max)
        self.adj = np.asarray(adj, dtype=bool).reshape(n, n)
        self.adj_outside = np.asarray(adj_outside, dtype=bool).reshape(-1, 1)
        self.adj_hall = np.asarray(adj_hall, dtype=bool).reshape(-1, 1)
        self.obs_space = Box(shape=[n, 1], low=-np.inf, high=np.inf)
        self.action_space = Box(shape=[n, 1], low=0.0, high=1.0)

    @property
    def state_size(self):
        return len(self.temp_lower_bound)

    @property
    def action_size(self):
        return self.state_size

    def _pack(self):
        p = []
        for k in _KEYS:
            p += list(np.asarray(getattr(self, k), dtype=np.float64).reshape(-1))
        return 0, p + list(self.R_wall.reshape(-1)) + list(self.adj.astype(np.float64).reshape(-1))

    def __repr__(self):
        return f"HVAC({self.state_size})"

    def __str__(self):
        bounds = ", ".join(f"[{float(lo):.3f}, {float(hi):.3f}]" for lo, hi in zip(self.temp_lower_bound, self.temp_upper_bound))
        return (f"HVAC(\ntemp_bounds=[{bounds}],\nR_wall=\n{self.R_wall},\nadj=\n{self.adj},\n"
                f"adj_outside={self.adj_outside.squeeze().tolist()},\nadj_hall={self.adj_hall.squeeze().tolist()}\n)")

    @classmethod
    def load(cls, config):
        return cls(**config)
