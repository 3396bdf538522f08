"""NavigationLQR: linear navigation x' = x + u with cost ||x-g||^2 + beta ||u||^2 and an optional
action box.  Mirror of tfmpc/envs/lqr/navigation/__init__.py:8-68."""
import numpy as np

from ..diffenv import Box, DiffEnv
from ..gymenv import GymEnv


class NavigationLQR(DiffEnv, GymEnv):
    _kind = 0

    def __init__(self, goal, beta, low=None, high=None):
        GymEnv.__init__(self)
        self.goal = np.asarray(goal.detach().cpu() if hasattr(goal, "detach") else goal, dtype=np.float64).reshape(-1, 1)
        self.beta = float(beta)
        low = -np.inf if low is None else low
        high = np.inf if high is None else high
        shape = self.goal.shape
        self.obs_space = Box(-np.inf, np.inf, shape=shape)
        self.action_space = Box(low, high, shape=shape)

    @property
    def action_size(self):
        return self.state_size

    @property
    def state_size(self):
        return self.goal.shape[0]

    def _pack(self):
        p = list(self.goal.reshape(-1)) + [self.beta] + list(self.action_space.low.reshape(-1).astype(np.float64)) \
            + list(self.action_space.high.reshape(-1).astype(np.float64))
        return 0, p

    @classmethod
    def load(cls, config):
        return cls(np.asarray(config["goal"], dtype=np.float64).reshape(-1, 1), config["beta"], config.get("low"), config.get("high"))

    def __repr__(self):
        bounds = ""
        if self.action_space.is_bounded():
            bounds = f", bounds=[{self.action_space.low.squeeze().tolist()}, {self.action_space.high.squeeze().tolist()}]"
        return f"NavigationLQR(goal={self.goal.squeeze().tolist()}, beta={self.beta}{bounds})"
