"""Navigation: 2-D nonlinear navigation with sigmoid deceleration zones and bounded actions.
Mirror of the reference class tfmpc/envs/navigation/__init__.py:9-96 (same constructor, properties,
`load`), computed by the CUDA kernels behind DiffEnv."""
import numpy as np

from .diffenv import Box, DiffEnv
from .gymenv import GymEnv


class Navigation(DiffEnv, GymEnv):
    _kind = 1

    def __init__(self, goal, deceleration, low, high):
        GymEnv.__init__(self)
        self.goal = np.asarray(goal, dtype=np.float64).reshape(-1, 1)
        self.deceleration = {"center": np.asarray(deceleration["center"], dtype=np.float64).reshape(-1, 2, 1),
                             "decay": np.asarray(deceleration["decay"], dtype=np.float64).reshape(-1)}
        self.obs_space = Box(low=np.array([-np.inf, -np.inf]), high=np.array([np.inf, np.inf]))
        self.action_space = Box(low=np.array(low, dtype=np.float32), high=np.array(high, dtype=np.float32))

    @property
    def action_size(self):
        return self.state_size

    @property
    def state_size(self):
        return self.goal.shape[0]

    def _pack(self):
        nz = len(self.deceleration["decay"])
        p = (list(self.goal.reshape(-1)) + list(self.action_space.low.reshape(-1)) + list(self.action_space.high.reshape(-1))
             + list(self.deceleration["center"].reshape(-1)) + list(self.deceleration["decay"]))
        return nz, p

    def __repr__(self):
        goal = self.goal.squeeze().tolist()
        bounds = f"[{self.action_space.low.squeeze().tolist()}, {self.action_space.high.squeeze().tolist()}]"
        decay = ", ".join(f"{d:.4f}" for d in self.deceleration["decay"].tolist())
        return f"Navigation(goal={goal}, deceleration={{center={self.deceleration['center'].tolist()}, decay=[{decay}]}}, bounds={bounds})"

    @classmethod
    def load(cls, config):
        return cls(config["goal"], dict(config["deceleration"]), config["low"], config["high"])
