"""Reservoir: n-reservoir water network, actions in [0,1], piecewise-linear cost.
Mirror of tfmpc/envs/reservoir/__init__.py:9-129."""
import numpy as np

from .diffenv import Box, DiffEnv
from .gymenv import GymEnv

_KEYS = ["max_res_cap", "lower_bound", "upper_bound", "low_penalty", "high_penalty", "set_point_penalty", "rain_shape", "rain_scale"]


class Reservoir(DiffEnv, GymEnv):
    _kind = 2

    def __init__(self, max_res_cap, lower_bound, upper_bound, low_penalty, high_penalty, set_point_penalty, downstream,
                 rain_shape, rain_scale):
        GymEnv.__init__(self)
        col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1, 1)  # noqa: E731
        self.max_res_cap, self.lower_bound, self.upper_bound = col(max_res_cap), col(lower_bound), col(upper_bound)
        self.low_penalty, self.high_penalty, self.set_point_penalty = col(low_penalty), col(high_penalty), col(set_point_penalty)
        self.rain_shape, self.rain_scale = col(rain_shape), col(rain_scale)
        n = self.lower_bound.shape[0]
        self.downstream = np.asarray(downstream, dtype=np.float64).reshape(n, n)
        self.obs_space = Box(low=np.zeros_like(self.max_res_cap), high=self.max_res_cap)
        self.action_space = Box(shape=[n, 1], low=0.0, high=1.0)

    @property
    def state_size(self):
        return len(self.lower_bound)

    @property
    def action_size(self):
        return self.state_size

    def _pack(self):
        p = []
        for k in _KEYS:
            p += list(getattr(self, k).reshape(-1))
        return 0, p + list(self.downstream.reshape(-1))

    def __repr__(self):
        return f"Reservoir({self.state_size})"

    def __str__(self):
        bounds = ", ".join(f"[{float(l):.2f}, {float(u):.2f}]" for l, u in zip(self.lower_bound, self.upper_bound))
        rain = ", ".join(f"Gamma(shape={float(a):.2f}, scale={float(b):.2f})" for a, b in zip(self.rain_shape, self.rain_scale))
        return f"Reservoir(\nbounds={bounds},\ntopology=\n{self.downstream},\nrain={rain})"

    @classmethod
    def load(cls, config):
        return cls(**{k: np.asarray(v, dtype=np.float64) for k, v in config.items()})
