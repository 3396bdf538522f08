"""GymEnv: the plant-side interface of the reference's MPC loop (tfmpc/envs/gymenv.py:5-41):
setup / reset / step.  `step` uses the deterministic dynamics unless the env implements
`_plant_noise`."""


class GymEnv:

    def __init__(self):
        self._t = None
        self._state = None
        self._info = {}

    def setup(self, initial_state, horizon):
        self.initial_state = initial_state
        self.horizon = int(horizon)

    def step(self, action):
        self._t += 1
        next_state = self.transition(self._state, action)
        noise = getattr(self, "_plant_noise", None)
        if noise is not None:
            next_state = noise(self._state, action, next_state)
        cost = self.cost(self._state, action)
        done = self._t == self.horizon
        self._state = next_state
        return next_state, cost, done, self._info

    def reset(self):
        self._t = 0
        self._state = self.initial_state
        self._info = {}
        return self._state

    def render(self, mode="human"):
        pass

    def close(self):
        pass

    def seed(self, seed=None):
        pass
