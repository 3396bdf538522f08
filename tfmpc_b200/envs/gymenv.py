"""GymEnv: the plant-side interface of the reference's MPC loop (tfmpc/envs/gymenv.py:5-41): setup / reset / step.

`step` advances the PLANT: like the reference (gymenv.py:18) it calls transition(state, action, cec=False), i.e. the
environment's noise model is on -- truncated-normal position noise for Navigation, gamma rainfall for Reservoir, drawn on the
device (tfmpc_env_step_noisy, Philox keyed by env.seed()).  NavigationLQR and HVAC have no noise model (in the reference their
transition() does not even accept `cec`, so GymEnv.step raises there; here they step deterministically).
Set `env.cec_plant = True` for a certainty-equivalent (noise-free) plant, e.g. to replay a closed loop exactly."""


class GymEnv:

    cec_plant = False

    def __init__(self):
        self._t = None
        self._state = None
        self._info = {}

    def setup(self, initial_state, horizon):
        self.initial_state = initial_state
        self.horizon = int(horizon)

    def step(self, action):
        self._t += 1
        next_state = self.transition(self._state, action, cec=bool(self.cec_plant))
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
        """Seeds the plant noise (the reference's seed() is a stub: TensorFlow's global generator is used there)."""
        import os
        self._noise_seed = int.from_bytes(os.urandom(8), "little") if seed is None else int(seed)
        self._noise_calls = 0
        return [self._noise_seed]
