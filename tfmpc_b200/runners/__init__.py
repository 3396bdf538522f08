"""Closed-loop runner -- mirror of tfmpc/runners/__init__.py:8-49."""
import contextlib

import torch

from ..utils import trajectory


class Runner:

    def __init__(self, env, agent):
        self.env = env
        self.agent = agent

    def run(self, mode=None):
        state = self.env.reset()
        timestep = 0
        done = False
        states, actions, costs = [torch.as_tensor(state)], [], []
        while not done:
            action = self.agent(state, timestep)
            next_state, cost, done, info = self.env.step(action)
            if mode is not None:
                self.env.render(mode)
            state = next_state
            timestep = self.env._t
            states.append(state)
            actions.append(action)
            costs.append(cost)
        costs.append(self.env.final_cost(state))
        dev = actions[0].device
        states = torch.stack([s.to(dev) for s in states])
        actions = torch.stack(actions)
        costs = torch.stack(costs)
        if states.dim() == 3 and states.shape[-1] != 1:     # batched plants: [T+1,B,n] -> [B,T+1,n]
            return trajectory.BatchTrajectory(states.transpose(0, 1), actions.transpose(0, 1), costs.transpose(0, 1))
        return trajectory.Trajectory(states, actions, costs)

    @contextlib.contextmanager
    def __call__(self, initial_state, horizon):
        self.env.setup(initial_state, horizon)
        yield self
        self.env.close()
