"""Shrinking-horizon MPC agent -- mirror of tfmpc/agents/mpc.py:4-15: at plant step t re-solve the
iLQR problem from the current state over the remaining horizon H - t (fresh initial actions every
step, no warm start, as in the reference) and apply the first action.  `state` may be [n,1] or a
batch [B,n]; the solve for all B plants is one launch sequence."""
import torch


class MPC:

    def __init__(self, solver, horizon, seed=None):
        self.solver = solver
        self.horizon = int(horizon)
        self.seed = seed
        self.iterations = []          # per plant step: iteration counts of the solve

    def __call__(self, state, timestep):
        steps_to_go = self.horizon - int(timestep)
        seed = None if self.seed is None else self.seed + int(timestep)
        out = self.solver.solve_device(state, steps_to_go, seed=seed)
        self.iterations.append(out["stats"][:, 0])
        action = out["actions"][:, 0]                       # trajectory[0].action, mpc.py:13
        st = torch.as_tensor(state)
        if st.dim() == 2 and st.shape[-1] == 1 and st.shape[0] == self.solver.env.state_size:
            return action[0].unsqueeze(-1)                  # reference shape [m,1]
        return action[0] if st.dim() == 1 else action
