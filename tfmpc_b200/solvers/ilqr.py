"""Control-limited iLQR (Tassa, Mansard & Todorov 2014) -- mirror of the reference class
tfmpc/solvers/ilqr.py:22-387: same constructor kwargs and defaults, same stage methods
(start / derivatives / backward / forward) and solve(x0, T) -> (Trajectory, iteration).

Differences, all additive:
  * x0 may carry a leading batch axis ([B,n] or [B,n,1]); B problems are then solved in one
    launch sequence and solve returns (BatchTrajectory, iterations[B]).
  * `u_init=` / `seed=` on start() and solve() pin the initial action sequence, which the
    reference draws from an unseeded RNG (ilqr.py:69-70).
  * the whole outer loop (mu/delta schedule, convergence tests, line search) runs on the device.
"""
import logging
import os
import warnings
from collections import namedtuple

import numpy as np
import torch

from .. import _native as N
from .. import ops
from ..envs.diffenv import CostApprox, FinalCostApprox, TransitionApprox
from ..utils import trajectory

_Nominal = namedtuple("_Nominal", "states actions")


class iLQR:

    def __init__(self, env, **kwargs):
        self.env = env
        # solve (ilqr.py:27-29)
        self.atol = kwargs.get("atol", 5e-3)
        self.max_iterations = kwargs.get("max_iterations", 100)
        # backward (:31-33)
        self.mu_min = kwargs.get("mu_min", 1e-6)
        self.delta_0 = kwargs.get("delta_0", 2.0)
        # forward (:35-37)
        self.c1 = kwargs.get("c1", 0.0)
        self.alpha_min = kwargs.get("alpha_min", 1e-3)
        self.dtype = kwargs.get("dtype", torch.float32)
        self._config = kwargs
        if "logdir" in self._config:
            os.makedirs(self._config["logdir"], exist_ok=True)
            logging.basicConfig(filename=os.path.join(self._config["logdir"], "trace.log"), level=logging.DEBUG)

    # -- reference properties (:45-51)
    @property
    def low(self):
        return torch.as_tensor(self.env.action_space.low)

    @property
    def high(self):
        return torch.as_tensor(self.env.action_space.high)

    # -- plumbing
    def _native(self):
        return self.env.native(self.dtype)

    def _opts(self):
        return ops.make_opts(self.atol, self.max_iterations, self.mu_min, self.delta_0, self.c1, self.alpha_min)

    def _dev(self):
        N.require_cuda()
        return torch.device("cuda", torch.cuda.current_device())

    def _x0(self, x0, batched=None):
        """x0 as [B, n] plus "was a single problem".  Shapes are read the reference's way: [n,1] / [n] is one problem, [B,n] / [B,n,1]
        a batch.  For n = 1 the shape [1,1] is ambiguous (one column vector, or a batch of one): it is taken as ONE problem unless
        `batched=True` says otherwise (`batched=False` forces the single-problem reading of [n] / [n,1])."""
        n = self.env.state_size
        t = torch.as_tensor(np.asarray(x0) if not torch.is_tensor(x0) else x0).to(device=self._dev(), dtype=self.dtype)
        if batched is True:
            return t.reshape(-1, n).contiguous(), False
        if t.dim() >= 2 and t.shape[-1] == 1 and t.shape[-2] == n:
            t = t.squeeze(-1)
        single = t.dim() == 1
        if batched is False and not single:
            raise N.TfmpcError(f"x0 of shape {tuple(t.shape)} is not a single problem of state size {n}")
        return t.reshape(-1, n).contiguous(), single

    def _traj(self, t, size):
        """[T,size,1] / [T,size] / [B,T,size(,1)] -> ([B,T,size], was_single)"""
        t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t).to(device=self._dev(), dtype=self.dtype)
        if t.dim() >= 2 and t.shape[-1] == 1 and t.shape[-2] == size:
            t = t.squeeze(-1)
        single = t.dim() == 2
        return (t.unsqueeze(0) if single else t).contiguous(), single

    def initial_actions(self, batch, T, seed=None):
        """The reference's random start: ONE U(0,1) scalar per step, scaled to [low, high] in every
        action dimension, +-inf bounds replaced by +-1 (ilqr.py:59-70), drawn by the library
        on the device (tfmpc_ilqr_initial_actions: Philox keyed by `seed`; unseeded = a fresh seed from the OS)."""
        if seed is None:
            seed = int.from_bytes(os.urandom(8), "little")
        return ops.ilqr_initial_actions(self._native(), batch, T, seed, self.dtype, self._dev())

    # -- stages --------------------------------------------------------------------------
    def start(self, x0, T, u_init=None, seed=None):
        """ilqr.py:53-82 -> (states [T+1,n,1], actions [T,m,1], costs [T+1]) (leading B when batched)"""
        x0r, single = self._x0(x0)
        u = self.initial_actions(x0r.shape[0], T, seed) if u_init is None else self._traj(u_init, self.env.action_size)[0]
        xs, us, cs = ops.ilqr_start(self._native(), x0r, u)
        xs, us = xs.unsqueeze(-1), us.unsqueeze(-1)
        return (xs[0], us[0], cs[0]) if single else (xs, us, cs)

    def derivatives(self, states, actions):
        """ilqr.py:84-92 -> (TransitionApprox, CostApprox, FinalCostApprox) along the trajectory"""
        xs, single = self._traj(states, self.env.state_size)
        us, _ = self._traj(actions, self.env.action_size)
        B, T = us.shape[0], us.shape[1]
        n, m = self.env.state_size, self.env.action_size
        env = self._native()
        xr, ur = xs[:, :-1].reshape(B * T, n), us.reshape(B * T, m)
        lin = ops.env_linearize(env, xr, ur)
        f, _ = ops.env_step(env, xr, ur, want_cost=False)
        fq = ops.env_final_quad(env, xs[:, -1].contiguous())
        r = lambda a, *s: a.reshape(B, T, *s)  # noqa: E731
        tm = TransitionApprox(r(f, n, 1), r(lin["f_x"], n, n), r(lin["f_u"], n, m))
        cm = CostApprox(r(lin["l"]), r(lin["l_x"], n, 1), r(lin["l_u"], m, 1), r(lin["l_xx"], n, n), r(lin["l_uu"], m, m),
                        r(lin["l_ux"], m, n), r(lin["l_xu"], n, m))
        fm = FinalCostApprox(fq["l"], fq["l_x"].unsqueeze(-1), fq["l_xx"])
        if single:
            tm = TransitionApprox(*[a[0] for a in tm])
            cm = CostApprox(*[a[0] for a in cm])
            fm = FinalCostApprox(*[a[0] for a in fm])
        return tm, cm, fm

    def backward(self, T, actions, transition_model=None, cost_model=None, final_cost_model=None, mu=1.0, states=None):
        """ilqr.py:94-172 -> (K [T,m,n], k [T,m,1], J, dV1, dV2).

        With the reference's full signature (derivative models given) the generic dense kernel consumes exactly those
        models (tfmpc_ilqr_backward_staged: one warp per problem, TMA-staged blocks, all three controllers).  Called
        with `states=` instead of models, the environment-specialised kernel re-linearises analytically in registers
        (the path solve() uses)."""
        us, single = self._traj(actions, self.env.action_size)
        if transition_model is not None and cost_model is not None and final_cost_model is not None:
            B, T_ = us.shape[0], us.shape[1]
            dev = lambda a: torch.as_tensor(a).to(device=us.device, dtype=self.dtype)  # noqa: E731
            lift = (lambda a: dev(a).unsqueeze(0)) if single else dev
            tm = [lift(a) for a in transition_model]
            cm = [lift(a) for a in cost_model]
            fm = [lift(a) for a in final_cost_model]
            out = ops.ilqr_backward_staged(us, tm, cm, fm, self.env.action_space.low, self.env.action_space.high, float(mu))
        else:
            if states is None:
                raise N.TfmpcError("iLQR.backward needs the derivative models (reference signature) or states=")
            xs, _ = self._traj(states, self.env.state_size)
            out = ops.ilqr_backward(self._native(), xs, us, float(mu))
        K, k = out["K"], out["k"].unsqueeze(-1)
        if single:
            return K[0], k[0], out["J"][0], out["dV1"][0], out["dV2"][0]
        return K, k, out["J"], out["dV1"], out["dV2"]

    def forward(self, x, u, K, k, alpha=1.0):
        """ilqr.py:174-212 -> (states, actions, costs, J, residual)"""
        xs, single = self._traj(x, self.env.state_size)
        us, _ = self._traj(u, self.env.action_size)
        Kt = torch.as_tensor(K).to(device=xs.device, dtype=self.dtype)
        Kt = (Kt.unsqueeze(0) if Kt.dim() == 3 else Kt).contiguous()
        kt, _ = self._traj(k, self.env.action_size)
        out = ops.ilqr_forward(self._native(), xs, us, Kt, kt, float(alpha))
        s, a = out["states"].unsqueeze(-1), out["actions"].unsqueeze(-1)
        if single:
            return s[0], a[0], out["costs"][0], out["J"][0], out["residual"][0]
        return s, a, out["costs"], out["J"], out["residual"]

    # -- solve ---------------------------------------------------------------------------
    def solve_device(self, x0, T, u_init=None, seed=None, batched=None):
        """Batched solve, results left on the device: dict(states [B,T+1,n], actions [B,T,m],
        costs [B,T+1], stats [B,4] = iteration, backward passes, rollouts, status)."""
        x0r, _ = self._x0(x0, batched)
        u = self.initial_actions(x0r.shape[0], T, seed) if u_init is None else self._traj(u_init, self.env.action_size)[0]
        if u.shape[0] != x0r.shape[0]:
            raise N.TfmpcError(f"u_init batch {u.shape[0]} != x0 batch {x0r.shape[0]}")
        return ops.ilqr_solve(self._native(), x0r, u, self._opts())

    def solve(self, x0, T, show_progress=True, u_init=None, seed=None, batched=None):
        """ilqr.py:214-283 -> (Trajectory, iteration) [single problem] or (BatchTrajectory, iterations[B])."""
        _, single = self._x0(x0, batched)
        out = self.solve_device(x0, int(T), u_init=u_init, seed=seed, batched=batched)
        stats = out["stats"].cpu().numpy()
        logging.info(f"[SOLVE] mean iterations={stats[:, 0].mean():.2f} status={np.bincount(stats[:, 3], minlength=5).tolist()}")
        if single:
            if int(stats[0, 3]) not in (0, 1):      # not converged / max_iterations: the reference would have raised or exited (ilqr.py:305-313)
                warnings.warn(f"iLQR.solve stopped with status '{N.STATUS.get(int(stats[0, 3]), stats[0, 3])}' at iteration {int(stats[0, 0])}; "
                              "the trajectory returned is the last nominal", RuntimeWarning, stacklevel=2)
            return trajectory.Trajectory(out["states"][0], out["actions"][0], out["costs"][0]), int(stats[0, 0])
        return (trajectory.BatchTrajectory(out["states"], out["actions"], out["costs"], iterations=stats[:, 0], status=stats[:, 3]),
                stats[:, 0].copy())
