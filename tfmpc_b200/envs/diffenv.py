"""DiffEnv: the differentiable-environment interface of the reference
(tfmpc/envs/diffenv.py:6-101) backed by analytic CUDA linearisation kernels instead of
GradientTape autodiff."""
import os
from collections import namedtuple

import numpy as np
import torch

from .. import _native as N
from .. import ops

TransitionApprox = namedtuple("TransitionApprox", "f f_x f_u")
CostApprox = namedtuple("CostApprox", "l l_x l_u l_xx l_uu l_ux l_xu")
FinalCostApprox = namedtuple("FinalCostApprox", "l l_x l_xx")  # the reference names it "CostApprox" (quirk Q13)


class Box:
    """The slice of gym.spaces.Box the reference uses: low, high, is_bounded()."""

    def __init__(self, low, high, shape=None):
        if shape is not None:
            low = np.full(shape, low, dtype=np.float32) if np.isscalar(low) else np.broadcast_to(np.asarray(low, np.float32), shape).copy()
            high = np.full(shape, high, dtype=np.float32) if np.isscalar(high) else np.broadcast_to(np.asarray(high, np.float32), shape).copy()
        self.low = np.asarray(low, dtype=np.float32)
        self.high = np.asarray(high, dtype=np.float32)
        self.shape = self.low.shape

    def is_bounded(self):
        return bool(np.all(np.isfinite(self.low)) and np.all(np.isfinite(self.high)))


def _device():
    N.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class DiffEnv:
    """Base class.  Subclasses define `_kind`, `state_size`, `action_size`, `action_space` and
    `_pack()` -> (nz, [float params]) in the layout documented in include/tfmpc_b200.h."""

    _kind = None
    dtype = torch.float32

    # -- native handle -------------------------------------------------------------------
    def native(self, dtype=None):
        dtype = dtype or self.dtype
        prec = "f32" if dtype == torch.float32 else "f64"
        cache = self.__dict__.setdefault("_native_envs", {})
        if prec not in cache:
            nz, params = self._pack()
            cache[prec] = N.Env(prec, self._kind, self.state_size, self.action_size, nz, params)
        return cache[prec]

    # -- tensor plumbing -----------------------------------------------------------------
    def _rows(self, t, size, dtype=None):
        """[size,1] / [size] / [R,size,1] / [R,size] -> ([R,size] CUDA tensor, restore-shape fn)"""
        dtype = dtype or self.dtype
        t = torch.as_tensor(np.asarray(t) if not torch.is_tensor(t) else t)
        t = t.to(device=_device(), dtype=dtype)
        col = t.dim() >= 2 and t.shape[-1] == 1 and t.shape[-2] == size
        if col:
            t = t.squeeze(-1)
        single = t.dim() == 1
        rows = t.reshape(-1, size)
        return rows, single, col

    # -- reference API -------------------------------------------------------------------
    def transition(self, state, action, batch=False, cec=True):
        """cec=True: the deterministic dynamics the planner uses.  cec=False: the plant -- the environment's noise model on
        (navigation/__init__.py:45, reservoir/__init__.py:98-105), drawn on the device; every call advances the draw counter,
        `seed()` (GymEnv) restarts it.  Environments without a noise model ignore the flag."""
        x, single, col = self._rows(state, self.state_size)
        u, _, _ = self._rows(action, self.action_size)
        if cec:
            xn, _ = ops.env_step(self.native(), x, u, want_cost=False)
        else:
            if getattr(self, "_noise_seed", None) is None:
                self._noise_seed, self._noise_calls = int.from_bytes(os.urandom(8), "little"), 0
            xn, _ = ops.env_step_noisy(self.native(), x, u, self._noise_seed, self._noise_calls, want_cost=False)
            self._noise_calls += 1
        xn = xn[0] if single else xn
        return xn.unsqueeze(-1) if col else xn

    def cost(self, state, action, batch=False):
        x, single, _ = self._rows(state, self.state_size)
        u, _, _ = self._rows(action, self.action_size)
        _, c = ops.env_step(self.native(), x, u, want_next=False)
        return c[0] if single else c

    def final_cost(self, state):
        x, single, _ = self._rows(state, self.state_size)
        c = ops.env_final_cost(self.native(), x)
        return c[0] if single else c

    def get_linear_transition(self, state, action, batch=True):
        """diffenv.py:13-32 -> TransitionApprox(f [T,n,1], f_x [T,n,n], f_u [T,n,m]) (batch=False drops T)"""
        x, single, col = self._rows(state, self.state_size)
        u, _, _ = self._rows(action, self.action_size)
        env = self.native()
        lin = ops.env_linearize(env, x, u)
        f, _ = ops.env_step(env, x, u, want_cost=False)
        f = f.unsqueeze(-1)
        if single:
            return TransitionApprox(f[0], lin["f_x"][0], lin["f_u"][0])
        return TransitionApprox(f, lin["f_x"], lin["f_u"])

    def get_quadratic_cost(self, state, action, batch=True):
        """diffenv.py:34-83 -> CostApprox(l [T], l_x [T,n,1], l_u [T,m,1], l_xx, l_uu, l_ux, l_xu)"""
        x, single, col = self._rows(state, self.state_size)
        u, _, _ = self._rows(action, self.action_size)
        lin = ops.env_linearize(self.native(), x, u)
        vals = [lin["l"], lin["l_x"].unsqueeze(-1), lin["l_u"].unsqueeze(-1), lin["l_xx"], lin["l_uu"], lin["l_ux"], lin["l_xu"]]
        if single:
            vals = [v[0] for v in vals]
        return CostApprox(*vals)

    def get_quadratic_final_cost(self, state):
        """diffenv.py:85-101 -> FinalCostApprox(l [], l_x [n,1], l_xx [n,n])"""
        x, single, col = self._rows(state, self.state_size)
        q = ops.env_final_quad(self.native(), x)
        vals = [q["l"], q["l_x"].unsqueeze(-1), q["l_xx"]]
        if single:
            vals = [v[0] for v in vals]
        return FinalCostApprox(*vals)
