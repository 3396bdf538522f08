"""Linear Quadratic Regulator -- mirror of the reference class tfmpc/solvers/lqr.py:16-181
(same constructor, properties and methods), solving a BATCH of problems per CUDA launch.

Single-problem use is unchanged:  LQR(F, f, C, c).solve(x0 [n,1], T) -> Trajectory.
Batched use: give any of F, f, C, c a leading batch axis and/or pass x0 as [B,n] / [B,n,1];
`solve` then returns a BatchTrajectory.
"""
import ctypes as C
import json

import numpy as np
import torch

from .. import _native as N
from .. import ops
from ..utils import trajectory


def _dev():
    N.require_cuda()
    return torch.device("cuda", torch.cuda.current_device())


class LQR:

    def __init__(self, F, f, C, c, dtype=torch.float32):
        self.dtype = dtype
        to = lambda a: torch.as_tensor(np.asarray(a.detach().cpu() if torch.is_tensor(a) else a, dtype=np.float64)).to(dtype)  # noqa: E731
        # host copies in the reference's shapes: F [n,N], f [n,1], C [N,N], c [N,1] (+ optional leading B)
        self.F, self.f, self.C, self.c = to(F), to(f), to(C), to(c)
        self._dev_params = None

    # -- reference properties (lqr.py:24-34)
    @property
    def n_dim(self):
        return self.F.shape[-1]

    @property
    def state_size(self):
        return self.F.shape[-2]

    @property
    def action_size(self):
        return self.n_dim - self.state_size

    @property
    def batch_size(self):
        for t in (self.F, self.f, self.C, self.c):
            if t.dim() == 3:
                return t.shape[0]
        return None

    def _device_params(self):
        if self._dev_params is None:
            d = _dev()
            n, NN = self.state_size, self.n_dim
            F = self.F.to(d).contiguous()
            f = self.f.to(d).reshape(-1, n).contiguous() if self.f.dim() == 3 else self.f.to(d).reshape(n).contiguous()
            Cm = self.C.to(d).contiguous()
            c = self.c.to(d).reshape(-1, NN).contiguous() if self.c.dim() == 3 else self.c.to(d).reshape(NN).contiguous()
            self._dev_params = (F, f, Cm, c)
        return self._dev_params

    def _strides(self):
        F, f, Cm, c = self._device_params()
        n, NN = self.state_size, self.n_dim
        return (n * NN if F.dim() == 3 else 0, n if f.dim() == 2 else 0, NN * NN if Cm.dim() == 3 else 0, NN if c.dim() == 2 else 0)

    def _x(self, x, size, batched=None):
        """[size,1] / [size] is one problem, [B,size] / [B,size,1] a batch (the reference's shapes plus a batch axis).  For size = 1
        the shape [1,1] is ambiguous: it is read as ONE column vector unless `batched=True`."""
        t = torch.as_tensor(np.asarray(x) if not torch.is_tensor(x) else x).to(device=_dev(), dtype=self.dtype)
        if batched is True:
            return t.reshape(-1, size).contiguous(), False
        if t.dim() >= 2 and t.shape[-1] == 1 and t.shape[-2] == size:
            t = t.squeeze(-1)
        single = t.dim() == 1
        return t.reshape(-1, size).contiguous(), single

    def _lib(self):
        return N.load("f32" if self.dtype == torch.float32 else "f64")

    def _match_batch(self, t, what):
        """A batched solver (per-problem F / f / C / c) takes one row per problem, or a single row that is shared by all of them;
        anything else would index the parameters out of bounds on the device."""
        B = self.batch_size
        if B is None or t is None or t.shape[0] == B:
            return t
        if t.shape[0] == 1:
            return t.expand(B, *t.shape[1:]).contiguous()
        raise N.TfmpcError(f"{what}: {t.shape[0]} rows for a solver of {B} problems (expected {B}, or 1 to share)")

    def _step(self, x, u, want):
        n, m = self.state_size, self.action_size
        F, f, Cm, c = self._device_params()
        sF, sf, sC, sc = self._strides()
        xr, single = self._x(x, n)
        ur = self._x(u, m)[0] if u is not None else None
        xr, ur = self._match_batch(xr, "state"), self._match_batch(ur, "action")
        if ur is not None and ur.shape[0] != xr.shape[0]:
            raise N.TfmpcError(f"{ur.shape[0]} actions for {xr.shape[0]} states")
        single = single and xr.shape[0] == 1
        R = xr.shape[0]
        lib = self._lib()
        outs = {k: (torch.empty((R, n) if k == "next" else (R,), dtype=self.dtype, device=xr.device) if k == want else None)
                for k in ("next", "cost", "final")}
        P = lambda t: N.dev_ptr(lib, t)  # noqa: E731
        ptrs = [P(F), P(f), P(Cm), P(c), P(xr), P(ur), P(outs["next"]), P(outs["cost"]), P(outs["final"])]
        N.check(lib, lib.tfmpc_lqr_step(C.c_int64(R), n, m, ptrs[0].p, C.c_int64(sF), ptrs[1].p, C.c_int64(sf), ptrs[2].p, C.c_int64(sC),
                                        ptrs[3].p, C.c_int64(sc), ptrs[4].p, ptrs[5].p, ptrs[6].p, ptrs[7].p, ptrs[8].p, N.stream_ptr()))
        out = outs[want]
        if want == "next":
            out = out.unsqueeze(-1)
        return out[0] if single else out

    # -- reference methods
    def transition(self, x, u):   # lqr.py:36-39
        return self._step(x, u, "next")

    def cost(self, x, u):         # lqr.py:41-47
        return self._step(x, u, "cost")

    def final_cost(self, x):      # lqr.py:49-57
        return self._step(x, None, "final")

    def _solve(self, x0, T, terminal_zero=False, want_policy=True, want_value=True, batched=None):
        F, f, Cm, c = self._device_params()
        B = self.batch_size
        if x0 is None:
            x0 = torch.zeros(B or 1, self.state_size, dtype=self.dtype, device=F.device)
            single = B is None
        else:
            x0, single = self._x(x0, self.state_size, batched)
            single = single and B is None
            if B is not None and x0.shape[0] == 1:
                x0 = x0.expand(B, -1).contiguous()
        return ops.lqr_solve(F, f, Cm, c, x0, int(T), terminal_zero, want_policy, want_value), single

    def backward(self, T):
        """lqr.py:59-129 -> (policy, value_fn): lists over t of (K [m,n], k [m,1]) and (V [n,n], v [n,1], const)
        for a single problem; for a batched solver the entries carry a leading B axis."""
        out, single = self._solve(None, T)
        K, k, V, v, cst = out["K"], out["k"].unsqueeze(-1), out["V"], out["v"].unsqueeze(-1), out["const"]
        sel = (lambda t_, a: a[0, t_]) if single else (lambda t_, a: a[:, t_])
        policy = [(sel(t, K), sel(t, k)) for t in range(int(T))]
        value_fn = [(sel(t, V), sel(t, v), sel(t, cst)) for t in range(int(T))]
        return policy, value_fn

    def forward(self, policy, x0, T):
        """lqr.py:131-161: closed-loop rollout of a given policy -> (states [T+1,n,1], actions [T,m,1], costs [T+1])"""
        n, m = self.state_size, self.action_size
        F, f, Cm, c = self._device_params()
        sF, sf, sC, sc = self._strides()
        x0r, single = self._x(x0, n)
        x0r = self._match_batch(x0r, "x0")
        single = single and x0r.shape[0] == 1
        B = x0r.shape[0]
        T = int(T)
        K = torch.stack([torch.as_tensor(p[0]) for p in policy], dim=-3).to(device=x0r.device, dtype=self.dtype).reshape(-1, T, m, n)
        k = torch.stack([torch.as_tensor(p[1]).reshape(*torch.as_tensor(p[1]).shape[:-2], m) for p in policy], dim=-2)
        k = k.to(device=x0r.device, dtype=self.dtype).reshape(-1, T, m)
        if K.shape[0] == 1 and B > 1:
            K, k = K.expand(B, -1, -1, -1), k.expand(B, -1, -1)
        if K.shape[0] != B or k.shape[0] != B:
            raise N.TfmpcError(f"policy of {K.shape[0]} problems for {B} initial states")
        K, k = K.contiguous(), k.contiguous()
        lib = self._lib()
        states = torch.empty(B, T + 1, n, dtype=self.dtype, device=x0r.device)
        actions = torch.empty(B, T, m, dtype=self.dtype, device=x0r.device)
        costs = torch.empty(B, T + 1, dtype=self.dtype, device=x0r.device)
        P = lambda t: N.dev_ptr(lib, t)  # noqa: E731
        ptrs = [P(F), P(f), P(Cm), P(c), P(K), P(k), P(x0r), P(states), P(actions), P(costs)]
        N.check(lib, lib.tfmpc_lqr_forward(C.c_int64(B), n, m, T, ptrs[0].p, C.c_int64(sF), ptrs[1].p, C.c_int64(sf), ptrs[2].p,
                                           C.c_int64(sC), ptrs[3].p, C.c_int64(sc), ptrs[4].p, ptrs[5].p, ptrs[6].p, ptrs[7].p, ptrs[8].p,
                                           ptrs[9].p, N.stream_ptr()))
        states, actions = states.unsqueeze(-1), actions.unsqueeze(-1)
        if single:
            return states[0], actions[0], costs[0]
        return states, actions, costs

    def solve(self, x0, T, terminal_zero=False, batched=None):
        """lqr.py:163-166 -> Trajectory (single problem) or BatchTrajectory (`batched=True` forces the batch reading of x0)"""
        out, single = self._solve(x0, T, terminal_zero, want_policy=False, want_value=False, batched=batched)
        if single:
            if int(out["status"][0]) != 0:     # the reference raises here too (tf.linalg.inv of a singular Q_uu, lqr.py:84)
                raise N.TfmpcError("LQR.solve: Q_uu is singular at some timestep (status %d)" % int(out["status"][0]))
            return trajectory.Trajectory(out["states"][0], out["actions"][0], out["costs"][0])
        return trajectory.BatchTrajectory(out["states"], out["actions"], out["costs"], status=out["status"])

    def solve_device(self, x0, T, terminal_zero=False, want_policy=False, want_value=False):
        """Batched solve that leaves the results on the device (dict of CUDA tensors)."""
        return self._solve(x0, T, terminal_zero, want_policy, want_value)[0]

    # -- JSON interchange (lqr.py:168-181)
    def dump(self, file):
        json.dump({k: getattr(self, k).cpu().numpy().tolist() for k in ("F", "f", "C", "c")}, file)

    @classmethod
    def load(cls, file):
        config = json.load(file)
        return cls(**{k: np.array(v).astype("f") for k, v in config.items()})
