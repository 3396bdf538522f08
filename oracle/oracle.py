"""TEST INFRASTRUCTURE ONLY -- ctypes front-end of the CPU oracle (oracle/tfmpc_oracle.c).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product package (tfmpc_b200/) never does.

    from oracle import oracle
    o = oracle.Oracle("f32")                # or "f64"
    env = o.make_env(env_config_dict)       # the reference's env JSON format
    out = o.ilqr_solve(env, x0[B,n], u_init[B,T,m])

Arrays follow the reference layouts: states [B,T+1,n], actions [B,T,m], costs [B,T+1].
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_BUILD = os.path.join(_HERE, "_build")

KIND = {"NavigationLQR": 0, "Navigation": 1, "Reservoir": 2, "HVAC": 3}


def build(force=False):
    """Compile liboracle_f32.so / liboracle_f64.so with the committed Makefile."""
    libs = [os.path.join(_BUILD, f"liboracle_{p}.so") for p in ("f32", "f64")]
    src = os.path.join(_HERE, "tfmpc_oracle.c")
    stale = force or any(not os.path.exists(p) or os.path.getmtime(p) < os.path.getmtime(src) for p in libs)
    if stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return libs


def pack_env(cfg):
    """Reference env JSON -> (kind, n, m, nz, float64 parameter vector); layout documented at
    oracle_env_create in tfmpc_oracle.c."""
    kind = KIND[cfg["cls_name"]]
    c = cfg["config"]
    col = lambda v: np.asarray(v, dtype=np.float64).reshape(-1)  # noqa: E731
    if kind == 0:
        goal = col(c["goal"])
        n = goal.size
        low = c.get("low")
        high = c.get("high")
        # gym.spaces.Box stores its bounds as float32 (reference lqr/navigation/__init__.py:21), whatever the solver precision
        low = np.full(n, -np.inf if low is None else float(np.float32(low)))
        high = np.full(n, np.inf if high is None else float(np.float32(high)))
        return kind, n, n, 0, np.concatenate([goal, [float(c["beta"])], low, high])
    if kind == 1:
        centers = np.asarray(c["deceleration"]["center"], dtype=np.float64).reshape(-1, 2)
        decay = col(c["deceleration"]["decay"])
        f32 = lambda v: col(v).astype(np.float32).astype(np.float64)  # noqa: E731  (Box bounds are float32, navigation/__init__.py:21-24)
        return kind, 2, 2, len(decay), np.concatenate([col(c["goal"]), f32(c["low"]), f32(c["high"]), centers.reshape(-1), decay])
    if kind == 2:
        keys = ["max_res_cap", "lower_bound", "upper_bound", "low_penalty", "high_penalty", "set_point_penalty",
                "rain_shape", "rain_scale"]
        n = col(c["lower_bound"]).size
        return kind, n, n, 0, np.concatenate([col(c[k]) for k in keys] + [col(c["downstream"])])
    keys = ["temp_outside", "temp_hall", "temp_lower_bound", "temp_upper_bound", "R_outside", "R_hall", "capacity",
            "air_max", "adj_outside", "adj_hall"]
    n = col(c["temp_lower_bound"]).size
    return kind, n, n, 0, np.concatenate([col(c[k]) for k in keys] + [col(c["R_wall"]), col(c["adj"])])


class Env:
    def __init__(self, lib, handle, n, m, cfg):
        self._lib, self.handle, self.n, self.m, self.cfg = lib, handle, n, m, cfg

    def __del__(self):
        if self.handle:
            self._lib.oracle_env_destroy(self.handle)
            self.handle = None


class Oracle:
    def __init__(self, precision="f32"):
        assert precision in ("f32", "f64")
        build()
        self.precision = precision
        self.dtype = np.float32 if precision == "f32" else np.float64
        self.lib = C.CDLL(os.path.join(_BUILD, f"liboracle_{precision}.so"))
        self.lib.oracle_env_create.restype = C.c_void_p
        self.lib.oracle_env_create.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        self.lib.oracle_env_destroy.argtypes = [C.c_void_p]
        self.lib.oracle_max_threads.restype = C.c_int

    # ------------------------------------------------------------ helpers
    def _a(self, x, shape=None):
        a = np.ascontiguousarray(np.asarray(x, dtype=self.dtype))
        return a if shape is None else np.ascontiguousarray(a.reshape(shape))

    @staticmethod
    def _p(a):
        return a.ctypes.data_as(C.c_void_p) if a is not None else None

    def max_threads(self):
        return int(self.lib.oracle_max_threads())

    def make_env(self, cfg):
        kind, n, m, nz, params = pack_env(cfg)
        params = np.ascontiguousarray(params, dtype=np.float64)
        h = self.lib.oracle_env_create(kind, n, m, nz, self._p(params))
        if not h:
            raise ValueError("oracle_env_create rejected the configuration")
        return Env(self.lib, h, n, m, cfg)

    # ------------------------------------------------------------ envs
    def env_eval(self, env, x, u):
        x = self._a(x, (-1, env.n)); u = self._a(u, (-1, env.m)); B = x.shape[0]
        nxt = np.empty((B, env.n), self.dtype); cost = np.empty(B, self.dtype); fc = np.empty(B, self.dtype)
        self.lib.oracle_env_eval(C.c_void_p(env.handle), B, self._p(x), self._p(u), self._p(nxt), self._p(cost), self._p(fc))
        return nxt, cost, fc

    def env_linearize(self, env, x, u):
        n, m = env.n, env.m
        x = self._a(x, (-1, n)); u = self._a(u, (-1, m)); B = x.shape[0]
        z = lambda *s: np.zeros((B,) + s, self.dtype)  # noqa: E731
        out = dict(f_x=z(n, n), f_u=z(n, m), l=z(), l_x=z(n), l_u=z(m), l_xx=z(n, n), l_uu=z(m, m), l_xu=z(n, m),
                   fl=z(), fl_x=z(n), fl_xx=z(n, n))
        self.lib.oracle_env_linearize(C.c_void_p(env.handle), B, self._p(x), self._p(u),
                                      *[self._p(out[k]) for k in ("f_x", "f_u", "l", "l_x", "l_u", "l_xx", "l_uu", "l_xu",
                                                                  "fl", "fl_x", "fl_xx")])
        out["l_ux"] = np.ascontiguousarray(np.swapaxes(out["l_xu"], 1, 2))
        return out

    # ------------------------------------------------------------ box-QP
    def boxqp(self, H, q, low, high, x0):
        H = self._a(H); B, m = H.shape[0], H.shape[1]
        q = self._a(q, (B, m)); low = self._a(low, (B, m)); high = self._a(high, (B, m)); x = self._a(x0, (B, m)).copy()
        Hfree = np.zeros((B, m, m), self.dtype); free = np.zeros((B, m), np.int32); nfree = np.zeros(B, np.int32)
        status = np.zeros(B, np.int32)
        self.lib.oracle_boxqp(B, m, self._p(H), self._p(q), self._p(low), self._p(high), self._p(x), self._p(Hfree),
                              self._p(free), self._p(nfree), self._p(status))
        return dict(x=x, Hfree=Hfree, free=free.astype(bool), nfree=nfree, status=status)

    # ------------------------------------------------------------ iLQR
    def ilqr_start(self, env, x0, u_init):
        n, m = env.n, env.m
        u_init = self._a(u_init); B, T = u_init.shape[0], u_init.shape[1]
        u_init = self._a(u_init, (B, T, m)); x0 = self._a(x0, (B, n))
        xs = np.empty((B, T + 1, n), self.dtype); us = np.empty((B, T, m), self.dtype); cs = np.empty((B, T + 1), self.dtype)
        self.lib.oracle_ilqr_start(C.c_void_p(env.handle), B, T, self._p(x0), self._p(u_init), self._p(xs), self._p(us), self._p(cs))
        return xs, us, cs

    def ilqr_backward(self, env, states, actions, mu=1.0):
        n, m = env.n, env.m
        actions = self._a(actions); B, T = actions.shape[0], actions.shape[1]
        actions = self._a(actions, (B, T, m)); states = self._a(states, (B, T + 1, n))
        K = np.zeros((B, T, m, n), self.dtype); k = np.zeros((B, T, m), self.dtype)
        J = np.zeros(B, self.dtype); dV1 = np.zeros(B, self.dtype); dV2 = np.zeros(B, self.dtype)
        status = np.zeros(B, np.int32); branch = np.zeros((B, 3), np.int32)
        self.lib.oracle_ilqr_backward(C.c_void_p(env.handle), B, T, self._p(states), self._p(actions), C.c_double(mu),
                                      self._p(K), self._p(k), self._p(J), self._p(dV1), self._p(dV2), self._p(status), self._p(branch))
        return dict(K=K, k=k, J=J, dV1=dV1, dV2=dV2, status=status, branch=branch)

    def ilqr_forward(self, env, states, actions, K, k, alpha=1.0):
        n, m = env.n, env.m
        actions = self._a(actions); B, T = actions.shape[0], actions.shape[1]
        actions = self._a(actions, (B, T, m)); states = self._a(states, (B, T + 1, n))
        K = self._a(K, (B, T, m, n)); k = self._a(k, (B, T, m))
        xs = np.empty((B, T + 1, n), self.dtype); us = np.empty((B, T, m), self.dtype); cs = np.empty((B, T + 1), self.dtype)
        J = np.zeros(B, self.dtype); res = np.zeros(B, self.dtype)
        self.lib.oracle_ilqr_forward(C.c_void_p(env.handle), B, T, self._p(states), self._p(actions), self._p(K), self._p(k),
                                     C.c_double(alpha), self._p(xs), self._p(us), self._p(cs), self._p(J), self._p(res))
        return dict(states=xs, actions=us, costs=cs, J=J, residual=res)

    def ilqr_solve(self, env, x0, u_init, atol=5e-3, max_iterations=100, mu_min=1e-6, delta_0=2.0, c1=0.0, alpha_min=1e-3,
                   nthreads=0):
        n, m = env.n, env.m
        u_init = self._a(u_init); B, T = u_init.shape[0], u_init.shape[1]
        u_init = self._a(u_init, (B, T, m)); x0 = self._a(x0, (B, n))
        xs = np.empty((B, T + 1, n), self.dtype); us = np.empty((B, T, m), self.dtype); cs = np.empty((B, T + 1), self.dtype)
        stats = np.zeros((B, 4), np.int32)
        self.lib.oracle_ilqr_solve(C.c_void_p(env.handle), B, T, self._p(x0), self._p(u_init), C.c_double(atol), int(max_iterations),
                                   C.c_double(mu_min), C.c_double(delta_0), C.c_double(c1), C.c_double(alpha_min),
                                   self._p(xs), self._p(us), self._p(cs), self._p(stats), int(nthreads))
        return dict(states=xs, actions=us, costs=cs, iterations=stats[:, 0].copy(), n_backward=stats[:, 1].copy(),
                    n_rollouts=stats[:, 2].copy(), status=stats[:, 3].copy())

    # ------------------------------------------------------------ LQR
    def lqr_solve(self, F, f, C_, c, x0, T, terminal_zero=False, nthreads=0):
        """F [n,N] or [B,n,N] (likewise f, C, c); x0 [B,n]."""
        x0 = self._a(x0); x0 = x0.reshape(-1, x0.shape[-1]) if x0.ndim > 1 else x0.reshape(1, -1)
        B, n = x0.shape
        F = self._a(F); N = F.shape[-1]; m = N - n
        sF = int(F.ndim == 3); F = self._a(F, (-1, n, N))
        f = self._a(f); sf = int(f.size == B * n and B > 1 and f.ndim >= 2 and f.shape[0] == B); f = self._a(f, (-1, n))
        C_ = self._a(C_); sC = int(C_.ndim == 3); C_ = self._a(C_, (-1, N, N))
        c = self._a(c); sc = int(c.ndim >= 2 and c.shape[0] == B and c.size == B * N and B > 1); c = self._a(c, (-1, N))
        xs = np.empty((B, T + 1, n), self.dtype); us = np.empty((B, T, m), self.dtype); cs = np.empty((B, T + 1), self.dtype)
        K = np.empty((B, T, m, n), self.dtype); k = np.empty((B, T, m), self.dtype)
        V = np.empty((B, T, n, n), self.dtype); v = np.empty((B, T, n), self.dtype); cst = np.empty((B, T), self.dtype)
        status = np.zeros(B, np.int32)
        self.lib.oracle_lqr_solve(B, n, m, int(T), self._p(F), sF, self._p(f), sf, self._p(C_), sC, self._p(c), sc, self._p(x0),
                                  int(bool(terminal_zero)), self._p(xs), self._p(us), self._p(cs), self._p(K), self._p(k),
                                  self._p(V), self._p(v), self._p(cst), self._p(status), int(nthreads))
        return dict(states=xs, actions=us, costs=cs, K=K, k=k, V=V, v=v, const=cst, status=status)
