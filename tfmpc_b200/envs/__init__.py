"""Environment factories -- mirror of tfmpc/envs/__init__.py:9-37."""
import importlib

import numpy as np


def make_lqr(state_size, action_size):
    """Random LQR problem: F, f, c ~ N(0,1), C = sklearn make_spd_matrix (envs/__init__.py:9-18)."""
    from sklearn.datasets import make_spd_matrix

    from ..solvers.lqr import LQR
    n_dim = state_size + action_size
    F = np.random.normal(size=(state_size, n_dim))
    f = np.random.normal(size=(state_size, 1))
    C = make_spd_matrix(n_dim)
    c = np.random.normal(size=(n_dim, 1))
    return LQR(F, f, C, c)


def make_lqr_linear_navigation(goal, beta, dtype=None):
    """LQR form of linear navigation: F = [I I], f = 0, C = diag(2,..,2b,..), c = [-2g; 0]
    (envs/__init__.py:21-30).  `goal` may be [n,1] (one problem) or [B,n] / [B,n,1] (a batch that
    shares F, f, C and varies c -- BASELINE config C2).  The reference builds F by tiling I
    `action_size` times, which is only right for n = m = 2 (quirk Q3); [I I] is used for every n."""
    from ..solvers.lqr import LQR
    g = np.asarray(goal.detach().cpu() if hasattr(goal, "detach") else goal, dtype=np.float64)
    if g.ndim == 3:
        g = g[..., 0]
    batched = g.ndim == 2 and g.shape[1] != 1
    if not batched:
        g = g.reshape(1, -1)
    n = g.shape[1]
    F = np.concatenate([np.identity(n), np.identity(n)], axis=1)
    f = np.zeros((n, 1))
    C = np.diag([2.0] * n + [2.0 * float(beta)] * n)
    c = np.concatenate([-2.0 * g, np.zeros_like(g)], axis=1)[..., None]
    kw = {} if dtype is None else {"dtype": dtype}
    return LQR(F, f, C, c if batched else c[0], **kw)


_MODULES = {"navigation": "navigation", "reservoir": "reservoir", "hvac": "hvac", "lqr.navigation": "lqr.navigation",
            # the reference's own navlin.config.json names this non-existent module (quirk Q2)
            "navigation_lqr": "lqr.navigation"}


def make_env(config):
    """config = {"module", "cls_name", "config"} as in the reference's *.config.json (envs/__init__.py:33-37)."""
    module = _MODULES.get(config["module"], config["module"])
    module = importlib.import_module(f"{__name__}.{module}")
    return getattr(module, config["cls_name"]).load(dict(config["config"]))
