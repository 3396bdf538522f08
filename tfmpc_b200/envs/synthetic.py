"""Synthetic environment configs and initial-state samplers for the BASELINE.json
workloads that the reference does not ship:

  C3  nav.config.json (shipped)      x0 ~ U(-2,2)^2
  C4  20-reservoir chain              pattern of tfmpc/envs/reservoir/res4.config.json
  C5  32-room 4x8 HVAC grid           generalises tfmpc/envs/hvac/hvac6.config.json (a 2x3 grid)

Everything here is plain Python/NumPy producing dicts in the reference's own env
JSON format ({"module", "cls_name", "config", "initial_state"}), so the same dict
feeds the reference (golden generation), the oracle and this package's make_env.
No CUDA is touched on import.
"""
import numpy as np

NAV_CONFIG = {
    # verbatim values of the reference's tfmpc/envs/navigation/nav.config.json
    "module": "navigation",
    "cls_name": "Navigation",
    "config": {
        "goal": [[8.0], [9.0]],
        "deceleration": {"center": [[[5.0], [4.5]], [[1.5], [3.0]]], "decay": [1.15, 1.2]},
        "low": [[-1.0], [-1.0]],
        "high": [[1.0], [1.0]],
    },
    "initial_state": [[0.0], [0.0]],
}


def navigation_config():
    import copy
    return copy.deepcopy(NAV_CONFIG)


def navlqr_config(goal, beta, low=None, high=None):
    cfg = {"goal": [[float(g)] for g in np.ravel(goal)], "beta": float(beta)}
    if low is not None:
        cfg["low"] = float(low)
    if high is not None:
        cfg["high"] = float(high)
    return {"module": "lqr.navigation", "cls_name": "NavigationLQR", "config": cfg,
            "initial_state": [[0.0]] * len(cfg["goal"])}


def reservoir_config(n):
    """n-reservoir chain: constants of res4.config.json, bounds lb_i=20+10i, ub_i=min(80+50i, 900)."""
    col = lambda v: [[float(v)] for _ in range(n)]  # noqa: E731
    lower = [[20.0 + 10.0 * i] for i in range(n)]
    upper = [[min(80.0 + 50.0 * i, 900.0)] for i in range(n)]
    downstream = [[1 if j == i + 1 else 0 for j in range(n)] for i in range(n)]
    return {
        "module": "reservoir",
        "cls_name": "Reservoir",
        "config": {
            "max_res_cap": col(1000.0), "low_penalty": col(-5.0), "high_penalty": col(-100.0),
            "set_point_penalty": col(-0.1), "rain_shape": col(16.0), "rain_scale": col(1.25),
            "lower_bound": lower, "upper_bound": upper, "downstream": downstream,
        },
        "initial_state": [[(lo[0] + up[0]) / 2.0] for lo, up in zip(lower, upper)],
    }


def hvac_grid_config(rows, cols):
    """rows x cols room grid; at 2x3 this reproduces hvac6.config.json exactly:
    adj = right/down neighbours, adj_outside = corner rooms, adj_hall = all rooms."""
    n = rows * cols
    col = lambda v: [[v] for _ in range(n)]  # noqa: E731
    adj = [[False] * n for _ in range(n)]
    for r in range(rows):
        for c in range(cols):
            i = r * cols + c
            if c + 1 < cols:
                adj[i][i + 1] = True
            if r + 1 < rows:
                adj[i][i + cols] = True
    corners = {0, cols - 1, (rows - 1) * cols, n - 1}
    return {
        "module": "hvac",
        "cls_name": "HVAC",
        "config": {
            "temp_outside": col(6.0), "temp_hall": col(10.0),
            "temp_lower_bound": col(20.0), "temp_upper_bound": col(23.5),
            "R_outside": col(4.0), "R_hall": col(2.0),
            "R_wall": [[1.5] * n for _ in range(n)],
            "capacity": col(80.0), "air_max": col(10.0),
            "adj": adj,
            "adj_outside": [[i in corners] for i in range(n)],
            "adj_hall": col(True),
        },
        "initial_state": col(10.0),
    }


def sample_x0(cfg, batch, rng):
    """Initial states of SURVEY section 8(d): returns float64 [batch, n]; caller casts."""
    kind = cfg["cls_name"]
    c = cfg["config"]
    if kind == "Navigation":
        centers = np.array(c["deceleration"]["center"], dtype=np.float64).reshape(-1, 2)
        x0 = rng.uniform(-2.0, 2.0, size=(batch, 2))
        for _ in range(100):  # reject points on a zone centre (lambda is not differentiable there)
            bad = (np.linalg.norm(x0[:, None, :] - centers[None], axis=-1) < 1e-3).any(axis=1)
            if not bad.any():
                break
            x0[bad] = rng.uniform(-2.0, 2.0, size=(int(bad.sum()), 2))
        return x0
    if kind == "Reservoir":
        lb = np.array(c["lower_bound"], dtype=np.float64).reshape(-1)
        ub = np.array(c["upper_bound"], dtype=np.float64).reshape(-1)
        return lb + rng.uniform(size=(batch, lb.size)) * (ub - lb)
    if kind == "HVAC":
        n = len(c["temp_lower_bound"])
        return rng.normal(loc=10.0, scale=1.0, size=(batch, n))
    if kind == "NavigationLQR":
        n = len(c["goal"])
        return rng.normal(size=(batch, n))
    raise ValueError(f"unknown env class {kind}")


def sample_u_init(low, high, batch, horizon, rng):
    """One U(0,1) scalar per (problem, step) scaled to [low, high] in every action
    dimension -- what iLQR.start does (reference tfmpc/solvers/ilqr.py:59-70)."""
    low = np.asarray(low, dtype=np.float64).reshape(-1)
    high = np.asarray(high, dtype=np.float64).reshape(-1)
    lo = np.where(np.isinf(low), -1.0, low)
    hi = np.where(np.isinf(high), 1.0, high)
    r = rng.uniform(size=(batch, horizon, 1))
    return lo + r * (hi - lo)
