"""Runs the __host__ __device__ per-problem code that the CUDA kernels execute (small_core.cuh: analytic
linearisation, box-QP, backward/forward passes, the on-device mu/delta/convergence schedule) on the CPU
through tests/host_emulation, and checks it against the golden fixtures and the oracle.  This is how the
GPU-less build container validates device logic; the GPU parity tests (-m gpu) are the real gate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from _util import cfg_of, golden, golden_names

HERE = os.path.dirname(os.path.abspath(__file__))
ALPHAS = np.geomspace(1.0, 1e-3, 11)


@pytest.fixture(scope="module")
def emul():
    subprocess.check_call(["bash", os.path.join(HERE, "host_emulation", "build.sh")])
    libs = {p: C.CDLL(os.path.join(HERE, "host_emulation", f"libemul_{p}.so")) for p in ("f32", "f64")}

    def solve(prec, cfg, x0, u_init):
        from oracle import oracle
        dt = np.float32 if prec == "f32" else np.float64
        kind, n, m, nz, params = oracle.pack_env(cfg)
        x0 = np.ascontiguousarray(x0, dtype=dt); u = np.ascontiguousarray(u_init, dtype=dt)
        B, T = u.shape[0], u.shape[1]
        s = np.zeros((B, T + 1, n), dt); a = np.zeros((B, T, m), dt); c = np.zeros((B, T + 1), dt); st = np.zeros((B, 4), np.int32)
        params = np.ascontiguousarray(params, dtype=np.float64)
        p = lambda z: z.ctypes.data_as(C.c_void_p)  # noqa: E731
        rc = libs[prec].emul_ilqr_solve(kind, n, nz, p(params), C.c_double(5e-3), 100, C.c_double(1e-6), C.c_double(2.0), C.c_double(0.0),
                                       p(ALPHAS), C.c_int64(B), T, p(x0), p(u), p(s), p(a), p(c), p(st))
        assert rc == 0
        return s, a, c, st
    return solve


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("name", [n for n in golden_names("solve_nav")])
def test_device_logic_matches_reference(emul, prec, name):
    d = golden(name, prec)
    cfg = cfg_of(d)
    s, a, c, st = emul(prec, cfg, d["x0"][..., 0], d["u_init"][..., 0])
    assert (st[:, 0] == d["iterations"]).all()
    assert np.all(np.abs(c.sum(1) - d["costs"].sum(1)) <= 1e-4 * np.abs(d["costs"].sum(1)))
    assert np.max(np.abs(a - d["actions"])) < (1e-4 if prec == "f32" else 1e-6)


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_device_logic_matches_oracle_on_a_batch(emul, prec):
    from oracle import oracle
    from tfmpc_b200.envs import synthetic
    o = oracle.Oracle(prec)
    cfg = synthetic.navigation_config()
    rng = np.random.RandomState(5)
    B, T = 256, 50
    x0 = synthetic.sample_x0(cfg, B, rng)
    u0 = synthetic.sample_u_init([-1, -1], [1, 1], B, T, rng) * np.ones((1, 1, 2))
    r = o.ilqr_solve(o.make_env(cfg), x0, u0)
    s, a, c, st = emul(prec, cfg, x0, u0)
    same = st[:, 0] == r["iterations"]
    assert same.mean() >= (0.99 if prec == "f32" else 1.0)
    relc = np.abs(c.sum(1) - r["costs"].sum(1)) / np.abs(r["costs"].sum(1))
    assert np.all(relc[same] < 1e-5)
    assert (st[same, 1] == r["n_backward"][same]).all() and (st[same, 2] == r["n_rollouts"][same]).all()


@pytest.mark.parametrize("prec", ["f32", "f64"])
@pytest.mark.parametrize("T", [1, 2, 3, 4, 7])
def test_device_logic_short_horizons(emul, prec, T):
    """Horizons shorter than the prefetch depth of the rollouts (3-slot ring, loads two steps ahead) and of the backward
    sweep (one step ahead): the shared device code must not read past the trajectory and must agree with the oracle."""
    from oracle import oracle
    from tfmpc_b200.envs import synthetic
    o = oracle.Oracle(prec)
    cfg = synthetic.navigation_config()
    rng = np.random.RandomState(100 + T)
    B = 64
    x0 = synthetic.sample_x0(cfg, B, rng)
    u0 = synthetic.sample_u_init([-1, -1], [1, 1], B, T, rng) * np.ones((1, 1, 2))
    r = o.ilqr_solve(o.make_env(cfg), x0, u0)
    s, a, c, st = emul(prec, cfg, x0, u0)
    same = st[:, 0] == r["iterations"]
    assert same.mean() >= (0.97 if prec == "f32" else 1.0)
    relc = np.abs(c.sum(1) - r["costs"].sum(1)) / np.maximum(np.abs(r["costs"].sum(1)), 1e-6)
    assert np.all(relc[same] < 1e-5)
    assert (st[same, 1] == r["n_backward"][same]).all() and (st[same, 2] == r["n_rollouts"][same]).all()
