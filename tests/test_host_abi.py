"""CPU-side checks of the boundary: the C-ABI library loads here (no GPU), exports every symbol
include/tfmpc_b200.h declares, fails loudly instead of falling back, and the Python mirror keeps the
reference's import surface."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "tfmpc_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(tfmpc_[a-z0-9_]+)\s*\(", text)))


@pytest.fixture(scope="module")
def built():
    import __graft_entry__
    __graft_entry__.build()
    return True


@pytest.mark.parametrize("prec", ["f32", "f64"])
def test_library_exports_every_declared_symbol(built, prec):
    from tfmpc_b200 import _native
    lib = _native.load(prec)
    syms = declared_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), f"{s} declared in include/tfmpc_b200.h but not exported"
    assert lib.tfmpc_abi_version() == 1
    assert lib.tfmpc_real_bytes() == (4 if prec == "f32" else 8)


def test_default_opts_are_the_reference_defaults(built):
    from tfmpc_b200 import _native
    lib = _native.load("f32")
    o = _native.IlqrOpts()
    lib.tfmpc_ilqr_default_opts(C.byref(o))
    # reference tfmpc/solvers/ilqr.py:27-37
    assert (o.atol, o.max_iterations, o.mu_min, o.delta_0, o.c1, o.alpha_min) == (5e-3, 100, 1e-6, 2.0, 0.0, 1e-3)


def test_dlpack_validation(built):
    """tfmpc_dl_unpack: dtype, device and contiguity checks happen in the library, on borrowed capsules."""
    from tfmpc_b200 import _native
    lib = _native.load("f32")
    t = torch.arange(12, dtype=torch.float32).reshape(3, 4)
    p = _native.host_ptr(lib, t)
    assert p.value == t.data_ptr()
    with pytest.raises(_native.TfmpcError, match="CUDA"):
        _native.dev_ptr(lib, t)                                   # CPU tensor where a device tensor is required
    with pytest.raises(_native.TfmpcError, match="float32"):
        _native.host_ptr(lib, t.double())                         # wrong dtype for this build
    with pytest.raises(_native.TfmpcError, match="contiguous"):
        _native.host_ptr(lib, t.t())                              # non-contiguous view
    with pytest.raises(_native.TfmpcError, match="int32"):
        _native.host_ptr(lib, t, int32=True)
    assert _native.host_ptr(lib, torch.zeros(3, dtype=torch.int32), int32=True).value
    assert _native.host_ptr(lib, None).p.value is None


def test_argument_validation_and_error_strings(built):
    from tfmpc_b200 import _native
    lib = _native.load("f32")
    lib.tfmpc_last_error.restype = C.c_char_p
    h = C.c_void_p()
    arr = (C.c_double * 3)(1, 2, 3)
    assert lib.tfmpc_env_create(99, 2, 2, 0, arr, C.c_int64(3), C.byref(h)) == -1
    assert b"unknown environment kind" in lib.tfmpc_last_error()
    assert lib.tfmpc_env_create(0, 2, 2, 0, arr, C.c_int64(3), C.byref(h)) == -1
    assert b"expects 7 parameters" in lib.tfmpc_last_error()
    assert lib.tfmpc_env_create(1, 3, 3, 1, arr, C.c_int64(9), C.byref(h)) == -1
    assert lib.tfmpc_env_create(2, 40, 40, 0, arr, C.c_int64(3), C.byref(h)) == -1
    assert lib.tfmpc_lqr_solve(C.c_int64(1), 2, 2, 10, None, C.c_int64(0), None, C.c_int64(0), None, C.c_int64(0), None, C.c_int64(0),
                               None, 0, None, None, None, None, None, None, None, None, None, None) == -1
    # asynchronous solve: a completion event is mandatory, and all the synchronous form's argument checks apply
    dummy = C.c_void_p(0x1000)
    assert lib.tfmpc_ilqr_solve_async(None, C.c_int64(1), 5, dummy, dummy, None, dummy, dummy, dummy, dummy, dummy, C.c_int64(1 << 20), None, None) == -1
    assert b"completion event" in lib.tfmpc_last_error()
    assert lib.tfmpc_ilqr_solve_async(None, C.c_int64(1), 5, dummy, dummy, None, dummy, dummy, dummy, dummy, dummy, C.c_int64(1 << 20), None, dummy) == -1
    assert b"bad argument" in lib.tfmpc_last_error()
    # graph mode is a plain switch that reports the previous state (off by default)
    assert lib.tfmpc_set_graph_mode(1) == 0 and lib.tfmpc_set_graph_mode(0) == 1 and lib.tfmpc_set_graph_mode(0) == 0


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback(built):
    """Without a CUDA device the product path raises; it never computes on the CPU."""
    from tfmpc_b200 import _native, envs
    from tfmpc_b200.envs import synthetic
    from tfmpc_b200.solvers.ilqr import iLQR
    env = envs.make_env(synthetic.navigation_config())
    with pytest.raises(_native.TfmpcError, match="no CPU fallback"):
        iLQR(env).solve(np.zeros((2, 1)), 10)
    with pytest.raises(_native.TfmpcError, match="no CPU fallback"):
        envs.make_lqr_linear_navigation(np.array([[8.0], [-9.0]]), 5.0).solve(np.zeros((2, 1)), 10)
    lib = _native.load("f32")
    h = C.c_void_p()
    arr = (C.c_double * 7)(1, 2, 5, -1, -1, 1, 1)
    assert lib.tfmpc_env_create(0, 2, 2, 0, arr, C.c_int64(7), C.byref(h)) == -3   # TFMPC_E_CUDA


def test_missing_extension_fails_loudly(built, monkeypatch, tmp_path):
    from tfmpc_b200 import _native
    monkeypatch.setattr(_native, "_LIBDIR", str(tmp_path))
    monkeypatch.setattr(_native, "_LIBS", {})
    with pytest.raises(_native.TfmpcError, match="has not been built"):
        _native.load("f32")


def test_product_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under tfmpc_b200/ may import, load or execute it."""
    bad = []
    for dirpath, _, files in os.walk(os.path.join(ROOT, "tfmpc_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if re.search(r"(^|\n)\s*(import|from)\s+oracle\b|liboracle|tfmpc_oracle|libemul|oracle[/.]oracle|#include\s*\"[^\"]*oracle", text):
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_reference_import_surface():
    """Same module paths and names as the reference."""
    import tfmpc
    from tfmpc import agents, envs, launchers, runners  # noqa: F401
    from tfmpc.envs import make_env, make_lqr, make_lqr_linear_navigation  # noqa: F401
    from tfmpc.envs.diffenv import CostApprox, DiffEnv, FinalCostApprox, TransitionApprox  # noqa: F401
    from tfmpc.envs.hvac import HVAC
    from tfmpc.envs.lqr.navigation import NavigationLQR
    from tfmpc.envs.navigation import Navigation
    from tfmpc.envs.reservoir import Reservoir
    from tfmpc.solvers import ilqr, lqr
    from tfmpc.utils.trajectory import Trajectory, Transition  # noqa: F401
    for name in ("n_dim", "state_size", "action_size", "transition", "cost", "final_cost", "backward", "forward", "solve", "dump", "load"):
        assert hasattr(lqr.LQR, name)
    for name in ("low", "high", "start", "derivatives", "backward", "forward", "solve"):
        assert hasattr(ilqr.iLQR, name)
    for cls in (Navigation, NavigationLQR, Reservoir, HVAC):
        for name in ("transition", "cost", "final_cost", "state_size", "action_size", "load",
                     "get_linear_transition", "get_quadratic_cost", "get_quadratic_final_cost"):
            assert hasattr(cls, name), (cls, name)
    assert tfmpc.__version__


def test_env_configs_and_packing():
    from tfmpc_b200 import envs
    from tfmpc_b200.envs import synthetic
    nav = envs.make_env(synthetic.navigation_config())
    assert nav.state_size == nav.action_size == 2 and nav.action_space.is_bounded()
    assert len(nav._pack()[1]) == 6 + 3 * 2
    res = envs.make_env(synthetic.reservoir_config(20))
    assert res.state_size == 20 and len(res._pack()[1]) == 8 * 20 + 400
    hv = envs.make_env(synthetic.hvac_grid_config(4, 8))
    assert hv.state_size == 32 and len(hv._pack()[1]) == 10 * 32 + 2 * 1024
    free = envs.make_env(synthetic.navlqr_config([1.0, 2.0, 3.0], 0.5))
    assert not free.action_space.is_bounded() and free.state_size == 3
    # the reference's own (stale) navlin.config.json module name is accepted (quirk Q2)
    cfg = synthetic.navlqr_config([5.5, -10.0], 0.0, -1.0, 1.0)
    cfg["module"] = "navigation_lqr"
    assert envs.make_env(cfg).action_space.is_bounded()
    lq = envs.make_lqr_linear_navigation(np.array([[8.0], [-9.0]]), 5.0)
    assert lq.state_size == 2 and lq.action_size == 2 and lq.n_dim == 4 and lq.batch_size is None
    assert np.allclose(lq.C.numpy(), np.diag([2, 2, 10, 10])) and np.allclose(lq.c.numpy().ravel(), [-16, 18, 0, 0])
    lqb = envs.make_lqr_linear_navigation(np.zeros((7, 2)), 1.0)
    assert lqb.batch_size == 7
    np.random.seed(0)
    r = envs.make_lqr(3, 2)
    assert r.F.shape == (3, 5) and r.C.shape == (5, 5) and np.all(np.linalg.eigvalsh(r.C.numpy().astype(np.float64)) > 0)


def test_trajectory_container(tmp_path):
    """reference tests/test_trajectory.py:20-64 + CSV format of utils/trajectory.py:71-90"""
    from tfmpc_b200.utils.trajectory import BatchTrajectory, Trajectory
    T, n, m = 5, 3, 2
    rng = np.random.RandomState(0)
    s, a, c = rng.normal(size=(T + 1, n, 1)), rng.normal(size=(T, m, 1)), rng.normal(size=(T + 1,))
    tr = Trajectory(torch.as_tensor(s), torch.as_tensor(a), torch.as_tensor(c))
    assert tr.states.shape == (T + 1, n) and tr.actions.shape == (T, m) and len(tr) == T
    assert np.allclose(tr.initial_state, s[0, :, 0]) and np.allclose(tr.final_state, s[-1, :, 0])
    assert np.isclose(tr.total_cost, c.sum()) and np.allclose(tr.cumulative_cost, np.cumsum(c))
    assert np.allclose(tr.cost_to_go, np.cumsum(c[::-1])[::-1])
    tt = tr[2]
    assert np.allclose(tt.state, s[3, :, 0]) and np.allclose(tt.action, a[2, :, 0]) and np.isclose(tt.cost, c[2])
    assert "Trajectory(init=" in repr(tr) and str(tr).count("\n") == T + 2
    path = tmp_path / "run" / "data.csv"
    tr.save(str(path))
    import pandas as pd
    df = pd.read_csv(path, index_col="Timestep")
    assert list(df.columns) == ["x[1]", "x[2]", "x[3]", "u[1]", "u[2]", "costs"] and len(df) == T
    assert np.allclose(df["x[2]"].values, s[1:, 1, 0]) and np.allclose(df["costs"].values, c[:-1])
    bt = BatchTrajectory(np.stack([s[..., 0]] * 4), np.stack([a[..., 0]] * 4), np.stack([c] * 4), iterations=np.arange(4))
    assert len(bt) == 4 and np.allclose(bt.total_cost, c.sum()) and np.allclose(bt[1].states, tr.states)


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) prints one JSON line with the contract keys,
    on a small sample so that the CPU suite stays fast."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--workload", "c4", "--steps", "1", "--warmup", "0",
                          "--cpu-sample", "32"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for key in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
                "config", "cpu_baseline", "e2e"):
        assert key in d, key
    assert d["impl"] == "reference" and d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # both arms print the SAME config object (the driver compares them): it comes from one function
    sys.path.insert(0, root)
    import bench
    assert d["config"] == bench.arm_config("c4", 16384, 1) and set(d["config"]) >= {"workload", "batch_per_gpu", "global_batch", "horizon"}


def test_bench_parity_block_counts_what_it_says():
    """bench.py's `parity` block: fractions of the batch with the same iteration count / within one / converged cost within 1e-4 /
    same status, computed against an oracle result on the same problems."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, root)
    import bench
    stats = np.array([[5, 6, 8, 0], [7, 8, 9, 0], [9, 10, 12, 1], [3, 4, 5, 2]], dtype=np.int32)
    total = np.array([10.0, 20.0, 30.0, 40.0], dtype=np.float32)
    r = {"iterations": np.array([5, 8, 9, 3]), "costs": np.array([[10.0, 0.0], [20.0, 0.001], [30.01, 0.0], [20.0, 20.0]]), "status": np.array([0, 0, 1, 0])}
    p = bench.parity_block(stats, total, r)
    assert p["problems"] == 4 and p["same_iterations"] == 0.75 and p["within1"] == 1.0
    assert p["cost_within_1e-4"] == 0.75 and abs(p["cost_within_1e-4_given_same_iterations"] - 2 / 3) < 1e-12 and p["status_match"] == 0.75
