"""GPU tests of the persistent work-queue solver (tfmpc_b200/csrc/queue_core.cuh + ilqr_queue.cu) and of the
closed-form box-QP, through the C ABI.

* the queue kernel and the per-tick launch sequence are two SCHEDULES of the same per-problem arithmetic
  (small_core.cuh): with the same box-QP flavour their results must agree problem by problem;
* the closed-form box-QP (default of the fp32 product build) replaces the reference's projected-Newton iteration
  (optimization.py:6-101) by the exact minimiser it converges to: gated against the oracle on the fp32 noise band
  (fp64 build: >= 99 % same iteration count, which the fp32-vs-fp64 band itself does not reach);
* full BASELINE-size batches (C3: 65,536 problems) are compared with the oracle problem by problem.
"""
import numpy as np
import pytest
import torch

from test_gpu_parity import _agreement, _batch_case, _cu, _dt, _env, _np

pytestmark = pytest.mark.gpu

PRECS = ["f32", "f64"]
QP_NEWTON, QP_CLOSED = 0, 2


@pytest.fixture(params=PRECS)
def prec(request):
    return request.param


@pytest.fixture
def option(request):
    """set a library option for the duration of a test"""
    from tfmpc_b200 import ops
    undo = []

    def setter(name, value, prec):
        undo.append((name, ops.set_option(name, value, prec), prec))
    yield setter
    for name, old, prec in reversed(undo):
        ops.set_option(name, old, prec)


def _solve(cfg, prec, x0, u0, with_ws=False):
    from tfmpc_b200 import _native, ops
    nat = _env(cfg, prec).native(_dt(prec))
    out = ops.ilqr_solve(nat, _cu(x0, prec), _cu(u0, prec))
    torch.cuda.synchronize()
    g = {k: _np(v) for k, v in out.items()}
    if with_ws:
        ws = _native._WS_CACHE[(torch.device("cuda", torch.cuda.current_device()), torch.cuda.current_stream().cuda_stream)]
        g["counters"] = ops.queue_counters(ws, prec)
    return g


CASES = {
    "nav_h50": (lambda s: s.navigation_config(), 3000, 50),
    "nav_h7": (lambda s: s.navigation_config(), 300, 7),
    "navlqr_box": (lambda s: s.navlqr_config([5.5, -9.0], 5.0, -1.0, 1.0), 700, 10),
    "navlqr_free": (lambda s: s.navlqr_config([5.5, -9.0], 0.5), 700, 10),
    "navlqr1_box": (lambda s: s.navlqr_config([2.5], 0.5, -0.3, 0.3), 100, 9),
    "navlqr3_box": (lambda s: s.navlqr_config([1.0, -2.0, 3.0], 0.5, -0.4, 0.6), 500, 8),
    "navlqr4_free": (lambda s: s.navlqr_config([1.0, -2.0, 3.0, 0.5], 2.0), 333, 17),
}


@pytest.mark.parametrize("case", sorted(CASES))
def test_queue_schedule_equals_tick_schedule(prec, option, case):
    """Same arithmetic, different schedule.  The tick kernels use the reference's box-QP iteration, so the queue kernel
    is set to it as well.  fp64: identical results; fp32: the two kernels are separate compilations of the same
    expressions (FMA contraction may differ), so the gate is agreement of the iteration counts on >= 99.5 %."""
    from tfmpc_b200.envs import synthetic
    mk, B, T = CASES[case]
    cfg = mk(synthetic)
    x0, u0 = _batch_case(cfg, B, T, seed=21)
    option("qp", QP_NEWTON, prec)
    option("solver", 0, prec)
    ticks = _solve(cfg, prec, x0, u0)
    option("solver", 1, prec)
    queue = _solve(cfg, prec, x0, u0, with_ws=True)
    assert (queue["stats"][:, 3] != 5).all(), "queue kernel aborted (watchdog)"
    assert queue["counters"]["watchdog"] == 0
    same = queue["stats"][:, 0] == ticks["stats"][:, 0]
    if prec == "f64":
        for k in ("stats", "states", "actions", "costs"):
            assert np.array_equal(queue[k], ticks[k]), k
    else:
        assert same.mean() >= 0.995, same.mean()
        relc = np.abs(queue["costs"].sum(1) - ticks["costs"].sum(1)) / np.maximum(np.abs(ticks["costs"].sum(1)), 1e-6)
        assert np.all(relc[same] < 1e-5), relc[same].max()
        assert (queue["stats"][same] == ticks["stats"][same]).all()


def test_queue_counters_account_for_every_iteration(prec, option):
    """Every pop of a problem runs exactly one backward pass for it (bounded environment: no retries), so the queue's
    lane count equals the sum of the per-problem backward counters; a store pass happens at most once per pop."""
    from tfmpc_b200.envs import synthetic
    cfg = synthetic.navigation_config()
    x0, u0 = _batch_case(cfg, 5000, 50, seed=5)
    option("solver", 1, prec)
    g = _solve(cfg, prec, x0, u0, with_ws=True)
    c = g["counters"]
    assert c["watchdog"] == 0 and (g["stats"][:, 3] != 5).all()
    assert c["problem_iterations"] == int(g["stats"][:, 1].sum())
    assert c["store_passes"] <= c["problem_iterations"] and c["warp_iterations"] * 32 >= c["problem_iterations"]
    assert c["rounds"] >= c["warp_iterations"]          # at least one rollout round per warp iteration (except all-converged ones)  -- loose sanity


@pytest.mark.parametrize("case", ["nav_h50", "nav_h7", "navlqr_box", "navlqr_free", "navlqr1_box"])
@pytest.mark.parametrize("qp", [QP_NEWTON, QP_CLOSED])
def test_solo_engine_equals_lane_per_problem(prec, option, case, qp):
    """The solo engine (a warp that popped <= queue_solo_max problems works on ONE problem with all its lanes: Q assembly
    spread over the lanes, candidates of all step sizes kept in shared memory, a lone problem kept until it has converged)
    evaluates the expressions of the lane-per-problem path in the same order.  queue_w_target above the batch size makes
    every visit a solo visit.  Results are identical in both builds: the schedule never changes a problem's arithmetic."""
    from tfmpc_b200.envs import synthetic
    mk, B, T = CASES[case]
    cfg = mk(synthetic)
    x0, u0 = _batch_case(cfg, B, T, seed=77)
    option("solver", 1, prec)
    option("qp", qp, prec)
    option("queue_solo_max", 0, prec)
    lanes = _solve(cfg, prec, x0, u0, with_ws=True)
    for solo_max, w_target in ((1, 1 << 20), (4, B // 3), (32, 1 << 20)):
        option("queue_solo_max", solo_max, prec)
        option("queue_w_target", w_target, prec)
        option("queue_w_solo", 1, prec)
        solo = _solve(cfg, prec, x0, u0, with_ws=True)
        assert solo["counters"]["watchdog"] == 0 and (solo["stats"][:, 3] != 5).all()
        assert solo["counters"]["problem_iterations"] == int(solo["stats"][:, 1].sum()) or case == "navlqr_free"
        if solo_max == 32 or w_target >= B:      # every visit was a solo visit: one rollout round per iteration
            assert solo["counters"]["rounds"] == solo["counters"]["warp_iterations"]
        # bit-identical in BOTH builds: fp64 is compiled without FMA contraction; in fp32 the only code that differs between the
        # two shapes -- the Q assembly -- is written as explicit fma chains (r_fma), the rest is the same functions
        for k in ("stats", "states", "actions", "costs"):
            assert np.array_equal(solo[k], lanes[k]), (k, solo_max, float((solo["stats"][:, 0] == lanes["stats"][:, 0]).mean()))


def test_auto_mode_latency_when_alone_throughput_when_pipelined(option):
    """queue_mode = 0 (auto): a solve launched while no other stream of the device has a queue solve in flight runs in latency mode
    (2), one launched behind a solve that is still running on ANOTHER stream in throughput mode (1); back-to-back solves on the
    SAME stream never overlap, so they stay in latency mode.  The choice never changes results (fp64: bit-identical)."""
    from tfmpc_b200 import ops
    from tfmpc_b200.envs import synthetic
    cfg = synthetic.navigation_config()
    x0, u0 = _batch_case(cfg, 20000, 50, seed=9)
    nat = _env(cfg, "f32").native(_dt("f32"))
    dx0, du0 = _cu(x0, "f32"), _cu(u0, "f32")
    option("queue_mode", 0, "f32")
    last = lambda: ops.set_option("queue_last_mode", 0, "f32")  # noqa: E731  (read-only option: returns what the last launch chose)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):                # warm the side stream's workspace: its allocation would synchronise the device later
        ops.ilqr_solve(nat, dx0, du0)
    torch.cuda.synchronize()
    a = ops.ilqr_solve(nat, dx0, du0)
    assert last() == 2
    b = ops.ilqr_solve(nat, dx0, du0)            # same stream, the first one still running: still alone in the sense that matters
    assert last() == 2
    with torch.cuda.stream(side):
        c = ops.ilqr_solve(nat, dx0, du0)        # another stream while the main stream's solves (~5 ms) are in flight
        assert last() == 1
    torch.cuda.synchronize()
    d = ops.ilqr_solve(nat, dx0, du0)            # everything has drained
    assert last() == 2
    torch.cuda.synchronize()
    for k in ("stats", "states", "actions", "costs"):       # latency or throughput mode, solo engine or not: same results, bit for bit
        assert torch.equal(a[k], b[k]) and torch.equal(a[k], c[k]) and torch.equal(a[k], d[k]), k
    option("queue_mode", 1, "f64")
    nat64 = _env(cfg, "f64").native(_dt("f64"))
    x64, u64 = _cu(x0[:3000], "f64"), _cu(u0[:3000], "f64")
    t = {k: v.clone() for k, v in ops.ilqr_solve(nat64, x64, u64).items()}
    option("queue_mode", 2, "f64")
    l = ops.ilqr_solve(nat64, x64, u64)
    torch.cuda.synchronize()
    for k in t:
        assert torch.equal(t[k], l[k]), k


@pytest.mark.parametrize("case", ["nav_h50", "nav_h12", "navlqr_box", "navlqr1_box"])
def test_closed_form_qp_vs_oracle(prec, option, case):
    """Closed-form box-QP (m <= 2) against the oracle, which runs the reference's projected-Newton iteration.
    fp64: same iteration count on >= 99 % of the problems, converged cost within 1e-4 on >= 99.9 %.
    fp32: the noise-band gate of test_gpu_parity.py::test_ilqr_solve_vs_oracle_small."""
    from oracle import oracle
    from tfmpc_b200.envs import synthetic
    cfg, B, T = {
        "nav_h50": (synthetic.navigation_config(), 4096, 50),
        "nav_h12": (synthetic.navigation_config(), 1024, 12),
        "navlqr_box": (synthetic.navlqr_config([5.5, -9.0], 5.0, -1.0, 1.0), 1024, 10),
        "navlqr1_box": (synthetic.navlqr_config([2.5], 0.5, -0.3, 0.3), 256, 9),
    }[case]
    x0, u0 = _batch_case(cfg, B, T, seed=31)
    option("solver", 1, prec)
    option("qp", QP_CLOSED, prec)
    g = _solve(cfg, prec, x0, u0)
    orc = oracle.Oracle(prec)
    r = orc.ilqr_solve(orc.make_env(cfg), x0, u0)
    a = _agreement(g["stats"][:, 0], g["costs"].sum(1), r["iterations"], r["costs"].sum(1))
    if prec == "f64":
        assert a["same"] >= 0.99, a["same"]
        assert a["cost_ok"] >= 0.999, a["cost_ok"]
        assert (g["stats"][:, 3] == r["status"]).mean() >= 0.999
        return
    o64 = oracle.Oracle("f64")
    r64 = o64.ilqr_solve(o64.make_env(cfg), x0, u0)
    band = _agreement(r["iterations"], r["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))
    assert a["same"] >= max(0.93, band["same"] - 0.02), (a["same"], band["same"])
    assert a["within1"] >= max(0.95, band["within1"] - 0.02), (a["within1"], band["within1"])
    assert a["cost_ok"] >= max(0.98, band["cost_ok"] - 0.01), (a["cost_ok"], band["cost_ok"])
    assert a["cost_ok_same"] >= 0.99, a["cost_ok_same"]
    # and against the fp64 truth the closed form must not be worse than the reference's own iteration in fp32
    t = _agreement(g["stats"][:, 0], g["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))
    assert t["same"] >= band["same"] - 0.02, (t["same"], band["same"])
    assert (g["stats"][:, 3] == r["status"]).mean() > 0.99


def test_full_size_c3_parity_vs_oracle(option):
    """BASELINE config C3 at its full size: 65,536 nonlinear-navigation problems, H = 50, the product configuration
    (fp32, queue kernel, closed-form box-QP) against the fp32 oracle PROBLEM BY PROBLEM, with the fp32-vs-fp64 oracle
    agreement on the same batch as the noise band."""
    from oracle import oracle
    from tfmpc_b200.envs import synthetic
    cfg = synthetic.navigation_config()
    x0, u0 = _batch_case(cfg, 65536, 50, seed=1000)
    g = _solve(cfg, "f32", x0, u0)
    o32, o64 = oracle.Oracle("f32"), oracle.Oracle("f64")
    r = o32.ilqr_solve(o32.make_env(cfg), x0, u0)
    r64 = o64.ilqr_solve(o64.make_env(cfg), x0, u0)
    a = _agreement(g["stats"][:, 0], g["costs"].sum(1), r["iterations"], r["costs"].sum(1))
    band = _agreement(r["iterations"], r["costs"].sum(1), r64["iterations"], r64["costs"].sum(1))
    print(f"C3 full size: same={a['same']:.4f} within1={a['within1']:.4f} cost<=1e-4={a['cost_ok']:.4f} | fp32-vs-fp64 oracle band "
          f"same={band['same']:.4f} within1={band['within1']:.4f} cost<=1e-4={band['cost_ok']:.4f}")
    assert (g["stats"][:, 3] != 5).all()
    assert a["same"] >= band["same"] - 0.01 and a["same"] >= 0.95, (a["same"], band["same"])
    assert a["within1"] >= band["within1"] - 0.01, (a["within1"], band["within1"])
    assert a["cost_ok"] >= band["cost_ok"] - 0.005 and a["cost_ok"] >= 0.985, (a["cost_ok"], band["cost_ok"])
    assert a["cost_ok_same"] >= 0.999
    assert (g["stats"][:, 3] == r["status"]).mean() >= 0.998
    # mean iterations per solve agree to a fraction of a percent (the metric's numerator)
    assert abs((g["stats"][:, 0] + 1.0).mean() / (r["iterations"] + 1.0).mean() - 1) < 5e-3


def test_full_size_c3_fp64_build_is_exact(option):
    """The fp64 verification build with the reference's box-QP reproduces the fp64 oracle on a 16,384-problem slice of C3:
    same iteration / backward / rollout counts on >= 99.9 % (the box-QP's own 1e-8 stopping rule is the rest)."""
    from oracle import oracle
    from tfmpc_b200.envs import synthetic
    cfg = synthetic.navigation_config()
    x0, u0 = _batch_case(cfg, 16384, 50, seed=1000)
    option("qp", QP_NEWTON, "f64")
    g = _solve(cfg, "f64", x0, u0)
    o64 = oracle.Oracle("f64")
    r = o64.ilqr_solve(o64.make_env(cfg), x0, u0)
    same = g["stats"][:, 0] == r["iterations"]
    assert same.mean() >= 0.999, same.mean()
    assert (g["stats"][same, 1] == r["n_backward"][same]).all() and (g["stats"][same, 2] == r["n_rollouts"][same]).mean() >= 0.999
    relc = np.abs(g["costs"].sum(1) - r["costs"].sum(1)) / np.abs(r["costs"].sum(1))
    assert np.all(relc[same] < 1e-6), float(relc[same].max())      # the box-QP's own 1e-8 stopping rule (optimization.py:27) moves ill-conditioned steps by ~1e-8
