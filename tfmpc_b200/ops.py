"""Functional layer over the C ABI: torch CUDA tensors in, torch CUDA tensors out.

Every function enqueues its kernels on torch's current CUDA stream and returns without
synchronising.  Shapes follow the reference (batch-first): states [B,T+1,n], actions [B,T,m],
costs [B,T+1].  There is no fallback: non-CUDA tensors are rejected by the library itself
(tfmpc_dl_unpack), and a missing extension raises in _native.load().
"""
import ctypes as C

import torch

from . import _native as N


def _prec(t):
    if t.dtype == torch.float32:
        return "f32"
    if t.dtype == torch.float64:
        return "f64"
    raise N.TfmpcError(f"unsupported dtype {t.dtype}: use float32 (product) or float64 (verification build)")


def _c(t):
    return t if t.is_contiguous() else t.contiguous()


def _empty(like, *shape):
    return torch.empty(shape, dtype=like.dtype, device=like.device)


# ------------------------------------------------------------------ LQR
def lqr_solve(F, f, Cm, c, x0, T, terminal_zero=False, want_policy=True, want_value=True, out=None):
    """LQR.solve for B problems.  F [n,N] or [B,n,N]; f [n]/[B,n]; Cm [N,N]/[B,N,N]; c [N]/[B,N];
    x0 [B,n].  Returns dict(states, actions, costs, status[, K, k][, V, v, const]); `out` reuses the
    result tensors of an earlier call with the same shapes and flags."""
    N.require_cuda()
    x0 = _c(x0)
    B, n = x0.shape
    F, f, Cm, c = _c(F), _c(f), _c(Cm), _c(c)
    NN = F.shape[-1]
    m = NN - n
    lib = N.load(_prec(x0))
    sF = n * NN if F.dim() == 3 else 0
    sf = n if f.dim() == 2 else 0
    sC = NN * NN if Cm.dim() == 3 else 0
    sc = NN if c.dim() == 2 else 0
    for t, s, nb in ((F, sF, "F"), (f, sf, "f"), (Cm, sC, "C"), (c, sc, "c")):
        if s and t.shape[0] != B:
            raise N.TfmpcError(f"batched {nb} has leading size {t.shape[0]}, expected {B}")
    T = int(T)
    if out is None:
        out = dict(states=_empty(x0, B, T + 1, n), actions=_empty(x0, B, T, m), costs=_empty(x0, B, T + 1),
                   status=torch.empty(B, dtype=torch.int32, device=x0.device))
        if want_policy or want_value:
            out["K"] = _empty(x0, B, T, m, n)
            out["k"] = _empty(x0, B, T, m)
        if want_value:
            out["V"] = _empty(x0, B, T, n, n)
            out["v"] = _empty(x0, B, T, n)
            out["const"] = _empty(x0, B, T)
    if B == 0:       # nothing to solve: empty results, no launch (an empty tensor has no device pointer to hand over)
        return out
    P = lambda t, i32=False: N.dev_ptr(lib, t, i32)  # noqa: E731
    ptrs = [P(F), P(f), P(Cm), P(c), P(x0), P(out["states"]), P(out["actions"]), P(out["costs"]), P(out.get("K")), P(out.get("k")),
            P(out.get("V")), P(out.get("v")), P(out.get("const")), P(out["status"], True)]
    N.check(lib, lib.tfmpc_lqr_solve(C.c_int64(B), n, m, T, ptrs[0].p, C.c_int64(sF), ptrs[1].p, C.c_int64(sf), ptrs[2].p, C.c_int64(sC),
                                     ptrs[3].p, C.c_int64(sc), ptrs[4].p, int(bool(terminal_zero)), ptrs[5].p, ptrs[6].p, ptrs[7].p,
                                     ptrs[8].p, ptrs[9].p, ptrs[10].p, ptrs[11].p, ptrs[12].p, ptrs[13].p, N.stream_ptr()))
    return out


def lqr_solve_host(F, f, Cm, c, x0, T, terminal_zero=False, out=None):
    """Same solve through the HOST-buffer entry point: CPU tensors in, CPU tensors out; the
    library does the H2D copy, the solve and the D2H copy, and synchronises.  `out` reuses the
    (ideally pinned) result tensors of an earlier call."""
    N.require_cuda()
    x0 = _c(x0)
    B, n = x0.shape
    F, f, Cm, c = _c(F), _c(f), _c(Cm), _c(c)
    NN = F.shape[-1]
    m = NN - n
    lib = N.load(_prec(x0))
    sF = n * NN if F.dim() == 3 else 0
    sf = n if f.dim() == 2 else 0
    sC = NN * NN if Cm.dim() == 3 else 0
    sc = NN if c.dim() == 2 else 0
    T = int(T)
    if out is None:
        out = dict(states=torch.empty(B, T + 1, n, dtype=x0.dtype), actions=torch.empty(B, T, m, dtype=x0.dtype),
                   costs=torch.empty(B, T + 1, dtype=x0.dtype), status=torch.empty(B, dtype=torch.int32))
    P = lambda t, i32=False: N.host_ptr(lib, t, i32)  # noqa: E731
    ptrs = [P(F), P(f), P(Cm), P(c), P(x0), P(out["states"]), P(out["actions"]), P(out["costs"]), P(out["status"], True)]
    N.check(lib, lib.tfmpc_lqr_solve_host(C.c_int64(B), n, m, T, ptrs[0].p, C.c_int64(sF), ptrs[1].p, C.c_int64(sf), ptrs[2].p,
                                          C.c_int64(sC), ptrs[3].p, C.c_int64(sc), ptrs[4].p, int(bool(terminal_zero)), ptrs[5].p,
                                          ptrs[6].p, ptrs[7].p, ptrs[8].p, N.stream_ptr()))
    return out


# ------------------------------------------------------------------ environments
def env_step(env, x, u, want_next=True, want_cost=True):
    """x [R,n], u [R,m] -> (x_next [R,n] | None, cost [R] | None)"""
    x, u = _c(x), _c(u)
    R = x.shape[0]
    lib = env.lib
    xn = _empty(x, R, env.n) if want_next else None
    cost = _empty(x, R) if want_cost else None
    px, pu, pn, pc = (N.dev_ptr(lib, t) for t in (x, u, xn, cost))
    N.check(lib, lib.tfmpc_env_step(env.handle, C.c_int64(R), px.p, pu.p, pn.p, pc.p, N.stream_ptr()))
    return xn, cost


def env_step_noisy(env, x, u, seed, offset, want_cost=True):
    """tfmpc_env_step_noisy: the plant of GymEnv.step (transition(cec=False)): x [R,n], u [R,m] -> (x_next, cost | None) with the
    environment's noise model drawn on the device (Philox keyed by `seed`, call number `offset`)."""
    x, u = _c(x), _c(u)
    R = x.shape[0]
    lib = env.lib
    xn = _empty(x, R, env.n)
    cost = _empty(x, R) if want_cost else None
    px, pu, pn, pc = (N.dev_ptr(lib, t) for t in (x, u, xn, cost))
    N.check(lib, lib.tfmpc_env_step_noisy(env.handle, C.c_int64(R), px.p, pu.p, pn.p, pc.p, C.c_uint64(int(seed) & (2 ** 64 - 1)),
                                          C.c_uint64(int(offset)), N.stream_ptr()))
    return xn, cost


def env_has_noise_model(env):
    return bool(env.lib.tfmpc_env_has_noise_model(env.handle))


def ilqr_initial_actions(env, B, T, seed, dtype, device):
    """tfmpc_ilqr_initial_actions: iLQR.start's random initial actions [B,T,m], drawn on the device."""
    u = torch.empty(int(B), int(T), env.m, dtype=dtype, device=device)
    lib = env.lib
    pu = N.dev_ptr(lib, u)
    N.check(lib, lib.tfmpc_ilqr_initial_actions(env.handle, C.c_int64(int(B)), int(T), C.c_uint64(int(seed) & (2 ** 64 - 1)), pu.p, N.stream_ptr()))
    return u


def env_final_cost(env, x):
    x = _c(x)
    R = x.shape[0]
    lib = env.lib
    cost = _empty(x, R)
    px, pc = N.dev_ptr(lib, x), N.dev_ptr(lib, cost)
    N.check(lib, lib.tfmpc_env_final_cost(env.handle, C.c_int64(R), px.p, pc.p, N.stream_ptr()))
    return cost


def env_linearize(env, x, u):
    """x [R,n], u [R,m] -> dict(f_x [R,n,n], f_u [R,n,m], l [R], l_x [R,n], l_u [R,m], l_xx, l_uu, l_ux, l_xu)"""
    x, u = _c(x), _c(u)
    R, n, m = x.shape[0], env.n, env.m
    lib = env.lib
    out = dict(f_x=_empty(x, R, n, n), f_u=_empty(x, R, n, m), l=_empty(x, R), l_x=_empty(x, R, n), l_u=_empty(x, R, m),
               l_xx=_empty(x, R, n, n), l_uu=_empty(x, R, m, m), l_ux=_empty(x, R, m, n), l_xu=_empty(x, R, n, m))
    ptrs = [N.dev_ptr(lib, t) for t in (x, u, out["f_x"], out["f_u"], out["l"], out["l_x"], out["l_u"], out["l_xx"], out["l_uu"],
                                        out["l_ux"], out["l_xu"])]
    N.check(lib, lib.tfmpc_env_linearize(env.handle, C.c_int64(R), *[p.p for p in ptrs], N.stream_ptr()))
    return out


def env_final_quad(env, x):
    x = _c(x)
    R, n = x.shape[0], env.n
    lib = env.lib
    out = dict(l=_empty(x, R), l_x=_empty(x, R, n), l_xx=_empty(x, R, n, n))
    ptrs = [N.dev_ptr(lib, t) for t in (x, out["l"], out["l_x"], out["l_xx"])]
    N.check(lib, lib.tfmpc_env_final_quad(env.handle, C.c_int64(R), *[p.p for p in ptrs], N.stream_ptr()))
    return out


# ------------------------------------------------------------------ box-QP
def boxqp(H, q, low, high, x0):
    """H [B,m,m], q/low/high/x0 [B,m] -> dict(x, Hfree, free, nfree, status)"""
    N.require_cuda()
    H, q, low, high = _c(H), _c(q), _c(low), _c(high)
    B, m = q.shape
    lib = N.load(_prec(H))
    x = x0.clone().contiguous()
    Hfree = _empty(H, B, m, m)
    free = torch.empty(B, m, dtype=torch.int32, device=H.device)
    nfree = torch.empty(B, dtype=torch.int32, device=H.device)
    status = torch.empty(B, dtype=torch.int32, device=H.device)
    ptrs = [N.dev_ptr(lib, t) for t in (H, q, low, high, x, Hfree)] + [N.dev_ptr(lib, t, True) for t in (free, nfree, status)]
    N.check(lib, lib.tfmpc_boxqp(C.c_int64(B), m, *[p.p for p in ptrs], N.stream_ptr()))
    return dict(x=x, Hfree=Hfree, free=free.bool(), nfree=nfree, status=status)


# ------------------------------------------------------------------ iLQR
def ilqr_start(env, x0, u_init):
    x0, u_init = _c(x0), _c(u_init)
    B, T = u_init.shape[0], u_init.shape[1]
    lib = env.lib
    states, actions, costs = _empty(x0, B, T + 1, env.n), _empty(x0, B, T, env.m), _empty(x0, B, T + 1)
    ptrs = [N.dev_ptr(lib, t) for t in (x0, u_init, states, actions, costs)]
    N.check(lib, lib.tfmpc_ilqr_start(env.handle, C.c_int64(B), T, *[p.p for p in ptrs], N.stream_ptr()))
    return states, actions, costs


def ilqr_backward(env, states, actions, mu=1.0):
    states, actions = _c(states), _c(actions)
    B, T = actions.shape[0], actions.shape[1]
    lib = env.lib
    K, k = _empty(states, B, T, env.m, env.n), _empty(states, B, T, env.m)
    J, dV1, dV2 = _empty(states, B), _empty(states, B), _empty(states, B)
    status = torch.empty(B, dtype=torch.int32, device=states.device)
    ptrs = [N.dev_ptr(lib, t) for t in (states, actions)]
    outs = [N.dev_ptr(lib, t) for t in (K, k, J, dV1, dV2)] + [N.dev_ptr(lib, status, True)]
    N.check(lib, lib.tfmpc_ilqr_backward(env.handle, C.c_int64(B), T, ptrs[0].p, ptrs[1].p, C.c_double(float(mu)), *[p.p for p in outs],
                                         N.stream_ptr()))
    return dict(K=K, k=k, J=J, dV1=dV1, dV2=dV2, status=status)


def ilqr_backward_staged(actions, tm, cm, fm, low, high, mu=1.0):
    """iLQR.backward from explicit derivative models (the reference's signature, ilqr.py:94): generic dense kernel, any
    n, m <= 32.  actions [B,T,m]; tm = (f, f_x [B,T,n,n], f_u [B,T,n,m]); cm = (l [B,T], l_x [B,T,n], l_u [B,T,m], l_xx, l_uu,
    l_ux, l_xu [B,T,n,m]); fm = (l [B], l_x [B,n], l_xx [B,n,n]); low/high: action bounds (length m, +-inf = unbounded)."""
    import numpy as np
    N.require_cuda()
    actions = _c(actions)
    B, T, m = actions.shape
    f_x, f_u = _c(tm[1]), _c(tm[2])
    n = f_x.shape[-1]
    l, l_x, l_u, l_xx, l_uu, l_xu = _c(cm[0]), _c(cm[1].reshape(B, T, n)), _c(cm[2].reshape(B, T, m)), _c(cm[3]), _c(cm[4]), _c(cm[6])
    fl, fl_x, fl_xx = _c(fm[0].reshape(B)), _c(fm[1].reshape(B, n)), _c(fm[2])
    lib = N.load(_prec(actions))
    lo = (C.c_double * m)(*[float(v) for v in np.asarray(low, dtype=np.float64).reshape(-1)])
    hi = (C.c_double * m)(*[float(v) for v in np.asarray(high, dtype=np.float64).reshape(-1)])
    K, k = _empty(actions, B, T, m, n), _empty(actions, B, T, m)
    J, dV1, dV2 = _empty(actions, B), _empty(actions, B), _empty(actions, B)
    status = torch.empty(B, dtype=torch.int32, device=actions.device)
    ins = [N.dev_ptr(lib, t) for t in (actions, f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_xu, fl, fl_x, fl_xx)]
    outs = [N.dev_ptr(lib, t) for t in (K, k, J, dV1, dV2)] + [N.dev_ptr(lib, status, True)]
    N.check(lib, lib.tfmpc_ilqr_backward_staged(C.c_int64(B), T, n, m, lo, hi, *[p.p for p in ins], C.c_double(float(mu)),
                                                *[p.p for p in outs], N.stream_ptr()))
    return dict(K=K, k=k, J=J, dV1=dV1, dV2=dV2, status=status)


def ilqr_forward(env, states, actions, K, k, alpha=1.0):
    states, actions, K, k = _c(states), _c(actions), _c(K), _c(k)
    B, T = actions.shape[0], actions.shape[1]
    lib = env.lib
    xs, us, cs = _empty(states, B, T + 1, env.n), _empty(states, B, T, env.m), _empty(states, B, T + 1)
    J, res = _empty(states, B), _empty(states, B)
    ins = [N.dev_ptr(lib, t) for t in (states, actions, K, k)]
    outs = [N.dev_ptr(lib, t) for t in (xs, us, cs, J, res)]
    N.check(lib, lib.tfmpc_ilqr_forward(env.handle, C.c_int64(B), T, *[p.p for p in ins], C.c_double(float(alpha)), *[p.p for p in outs],
                                        N.stream_ptr()))
    return dict(states=xs, actions=us, costs=cs, J=J, residual=res)


def make_opts(atol=5e-3, max_iterations=100, mu_min=1e-6, delta_0=2.0, c1=0.0, alpha_min=1e-3):
    return N.IlqrOpts(float(atol), int(max_iterations), float(mu_min), float(delta_0), float(c1), float(alpha_min))


def ilqr_solve(env, x0, u_init, opts=None, out=None):
    """iLQR.solve for B problems, device tensors.  x0 [B,n], u_init [B,T,m].
    Returns dict(states, actions, costs, stats[B,4] int32 = iteration, backward passes, rollouts, status)."""
    x0, u_init = _c(x0), _c(u_init)
    B, T = u_init.shape[0], u_init.shape[1]
    lib = env.lib
    opts = opts or make_opts()
    if out is None:
        out = dict(states=_empty(x0, B, T + 1, env.n), actions=_empty(x0, B, T, env.m), costs=_empty(x0, B, T + 1),
                   stats=torch.empty(B, 4, dtype=torch.int32, device=x0.device))
    if B == 0:
        return out
    nbytes = lib.tfmpc_ilqr_workspace_bytes(env.handle, C.c_int64(B), T)
    if nbytes < 0:
        N.check(lib, int(nbytes))
    ws = N.workspace(x0.device, max(int(nbytes), 256))
    ins = [N.dev_ptr(lib, t) for t in (x0, u_init)]
    outs = [N.dev_ptr(lib, t) for t in (out["states"], out["actions"], out["costs"])] + [N.dev_ptr(lib, out["stats"], True)]
    N.check(lib, lib.tfmpc_ilqr_solve(env.handle, C.c_int64(B), T, ins[0].p, ins[1].p, C.byref(opts), *[p.p for p in outs],
                                      C.c_void_p(ws.data_ptr()), C.c_int64(ws.numel()), N.stream_ptr()))
    return out


def set_option(name, value, precision="f32"):
    """tfmpc_set_option: runtime options of the small-environment solve ("solver": 1 queue / 0 ticks, "qp": 2 closed form /
    0 the reference's projected-Newton iteration, "queue_mode": 0 auto / 1 throughput / 2 latency, and the scheduling knobs
    "queue_warps_per_sm", "queue_w_target", "queue_w_solo", "queue_solo_max", "queue_drain_solo", "queue_patience",
    "queue_trace"; "queue_last_mode" is read-only) -- include/tfmpc_b200.h has the list.  Returns the previous value."""
    lib = N.load(precision)
    rc = lib.tfmpc_set_option(str(name).encode(), int(value))
    if rc < 0:
        N.check(lib, rc)
    return int(rc)


def queue_counters(workspace, precision="f32"):
    """Scheduling counters the queue solver left in `workspace` (the tensor handed to the last solve; synchronises):
    warp iterations, problem iterations, rollout rounds (search + store passes), store passes, watchdog flag."""
    lib = N.load(precision)
    out = (C.c_int32 * 5)()
    N.check(lib, lib.tfmpc_ilqr_queue_counters(C.c_void_p(workspace.data_ptr()), out, N.stream_ptr()))
    return dict(zip(("warp_iterations", "problem_iterations", "rounds", "store_passes", "watchdog"), [int(v) for v in out]))


def queue_trace(env, B, T, workspace, max_records=1 << 18):
    """tfmpc_ilqr_queue_trace -> int64 array [records, 8] (acquire start ns (low 32 bits), wait ns, work ns, packed lanes/rounds/warp,
    set-up ns, backward ns, search ns, store-pass ns)."""
    import numpy as np
    out = np.zeros((max_records, 8), dtype=np.uint32)
    env.lib.tfmpc_ilqr_queue_trace.restype = C.c_int64
    n = env.lib.tfmpc_ilqr_queue_trace(env.handle, C.c_int64(B), int(T), C.c_void_p(workspace.data_ptr()), out.ctypes.data_as(C.c_void_p),
                                       C.c_int64(max_records), N.stream_ptr())
    if n < 0:
        N.check(env.lib, int(n))
    return out[: int(n)].astype(np.int64)


def set_graph_mode(on, precision="f32"):
    """tfmpc_set_graph_mode: CUDA-graph replay of repeated solves on the same buffers; returns the previous mode."""
    return bool(N.load(precision).tfmpc_set_graph_mode(int(bool(on))))


def ilqr_workspace(env, B, T, device=None):
    """A private workspace tensor for one in-flight ilqr_solve_async call."""
    nbytes = env.lib.tfmpc_ilqr_workspace_bytes(env.handle, C.c_int64(B), int(T))
    if nbytes < 0:
        N.check(env.lib, int(nbytes))
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device or torch.device("cuda", torch.cuda.current_device()))


def ilqr_solve_async(env, x0, u_init, out, workspace, done, opts=None):
    """Asynchronous iLQR.solve (tfmpc_ilqr_solve_async): the current stream does not wait for the results, so
    back-to-back calls overlap one batch's straggler ticks with the next batch's head.  `out` (as returned by
    ilqr_solve) and `workspace` (ilqr_workspace) must not be shared by calls in flight; `done` is a
    torch.cuda.Event that completes when `out` is ready -- wait on it before touching `out` or reusing either."""
    x0, u_init = _c(x0), _c(u_init)
    B, T = u_init.shape[0], u_init.shape[1]
    lib = env.lib
    opts = opts or make_opts()
    if not done.cuda_event:
        done.record()          # torch creates the cudaEvent_t lazily; the library re-records it behind the results
    ins = [N.dev_ptr(lib, t) for t in (x0, u_init)]
    outs = [N.dev_ptr(lib, t) for t in (out["states"], out["actions"], out["costs"])] + [N.dev_ptr(lib, out["stats"], True)]
    N.check(lib, lib.tfmpc_ilqr_solve_async(env.handle, C.c_int64(B), T, ins[0].p, ins[1].p, C.byref(opts), *[p.p for p in outs],
                                            C.c_void_p(workspace.data_ptr()), C.c_int64(workspace.numel()), N.stream_ptr(),
                                            C.c_void_p(done.cuda_event)))
    return out


def ilqr_solve_host(env, x0, u_init, opts=None, out=None):
    """Same solve through the HOST-buffer entry point (CPU tensors in and out, copies inside)."""
    x0, u_init = _c(x0), _c(u_init)
    B, T = u_init.shape[0], u_init.shape[1]
    lib = env.lib
    opts = opts or make_opts()
    if out is None:
        out = dict(states=torch.empty(B, T + 1, env.n, dtype=x0.dtype), actions=torch.empty(B, T, env.m, dtype=x0.dtype),
                   costs=torch.empty(B, T + 1, dtype=x0.dtype), stats=torch.empty(B, 4, dtype=torch.int32))
    ins = [N.host_ptr(lib, t) for t in (x0, u_init)]
    outs = [N.host_ptr(lib, t) for t in (out["states"], out["actions"], out["costs"])] + [N.host_ptr(lib, out["stats"], True)]
    N.check(lib, lib.tfmpc_ilqr_solve_host(env.handle, C.c_int64(B), T, ins[0].p, ins[1].p, C.byref(opts), *[p.p for p in outs],
                                           N.stream_ptr()))
    return out


def ilqr_host_scratch(env, B, T, device=None):
    """Device scratch for one in-flight ilqr_solve_host_async call."""
    env.lib.tfmpc_ilqr_solve_host_scratch_bytes.restype = C.c_int64
    nbytes = env.lib.tfmpc_ilqr_solve_host_scratch_bytes(env.handle, C.c_int64(B), int(T))
    if nbytes < 0:
        N.check(env.lib, int(nbytes))
    return torch.empty(max(int(nbytes), 256), dtype=torch.uint8, device=device or torch.device("cuda", torch.cuda.current_device()))


def ilqr_solve_host_async(env, x0, u_init, out, scratch, opts=None):
    """tfmpc_ilqr_solve_host_async: HOST tensors in (x0 [B,n], u_init [B,T,m]) and out (dict as returned by ilqr_solve_host);
    copies and solve are enqueued on the current stream, nothing is synchronised.  Use pinned tensors."""
    B, T = u_init.shape[0], u_init.shape[1]
    lib = env.lib
    opts = opts or make_opts()
    ins = [N.host_ptr(lib, t) for t in (x0, u_init)]
    outs = [N.host_ptr(lib, t) for t in (out["states"], out["actions"], out["costs"])] + [N.host_ptr(lib, out["stats"], True)]
    N.check(lib, lib.tfmpc_ilqr_solve_host_async(env.handle, C.c_int64(B), T, ins[0].p, ins[1].p, C.byref(opts), *[p.p for p in outs],
                                                 C.c_void_p(scratch.data_ptr()), C.c_int64(scratch.numel()), N.stream_ptr()))
    return out
