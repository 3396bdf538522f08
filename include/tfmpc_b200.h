/*
 * tfmpc_b200 -- C ABI of the B200-native batched LQR / iLQR solver.
 *
 * This is the drop-in boundary for the hot path of thiagopbueno/tf-mpc v0.7.0.  The
 * reference has no FFI of its own (it is pure Python on TensorFlow), so each entry point
 * cites the reference *Python* interface it replaces (paths relative to the reference
 * repo root).  The Python mirror in tfmpc_b200/ (same class / method names as the
 * reference) reaches these symbols through ctypes (tfmpc_b200/_native.py); tensors are
 * handed over zero-copy: the shim exports each torch tensor as a DLPack capsule, asks
 * tfmpc_dl_unpack() to validate it and to return the raw pointer, and passes plain
 * pointers + sizes to the compute entry points below.
 *
 * Two builds of the same sources:  libtfmpc_b200.so      tfmpc_real = float   (product)
 *                                  libtfmpc_b200_f64.so  tfmpc_real = double  (verification build)
 *
 * Conventions
 *  - every function returns 0 on success, a negative TFMPC_E_* code on failure; the text
 *    of the last failure on the calling thread is tfmpc_last_error().  No exception or
 *    abort ever crosses this boundary.
 *  - all tensor arguments are DEVICE pointers to dense row-major arrays in the reference's
 *    layouts, unless the function name ends in _host.  Nothing is retained after return.
 *  - all work is enqueued on the caller's stream (a cudaStream_t passed as void*; NULL =
 *    legacy default stream); no hidden synchronisation, no global mutable state, so calls
 *    are re-entrant across streams and threads.  The *_host variants are the exception:
 *    they copy host->device, solve, copy device->host and synchronise their stream.
 *  - per-problem numerical trouble (non-PD Q_uu, NaN, iteration cap) is reported in the
 *    per-problem status / stats arrays, never as a call failure.
 *  - there is no CPU fallback: without a CUDA device every compute entry point fails with
 *    TFMPC_E_CUDA.
 */
#ifndef TFMPC_B200_H
#define TFMPC_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#ifdef TFMPC_F64
typedef double tfmpc_real;
#else
typedef float tfmpc_real;
#endif

#define TFMPC_ABI_VERSION 1
#define TFMPC_MAX_DIM 32   /* max state / action dimension of an environment */
#define TFMPC_MAX_ZONES 8  /* max deceleration zones of the Navigation environment */

enum {
  TFMPC_OK = 0,
  TFMPC_E_INVALID = -1,     /* bad argument (null pointer, size out of range, ...) */
  TFMPC_E_UNSUPPORTED = -2, /* valid request this build has no kernel for */
  TFMPC_E_CUDA = -3,        /* CUDA runtime error, text in tfmpc_last_error() */
  TFMPC_E_DLPACK = -4,      /* DLPack tensor failed validation */
  TFMPC_E_WORKSPACE = -5    /* workspace too small */
};

/* environment kinds (reference classes) */
enum {
  TFMPC_ENV_NAVLQR = 0,     /* tfmpc/envs/lqr/navigation/__init__.py:8  NavigationLQR */
  TFMPC_ENV_NAVIGATION = 1, /* tfmpc/envs/navigation/__init__.py:9      Navigation    */
  TFMPC_ENV_RESERVOIR = 2,  /* tfmpc/envs/reservoir/__init__.py:9       Reservoir     */
  TFMPC_ENV_HVAC = 3        /* tfmpc/envs/hvac/__init__.py:8            HVAC          */
};

/* per-problem status written by the solvers */
enum {
  TFMPC_ST_CONVERGED = 0,   /* g_norm < atol or residual < atol (ilqr.py:245-257) */
  TFMPC_ST_MAXITER = 1,     /* max_iterations exhausted (ilqr.py:227) */
  TFMPC_ST_NONPD = 2,       /* a box-QP / inverse factorisation failed (optimization.py:47-51; the reference aborts) */
  TFMPC_ST_REGLOOP = 3,     /* regularisation loop guard tripped (the reference would spin, ilqr.py:238) */
  TFMPC_ST_NAN = 4,         /* NaN in the gradient norm */
  TFMPC_ST_ABORTED = 5,     /* the solve kernel gave up before this problem finished (device watchdog); results undefined */
  TFMPC_ST_TICKS = 6        /* per-tick launch sequence only (option "solver" = 0): the fixed budget of max_iterations + 24 ticks ran
                             * out before the problem finished (more than 24 rejected line searches in total); the trajectory is the
                             * last nominal.  The work-queue, lane-per-state and dense solvers have no such budget: they allow the
                             * reference's unbounded retries up to the 200-per-iteration guard of TFMPC_ST_REGLOOP. */
};

typedef struct tfmpc_env tfmpc_env_t; /* opaque */

/* iLQR hyper-parameters: the reference's constructor kwargs and defaults, ilqr.py:27-37 */
typedef struct {
  double atol;          /* 5e-3 */
  int32_t max_iterations; /* 100 */
  double mu_min;        /* 1e-6 */
  double delta_0;       /* 2.0 */
  double c1;            /* 0.0 */
  double alpha_min;     /* 1e-3 */
} tfmpc_ilqr_opts_t;

/* ---------------------------------------------------------------- library */
int tfmpc_abi_version(void);
/* sizeof(tfmpc_real) of this build: 4 or 8 */
int tfmpc_real_bytes(void);
const char *tfmpc_last_error(void);
void tfmpc_ilqr_default_opts(tfmpc_ilqr_opts_t *opts);

/* Validate a borrowed DLManagedTensor* (DLPack ABI v0.x, as produced by
 * torch.utils.dlpack.to_dlpack) and return its data pointer.  The deleter is NEVER called.
 *   want_device: 2 = must be CUDA, 1 = must be CPU, 0 = either
 *   want_code:   0 = tfmpc_real of this build, 1 = int32
 * shape must have room for 8 entries.  Fails unless the tensor is compact row-major. */
int tfmpc_dl_unpack(const void *dl_managed_tensor, int want_device, int want_code, void **data, int32_t *ndim,
                    int64_t *shape, int32_t *device_id);

/* ---------------------------------------------------------------- LQR
 * Replaces LQR.solve = LQR.backward + LQR.forward (tfmpc/solvers/lqr.py:59-166) for B
 * independent problems.  F [n,n+m], f [n], C [n+m,n+m], c [n+m] per problem; a batch stride
 * (sF, sf, sC, sc, in elements) of 0 shares that array across the batch (config C2 shares F, f,
 * C and varies c).  Outputs: states [B,T+1,n], actions [B,T,m], costs [B,T+1]; optional (nullable)
 * K [B,T,m,n], k [B,T,m], V [B,T,n,n], v [B,T,n], cst [B,T] = the policy / value_fn lists of
 * lqr.py:126-129.  status [B] (nullable): 0 or TFMPC_ST_NONPD when Q_uu was singular.
 * terminal_zero != 0 selects V_T = v_T = 0 without a final cost (the README.md:74-92 table);
 * 0 is the v0.7.0 source (lqr.py:67-68,154-155). */
int tfmpc_lqr_solve(int64_t B, int n, int m, int T, const tfmpc_real *F, int64_t sF, const tfmpc_real *f, int64_t sf,
                    const tfmpc_real *C, int64_t sC, const tfmpc_real *c, int64_t sc, const tfmpc_real *x0,
                    int terminal_zero, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs, tfmpc_real *K,
                    tfmpc_real *k, tfmpc_real *V, tfmpc_real *v, tfmpc_real *cst, int32_t *status, void *stream);

/* LQR.forward(policy, x0, T) (lqr.py:131-161) for a caller-supplied policy K [B,T,m,n], k [B,T,m]. */
int tfmpc_lqr_forward(int64_t B, int n, int m, int T, const tfmpc_real *F, int64_t sF, const tfmpc_real *f, int64_t sf,
                      const tfmpc_real *C, int64_t sC, const tfmpc_real *c, int64_t sc, const tfmpc_real *K,
                      const tfmpc_real *k, const tfmpc_real *x0, tfmpc_real *states, tfmpc_real *actions,
                      tfmpc_real *costs, void *stream);
/* LQR.transition / LQR.cost / LQR.final_cost (lqr.py:36-57) for R rows: x [R,n], u [R,m] (may be NULL
 * when only final_cost is wanted) -> x_next [R,n], cost [R], final_cost [R]; any output may be NULL. */
int tfmpc_lqr_step(int64_t R, int n, int m, const tfmpc_real *F, int64_t sF, const tfmpc_real *f, int64_t sf,
                   const tfmpc_real *C, int64_t sC, const tfmpc_real *c, int64_t sc, const tfmpc_real *x,
                   const tfmpc_real *u, tfmpc_real *x_next, tfmpc_real *cost, tfmpc_real *final_cost, void *stream);

/* ---------------------------------------------------------------- environments
 * Replaces Env.load(config) of the four env classes (e.g. navigation/__init__.py:86-96).
 * `params` is a HOST array of doubles, copied; layout per kind:
 *   NAVLQR      goal[n] beta low[n] high[n]                       (+-inf = unbounded)
 *   NAVIGATION  goal[2] low[2] high[2] center[nz][2] decay[nz]
 *   RESERVOIR   max_res_cap lower_bound upper_bound low_penalty high_penalty set_point_penalty
 *               rain_shape rain_scale (each [n]) downstream[n][n]
 *   HVAC        temp_outside temp_hall temp_lower_bound temp_upper_bound R_outside R_hall capacity
 *               air_max adj_outside adj_hall (each [n]) R_wall[n][n] adj[n][n]
 */
int tfmpc_env_create(int kind, int n, int m, int nz, const double *params, int64_t nparams, tfmpc_env_t **env);
int tfmpc_env_destroy(tfmpc_env_t *env);
/* state_size, action_size, action_space.is_bounded(), and low/high (host copies, length m) */
int tfmpc_env_info(const tfmpc_env_t *env, int32_t *n, int32_t *m, int32_t *bounded, double *low, double *high);

/* transition(state, action, batch=True) and cost(state, action, batch=True) for R rows:
 * x [R,n], u [R,m] -> x_next [R,n] (nullable), cost [R] (nullable).
 * (navigation/__init__.py:34-54 etc.; deterministic cec=True dynamics) */
int tfmpc_env_step(const tfmpc_env_t *env, int64_t R, const tfmpc_real *x, const tfmpc_real *u, tfmpc_real *x_next,
                   tfmpc_real *cost, void *stream);
/* GymEnv.step's plant (tfmpc/envs/gymenv.py:15-25 -> transition(state, action, cec=False)): tfmpc_env_step, then the
 * environment's noise model applied to x_next in place, drawn on the device from a counter-based generator (Philox4x32-10)
 * keyed by `seed`; `offset` numbers the call (the same (seed, offset) reproduces the same draws, advance it every step):
 *   Navigation  x' += truncated normal(0, sigma = 0.2, re-drawn beyond 2 sigma)      navigation/__init__.py:45
 *   Reservoir   rainfall ~ Gamma(rain_shape, scale = rain_scale) instead of its mean  reservoir/__init__.py:98-105
 * NavigationLQR and HVAC have no noise model in the reference (their transition() takes no `cec`): deterministic.
 * tfmpc_env_has_noise_model() tells the two groups apart. */
int tfmpc_env_step_noisy(const tfmpc_env_t *env, int64_t R, const tfmpc_real *x, const tfmpc_real *u, tfmpc_real *x_next,
                         tfmpc_real *cost, uint64_t seed, uint64_t offset, void *stream);
int tfmpc_env_has_noise_model(const tfmpc_env_t *env);
/* final_cost(state): x [R,n] -> cost [R] */
int tfmpc_env_final_cost(const tfmpc_env_t *env, int64_t R, const tfmpc_real *x, tfmpc_real *cost, void *stream);
/* DiffEnv.get_linear_transition + get_quadratic_cost (tfmpc/envs/diffenv.py:13-83) with analytic
 * derivatives: x [R,n], u [R,m] -> f_x [R,n,n] f_u [R,n,m] l [R] l_x [R,n] l_u [R,m] l_xx [R,n,n]
 * l_uu [R,m,m] l_ux [R,m,n] l_xu [R,n,m]; any output may be NULL. */
int tfmpc_env_linearize(const tfmpc_env_t *env, int64_t R, const tfmpc_real *x, const tfmpc_real *u, tfmpc_real *f_x,
                        tfmpc_real *f_u, tfmpc_real *l, tfmpc_real *l_x, tfmpc_real *l_u, tfmpc_real *l_xx,
                        tfmpc_real *l_uu, tfmpc_real *l_ux, tfmpc_real *l_xu, void *stream);
/* DiffEnv.get_quadratic_final_cost (diffenv.py:85-101): x [R,n] -> l [R], l_x [R,n], l_xx [R,n,n] */
int tfmpc_env_final_quad(const tfmpc_env_t *env, int64_t R, const tfmpc_real *x, tfmpc_real *l, tfmpc_real *l_x,
                         tfmpc_real *l_xx, void *stream);

/* ---------------------------------------------------------------- box-QP
 * utils/optimization.py:6-101 projected_newton_qp for B independent problems of size m:
 * H [B,m,m], q/low/high [B,m], x [B,m] in (start) / out (solution); Hfree [B,m,m] receives the
 * Cholesky factor of H[free,free] in its leading nfree x nfree block; free [B,m] int32 flags;
 * nfree [B]; status [B]. */
int tfmpc_boxqp(int64_t B, int m, const tfmpc_real *H, const tfmpc_real *q, const tfmpc_real *low,
                const tfmpc_real *high, tfmpc_real *x, tfmpc_real *Hfree, int32_t *isfree, int32_t *nfree,
                int32_t *status, void *stream);

/* ---------------------------------------------------------------- iLQR stages (tests/test_ilqr.py:48-109)
 * iLQR.start (ilqr.py:53-82) with the initial actions supplied instead of drawn:
 * x0 [B,n], u_init [B,T,m] -> states [B,T+1,n], actions [B,T,m] (copy of u_init), costs [B,T+1]. */
int tfmpc_ilqr_start(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *x0, const tfmpc_real *u_init,
                     tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs, void *stream);
/* The random initial actions iLQR.start draws (ilqr.py:59-70): u_init [B,T,m] with ONE uniform scalar per (problem, step)
 * broadcast over the action dimensions and scaled to [low, high] (infinite bounds -> -1 / +1); device-side Philox keyed by
 * `seed`, reproducible for a given (seed, B, T). */
int tfmpc_ilqr_initial_actions(const tfmpc_env_t *env, int64_t B, int T, uint64_t seed, tfmpc_real *u_init, void *stream);
/* iLQR.derivatives + iLQR.backward fused (ilqr.py:84-172, controllers :357-387): linearises along
 * (states, actions) and sweeps t = T-1..0.  -> K [B,T,m,n], k [B,T,m], J/dV1/dV2 [B], status [B]
 * (0 ok, 1 = unconstrained Cholesky failed -- the caller's retry rule is ilqr.py:305-309,
 * 2 = box-QP factorisation failed). */
int tfmpc_ilqr_backward(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *states, const tfmpc_real *actions,
                        double mu, tfmpc_real *K, tfmpc_real *k, tfmpc_real *J, tfmpc_real *dV1, tfmpc_real *dV2,
                        int32_t *status, void *stream);
/* iLQR.backward exactly as the reference defines it (ilqr.py:94-172): consumes caller-supplied derivative models
 * (TransitionApprox / CostApprox / FinalCostApprox, diffenv.py:6-8) for B problems and any n, m <= 32:
 * f_x [B,T,n,n] f_u [B,T,n,m] l [B,T] l_x [B,T,n] l_u [B,T,m] l_xx [B,T,n,n] l_uu [B,T,m,m] l_xu [B,T,n,m],
 * fl [B] fl_x [B,n] fl_xx [B,n,n], actions [B,T,m]; low/high [m] are HOST arrays (+-inf = unbounded,
 * env.action_space.low/high).  All three controllers (:357-387) with the dispatch rule of :136-143.
 * -> K [B,T,m,n], k [B,T,m], J/dV1/dV2 [B], status [B] (0, 1 = Cholesky of Q_uu_reg failed, 2 = box-QP failed).
 * One warp per problem; each timestep's block is staged into shared memory with TMA bulk copies. */
int tfmpc_ilqr_backward_staged(int64_t B, int T, int n, int m, const double *low, const double *high, const tfmpc_real *actions,
                               const tfmpc_real *f_x, const tfmpc_real *f_u, const tfmpc_real *l, const tfmpc_real *l_x,
                               const tfmpc_real *l_u, const tfmpc_real *l_xx, const tfmpc_real *l_uu, const tfmpc_real *l_xu,
                               const tfmpc_real *fl, const tfmpc_real *fl_x, const tfmpc_real *fl_xx, double mu, tfmpc_real *K,
                               tfmpc_real *k, tfmpc_real *J, tfmpc_real *dV1, tfmpc_real *dV2, int32_t *status, void *stream);
/* iLQR.forward (ilqr.py:174-212) -> xs [B,T+1,n], us [B,T,m], cs [B,T+1], J [B], residual [B] */
int tfmpc_ilqr_forward(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *states, const tfmpc_real *actions,
                       const tfmpc_real *K, const tfmpc_real *k, double alpha, tfmpc_real *xs, tfmpc_real *us,
                       tfmpc_real *cs, tfmpc_real *J, tfmpc_real *residual, void *stream);

/* ---------------------------------------------------------------- iLQR solve
 * iLQR.solve (ilqr.py:214-283) for B problems with the whole schedule (mu, delta, convergence,
 * line search, ilqr.py:236-277,285-355) on the device.  x0 [B,n], u_init [B,T,m] (the reference
 * draws these at random, ilqr.py:70; pass them explicitly) -> states [B,T+1,n], actions [B,T,m],
 * costs [B,T+1], stats [B,4] int32 = {iteration index returned by the reference, backward passes,
 * reference-semantics rollouts, TFMPC_ST_* status}.
 * workspace: device scratch of at least tfmpc_ilqr_workspace_bytes() bytes, 256-byte aligned. */
int64_t tfmpc_ilqr_workspace_bytes(const tfmpc_env_t *env, int64_t B, int T);
int tfmpc_ilqr_solve(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *x0, const tfmpc_real *u_init,
                     const tfmpc_ilqr_opts_t *opts, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs,
                     int32_t *stats, void *workspace, int64_t workspace_bytes, void *stream);
/* Runtime options of the small-environment solve (NavigationLQR n <= 4, Navigation).  Returns the previous value, or
 * TFMPC_E_INVALID for an unknown name.  Options and their environment-variable defaults:
 *   "solver"             1 = persistent work-queue kernel (default), 0 = per-tick launch sequence (round-1 path)   TFMPC_SOLVER=queue|ticks
 *   "qp"                 box-QP of the constrained controller (ilqr.py:364-387) for m <= 2:
 *                        2 = closed form (default of the fp32 build), 0 = the reference's projected-Newton iteration
 *                        (optimization.py:6-101; default of the fp64 verification build)                        TFMPC_QP=closed|newton
 *   "queue_warps_per_sm" resident warps per SM of the queue kernel (<= 18; 0 = the mode decides: 18 / 13)           TFMPC_QUEUE_WPS
 *   "queue_mode"         scheduling policy of the queue kernel: 1 = throughput (several batches in flight: the last
 *                        problems of a batch stay in full warps), 2 = latency (a batch that has the GPU to
 *                        itself: the last problems spread over every warp slot and, once there are fewer problems than
 *                        slots, each gets a whole warp -- the solo engine), 0 = auto (default): latency when no other
 *                        stream of the device has a queue solve in flight at launch time                         TFMPC_QUEUE_MODE
 *   "queue_w_target"     warps the queue plans its pop size for: a warp pops clamp(ceil(unfinished / w_target), 1, 32)
 *                        problems (0 = the mode decides: one warp per four SMs / 15 warps per SM)                      TFMPC_QUEUE_WTARGET
 *   "queue_solo_max"     a warp that popped <= this many problems runs them one after another on the solo engine (all
 *                        lanes on one problem, everything in shared memory; a lone problem stays until it has converged);
 *                        0 = never, 255 = the mode decides (0 / 1)                                               TFMPC_QUEUE_SOLO
 *   "queue_w_solo"       once <= this many problems are unfinished every warp pops one (0 = the mode decides: the pop-size
 *                        target / the number of warp slots)                                                      TFMPC_QUEUE_WSOLO
 *   "queue_drain_solo"   throughput mode only: 1 (default) = a warp that popped a lone problem takes it on the solo engine once NO
 *                        solve of the device is in its bulk phase any more (the pipeline is draining: nothing else wants the issue
 *                        slots; +2.5 % on a 20-batch run), 0 = never                                               TFMPC_QUEUE_DRAIN_SOLO
 *   "queue_patience"     idle polls before a warp takes fewer problems than planned (default 0)                 TFMPC_QUEUE_PATIENCE
 *   "queue_trace"        1 = record one scheduling-trace record per warp iteration (diagnostics)                 TFMPC_QUEUE_TRACE
 *   "queue_last_mode"    read-only: returns the mode (1 / 2) the last launch chose; the value passed is ignored */
int tfmpc_set_option(const char *name, int value);
/* Diagnostics: copies the control block the queue solver left in `workspace` (after the solve has completed on `stream`):
 * out[0] = warp iterations, out[1] = problem iterations (lanes), out[2] = rollout rounds incl. store passes,
 * out[3] = store passes (problems), out[4] = watchdog flag.  Synchronises the stream. */
int tfmpc_ilqr_queue_counters(const void *workspace, int32_t *out, void *stream);
/* Diagnostics: the scheduling trace of the last solve of (env, B, T) in `workspace` (option "queue_trace" on): records of
 * 8 uint32 = {acquire start (low 32 bits of the ns timer), ns waiting for tickets, ns working, lanes | rounds << 8 |
 * warp slot << 16, ns state set-up, ns backward, ns search rounds, ns store pass}.  Returns the number of records copied to `out` (HOST memory), or a negative error.  Synchronises. */
int64_t tfmpc_ilqr_queue_trace(const tfmpc_env_t *env, int64_t B, int T, const void *workspace, uint32_t *out, int64_t max_records,
                               void *stream);

/* CUDA-graph replay of the tick solve's launch sequence (small environments): with on != 0 the second and later calls
 * with the same (environment, B, T, options, buffer addresses) replay two captured graphs instead of enqueueing 253
 * kernels (host cost per solve 1.4 ms -> 0.3 ms, results bit-identical).  Off by default;
 * the environment variable TFMPC_GRAPH=1 sets the initial mode.  Returns the previous mode. */
int tfmpc_set_graph_mode(int on);
/* Asynchronous form for callers that keep several batches in flight.  Inputs are read in `stream` order, but `stream`
 * does NOT wait for the results: the straggler part of the solve (the ticks after most problems have converged) runs
 * on an internal high-priority stream, so the next call on the same `stream` starts at once and its throughput-bound
 * head overlaps this call's latency-bound tail.  `done_event` (a cudaEvent_t of the same device, created by the
 * caller) is recorded behind the results: wait on it (cudaStreamWaitEvent / cudaEventSynchronize) before reading
 * states/actions/costs/stats or reusing them or the workspace.  Concurrent calls need distinct outputs and workspaces. */
int tfmpc_ilqr_solve_async(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *x0, const tfmpc_real *u_init,
                           const tfmpc_ilqr_opts_t *opts, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs,
                           int32_t *stats, void *workspace, int64_t workspace_bytes, void *stream, void *done_event);
/* Same call with HOST buffers (pageable or pinned): copies inputs to the device, solves, copies
 * the results back and synchronises.  Device scratch is cached inside the env handle. */
int tfmpc_ilqr_solve_host(tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *x0, const tfmpc_real *u_init,
                          const tfmpc_ilqr_opts_t *opts, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs,
                          int32_t *stats, void *stream);
/* The same, without the synchronisation, for callers that keep several batches in flight from ONE host thread: the
 * host->device copies, the solve and the device->host copies are enqueued on `stream` and the call returns (with PINNED
 * host buffers the copies are truly asynchronous; pageable buffers make them synchronous, as with cudaMemcpyAsync).
 * `scratch` is caller-owned device memory of at least tfmpc_ilqr_solve_host_scratch_bytes() bytes (256-byte aligned) that
 * must not be shared by calls in flight on different streams.  Results are valid once `stream` has reached this point. */
int64_t tfmpc_ilqr_solve_host_scratch_bytes(const tfmpc_env_t *env, int64_t B, int T);
int tfmpc_ilqr_solve_host_async(const tfmpc_env_t *env, int64_t B, int T, const tfmpc_real *x0, const tfmpc_real *u_init,
                                const tfmpc_ilqr_opts_t *opts, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs,
                                int32_t *stats, void *scratch, int64_t scratch_bytes, void *stream);
int tfmpc_lqr_solve_host(int64_t B, int n, int m, int T, const tfmpc_real *F, int64_t sF, const tfmpc_real *f, int64_t sf,
                         const tfmpc_real *C, int64_t sC, const tfmpc_real *c, int64_t sc, const tfmpc_real *x0,
                         int terminal_zero, tfmpc_real *states, tfmpc_real *actions, tfmpc_real *costs, int32_t *status,
                         void *stream);

/* Measurement utility for bench.py: best-of-5 sustained FP32 FMA throughput of the current device in
 * TFLOP/s (2 flops per FMA) -- the denominator of the CUDA-core roofline.  Synchronous. */
int tfmpc_measure_fp32_peak(double *tflops, double *kernel_ms);

/* number of kernels this library has launched on behalf of the calling process (monotonic) */
int64_t tfmpc_kernel_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* TFMPC_B200_H */
