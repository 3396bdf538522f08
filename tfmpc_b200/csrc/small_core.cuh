// Per-problem iLQR building blocks for SMALL environments (n, m <= 4): everything lives in
// registers of ONE thread.  Used by the thread-per-problem kernels in ilqr_small.cu.
// All functions are __host__ __device__ so that tests/host_emulation can execute the very
// same code on the CPU (a debugging aid for this GPU-less build container; the shipped
// library contains no host path).
//
// Reference lines (relative to the reference repo root) are cited per block.
#pragma once
#include "common.cuh"

template <int N, int M>
struct Lin {  // TransitionApprox / CostApprox of tfmpc/envs/diffenv.py:6-8, one timestep
  real f_x[N * N], f_u[N * M], l, l_x[N], l_u[M], l_xx[N * N], l_uu[M * M], l_xu[N * M];
};

// ------------------------------------------------------------------ environments
// Internal kind: Navigation with at most 2 deceleration zones (every reference configuration, nav.config.json has 2).
// Same arithmetic as TFMPC_ENV_NAVIGATION; only the compile-time bound of the (unrolled, predicated) zone loops differs,
// which makes a timestep's code 4x shorter -- the instruction cache is shared by warps sitting in different phases.
#define TFMPC_ENV_NAVIGATION_Z2 101
template <int KIND>
struct ZoneBound { static constexpr int v = KIND == TFMPC_ENV_NAVIGATION_Z2 ? 2 : MAXZ; };

// Navigation: tfmpc/envs/navigation/__init__.py:34-74
template <int ZM = MAXZ>
HD real nav_lambda(const EnvSmall &e, const real *x, real *lam_z, real *r_z) {
  real lam = (real)1;
#pragma unroll
  for (int z = 0; z < ZM; z++) {
    if (z < e.nz) {
      real d0 = x[0] - e.center[z][0], d1 = x[1] - e.center[z][1];
      real r = r_sqrt(d0 * d0 + d1 * d1);
      real l = (real)2 / ((real)1 + r_exp(-e.decay[z] * r)) - (real)1;
      if (lam_z) { lam_z[z] = l; r_z[z] = r; }
      lam *= l;
    }
  }
  return lam;
}

template <int KIND, int N, int M>
HD void env_step(const EnvSmall &e, const real *x, const real *u, real *xn) {
  if (KIND == TFMPC_ENV_NAVLQR) {  // lqr/navigation/__init__.py:30-32
#pragma unroll
    for (int i = 0; i < N; i++) xn[i] = x[i] + u[i];
  } else {  // navigation/__init__.py:34-48, cec=True
    real lam = nav_lambda<ZoneBound<KIND>::v>(e, x, nullptr, nullptr);
#pragma unroll
    for (int i = 0; i < N; i++) xn[i] = x[i] + lam * u[i];
  }
}

template <int KIND, int N, int M>
HD real env_cost(const EnvSmall &e, const real *x, const real *u) {
  real c1 = 0;
#pragma unroll
  for (int i = 0; i < N; i++) c1 += (x[i] - e.goal[i]) * (x[i] - e.goal[i]);
  if (KIND == TFMPC_ENV_NAVLQR) {  // lqr/navigation/__init__.py:34-41
    real c2 = 0;
#pragma unroll
    for (int i = 0; i < M; i++) c2 += u[i] * u[i];
    return c1 + e.beta * c2;
  }
  return c1;  // navigation/__init__.py:50-54 (no action cost)
}

template <int KIND, int N, int M>
HD real env_final_cost(const EnvSmall &e, const real *x) {  // :43-47 / :56-60
  real c1 = 0;
#pragma unroll
  for (int i = 0; i < N; i++) c1 += (x[i] - e.goal[i]) * (x[i] - e.goal[i]);
  return c1;
}

// Analytic DiffEnv.get_linear_transition / get_quadratic_cost (diffenv.py:13-83); closed forms
// pinned by the reference's tests/test_env_navigation.py:62-176 and test_env_lqr_navigation.py:28-135.
template <int KIND, int N, int M>
HD void env_linearize(const EnvSmall &e, const real *x, const real *u, Lin<N, M> &L) {
#pragma unroll
  for (int i = 0; i < N * N; i++) { L.f_x[i] = 0; L.l_xx[i] = 0; }
#pragma unroll
  for (int i = 0; i < N * M; i++) { L.f_u[i] = 0; L.l_xu[i] = 0; }
#pragma unroll
  for (int i = 0; i < M * M; i++) L.l_uu[i] = 0;
  L.l = env_cost<KIND, N, M>(e, x, u);
  if (KIND == TFMPC_ENV_NAVLQR) {
#pragma unroll
    for (int i = 0; i < N; i++) {
      L.f_x[i * N + i] = 1; L.f_u[i * M + i] = 1;
      L.l_x[i] = (real)2 * (x[i] - e.goal[i]); L.l_xx[i * N + i] = 2;
    }
#pragma unroll
    for (int i = 0; i < M; i++) { L.l_u[i] = r_mul((real)2 * e.beta, u[i]); L.l_uu[i * M + i] = (real)2 * e.beta; }   // (r_mul: must not fuse into Q_u = l_u + ...)
  } else {
    constexpr int ZM = ZoneBound<KIND>::v;
    real lam_z[ZM], r_z[ZM], g0 = 0, g1 = 0;
    real lam = nav_lambda<ZM>(e, x, lam_z, r_z);
#pragma unroll
    for (int z = 0; z < ZM; z++) {
      if (z < e.nz) {
        real ex = r_exp(-e.decay[z] * r_z[z]);  // d lambda_z / d r = 2 d e^{-dr} / (1 + e^{-dr})^2
        real h = (real)2 * e.decay[z] * ex / (((real)1 + ex) * ((real)1 + ex));
        real others = 1;
#pragma unroll
        for (int y = 0; y < ZM; y++)
          if (y < e.nz && y != z) others *= lam_z[y];
        g0 += h * (x[0] - e.center[z][0]) / r_z[z] * others;
        g1 += h * (x[1] - e.center[z][1]) / r_z[z] * others;
      }
    }
    // f_x = I + u (grad lambda)^T, f_u = lambda I
    L.f_x[0] = (real)1 + u[0] * g0; L.f_x[1] = (real)0 + u[0] * g1;
    L.f_x[2] = (real)0 + u[1] * g0; L.f_x[3] = (real)1 + u[1] * g1;
    L.f_u[0] = lam; L.f_u[3] = lam;
#pragma unroll
    for (int i = 0; i < 2; i++) { L.l_x[i] = (real)2 * (x[i] - e.goal[i]); L.l_u[i] = 0; L.l_xx[i * 2 + i] = 2; }
  }
}

template <int KIND, int N, int M>
HD void env_final_quad(const EnvSmall &e, const real *x, real &l, real *l_x, real *l_xx) {  // diffenv.py:85-101
  l = env_final_cost<KIND, N, M>(e, x);
#pragma unroll
  for (int i = 0; i < N * N; i++) l_xx[i] = 0;
#pragma unroll
  for (int i = 0; i < N; i++) { l_x[i] = (real)2 * (x[i] - e.goal[i]); l_xx[i * N + i] = 2; }
}

// ------------------------------------------------------------------ dense helpers (static indexing only)
// Cholesky of the free block of H embedded in the full D x D matrix: clamped rows/columns are
// replaced by identity rows, which leaves the free block's factor bit-identical to the factor
// of the compacted H[free,free] while keeping every index compile-time (registers, no local
// memory).  Failure rule = Eigen LLT / LAPACK potrf: a non-positive (or NaN) pivot.
template <int D>
HD int chol_masked(const real *H, const bool *fr, real *L) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) L[i * D + j] = (fr[i] && fr[j]) ? H[i * D + j] : (i == j ? (real)1 : (real)0);
  int fail = 0;
#pragma unroll
  for (int j = 0; j < D; j++) {
    real s = L[j * D + j];
#pragma unroll
    for (int k = 0; k < j; k++) s -= L[j * D + k] * L[j * D + k];
    if (!(s > 0)) { fail = 1; s = 1; }
    real dj = r_sqrt(s);
    L[j * D + j] = dj;
#pragma unroll
    for (int i = j + 1; i < D; i++) {
      real t = L[i * D + j];
#pragma unroll
      for (int k = 0; k < j; k++) t -= L[i * D + k] * L[j * D + k];
      L[i * D + j] = t / dj;
    }
  }
  return fail;
}

template <int D>
HD void chol_solve(const real *L, real *b) {  // L L^T y = b, in place
#pragma unroll
  for (int i = 0; i < D; i++) {
    real s = b[i];
#pragma unroll
    for (int k = 0; k < i; k++) s -= L[i * D + k] * b[k];
    b[i] = s / L[i * D + i];
  }
#pragma unroll
  for (int i = D - 1; i >= 0; i--) {
    real s = b[i];
#pragma unroll
    for (int k = i + 1; k < D; k++) s -= L[k * D + i] * b[k];
    b[i] = s / L[i * D + i];
  }
}

// ------------------------------------------------------------------ box-QP
// tfmpc/utils/optimization.py:6-101 (projected_newton_qp) and :121-127 (_get_qp_indices).
template <int M>
HD real qp_value(const real *H, const real *q, const real *x) {  // :8-11
  real quad = 0, lin = 0;
#pragma unroll
  for (int i = 0; i < M; i++) {
    real hx = 0;
#pragma unroll
    for (int j = 0; j < M; j++) hx += H[i * M + j] * x[j];
    quad += x[i] * hx;
    lin += q[i] * x[i];
  }
  return (real)0.5 * quad + lin;
}

// x: in = start point, out = solution.  L: masked Cholesky factor of H[free,free].
// fr: free flags at exit.  Returns 0, or 2 if a factorisation failed (:47-51).
template <int M>
HD int boxqp(const real *H, const real *q, const real *lo, const real *hi, real *x, real *L, bool *fr) {
  const real rtol = (real)1e-8, armijo = (real)0.1, eps = (real)1e-6;
  bool clamped[M];
  real g[M], search[M], xc[M];
#pragma unroll
  for (int i = 0; i < M; i++) { clamped[i] = false; fr[i] = true; }
  real value = qp_value<M>(H, q, x), old_value = 0;
  int status = 0;
  for (int it = 0; it < 100; it++) {
    if (it > 0 && (old_value - value) < rtol * r_abs(old_value)) break;  // :27
    old_value = value;
    bool changed = false, allc = true;
#pragma unroll
    for (int i = 0; i < M; i++) {
      real s = 0;
#pragma unroll
      for (int j = 0; j < M; j++) s += H[i * M + j] * x[j];
      g[i] = q[i] + s;  // :34
    }
#pragma unroll
    for (int i = 0; i < M; i++) {  // :121-127
      bool c = (r_abs(x[i] - lo[i]) < eps && g[i] > 0) || (r_abs(hi[i] - x[i]) < eps && g[i] < 0);
      changed = changed || (c != clamped[i]);
      clamped[i] = c;
      fr[i] = !c;
      allc = allc && c;
    }
    if (it == 0 || changed) {  // :37-51
      if (chol_masked<M>(H, fr, L)) { status = 2; break; }
    }
    if (allc) break;  // :53
    real gn = 0;
#pragma unroll
    for (int i = 0; i < M; i++) gn += fr[i] ? g[i] * g[i] : (real)0;
    if (r_sqrt(gn) < eps) break;  // :58-62
    real rhs[M];
#pragma unroll
    for (int i = 0; i < M; i++) {  // grad_clamped = q + H (x * clamped), :65
      real s = 0;
#pragma unroll
      for (int j = 0; j < M; j++) s += H[i * M + j] * (clamped[j] ? x[j] : (real)0);
      rhs[i] = fr[i] ? q[i] + s : (real)0;
    }
    chol_solve<M>(L, rhs);
    real sdotg = 0;
#pragma unroll
    for (int i = 0; i < M; i++) {
      search[i] = fr[i] ? -rhs[i] - x[i] : (real)0;  // :70
      sdotg += search[i] * g[i];
    }
    if (sdotg >= 0) break;  // :75-79
    double step = 1.0;
    real vc;
    for (;;) {  // :82-95
      real st = (real)step;
      bool moved = false;
#pragma unroll
      for (int i = 0; i < M; i++) { xc[i] = r_clip(x[i] + st * search[i], lo[i], hi[i]); moved = moved || (xc[i] != x[i]); }
      if (!moved) {
        // Degenerate backtracking, fast-forwarded.  The candidate equals x bit for bit, so f(xc) == old_value
        // exactly, the Armijo ratio is -0 < 0.1, and every further halving of the step reproduces the same xc:
        // the reference spins here until step < 1e-22 (100 evaluations) and leaves with x and value unchanged.
        // About 2% of the box-QPs of the navigation workload end this way ("box-QP").
        vc = old_value;
        break;
      }
      vc = qp_value<M>(H, q, xc);
      if (!((vc - old_value) / (st * sdotg) < armijo)) break;
      step *= 0.6;
      if (step < 1e-22) {  // the reference evaluates xc, vc once more, then gives up
        st = (real)step;
#pragma unroll
        for (int i = 0; i < M; i++) xc[i] = r_clip(x[i] + st * search[i], lo[i], hi[i]);
        vc = qp_value<M>(H, q, xc);
        break;
      }
    }
#pragma unroll
    for (int i = 0; i < M; i++) x[i] = xc[i];
    value = vc;
  }
  return status;
}


enum { QP_NEWTON = 0, QP_COOP = 1, QP_CLOSED = 2 };  // box-QP flavour of the constrained controller

#ifdef __CUDACC__
// Warp-synchronous projected-Newton box-QP: the same algorithm and the same arithmetic as boxqp() above, executed by
// all 32 lanes of a warp TOGETHER, one problem per lane (`active` = this lane has a QP to solve).
//
// Why: in the thread-per-problem backward pass ~7% of the Armijo searches need more than one trial step and ~2% of
// them backtrack for tens of steps; with one problem per lane the whole warp waits for its slowest lane, and ncu showed
// more than half of all warp instructions of the backward kernel issued inside that loop with ~2 of 32 lanes active.
// Here every lane tries the full step (k = 0) on its own problem; the lanes that must backtrack are then served one at
// a time by the WHOLE warp: the owner broadcasts its problem (shuffles), lane j evaluates trial step k0 + j, and a
// ballot finds the first step the sequential loop would have stopped at.  Results are bit-identical to boxqp():
// each trial evaluates exactly the expressions of the sequential iteration k, with the same step size (real)(0.6^k).
template <int M>
__device__ __forceinline__ int boxqp_warp(bool active, const real *steps, int klast, const real *H, const real *q, const real *lo,
                                          const real *hi, real *x, real *L, bool *fr) {
  constexpr unsigned FULL = 0xffffffffu;
  const real rtol = (real)1e-8, armijo = (real)0.1, eps = (real)1e-6;
  const int lane = threadIdx.x & 31;
  bool clamped[M];
  real g[M], search[M], xc[M];
#pragma unroll
  for (int i = 0; i < M; i++) { clamped[i] = false; fr[i] = true; xc[i] = x[i]; }
  real value = active ? qp_value<M>(H, q, x) : (real)0, old_value = 0;
  int status = 0;
  bool act = active;
  for (int it = 0; it < 100; it++) {
    if (!__any_sync(FULL, act)) break;
    real sdotg = 0, vc = value;
    bool need = false;
    if (act && it > 0 && (old_value - value) < rtol * r_abs(old_value)) act = false;  // :27
    if (act) {
      old_value = value;
      bool changed = false, allc = true;
#pragma unroll
      for (int i = 0; i < M; i++) {
        real s = 0;
#pragma unroll
        for (int j = 0; j < M; j++) s += H[i * M + j] * x[j];
        g[i] = q[i] + s;  // :34
      }
#pragma unroll
      for (int i = 0; i < M; i++) {  // :121-127
        bool c = (r_abs(x[i] - lo[i]) < eps && g[i] > 0) || (r_abs(hi[i] - x[i]) < eps && g[i] < 0);
        changed = changed || (c != clamped[i]);
        clamped[i] = c;
        fr[i] = !c;
        allc = allc && c;
      }
      if (it == 0 || changed) {  // :37-51
        if (chol_masked<M>(H, fr, L)) { status = 2; act = false; }
      }
      if (act && allc) act = false;  // :53
      if (act) {
        real gn = 0;
#pragma unroll
        for (int i = 0; i < M; i++) gn += fr[i] ? g[i] * g[i] : (real)0;
        if (r_sqrt(gn) < eps) act = false;  // :58-62
      }
      if (act) {
        real rhs[M];
#pragma unroll
        for (int i = 0; i < M; i++) {  // grad_clamped = q + H (x * clamped), :65
          real s = 0;
#pragma unroll
          for (int j = 0; j < M; j++) s += H[i * M + j] * (clamped[j] ? x[j] : (real)0);
          rhs[i] = fr[i] ? q[i] + s : (real)0;
        }
        chol_solve<M>(L, rhs);
#pragma unroll
        for (int i = 0; i < M; i++) {
          search[i] = fr[i] ? -rhs[i] - x[i] : (real)0;  // :70
          sdotg += search[i] * g[i];
        }
        if (sdotg >= 0) act = false;  // :75-79
      }
      if (act) {  // trial step k = 0 (step = 1), :82-85
        bool moved = false;
#pragma unroll
        for (int i = 0; i < M; i++) { xc[i] = r_clip(x[i] + (real)1 * search[i], lo[i], hi[i]); moved = moved || (xc[i] != x[i]); }
        if (!moved) vc = old_value;  // degenerate, see boxqp()
        else {
          vc = qp_value<M>(H, q, xc);
          need = (vc - old_value) / ((real)1 * sdotg) < armijo;
        }
      }
    }
    // cooperative backtracking for the lanes whose full step failed the Armijo test
    unsigned nm = __ballot_sync(FULL, need);
    while (nm) {
      const int src = __ffs(nm) - 1;
      nm &= nm - 1;
      real bH[M * M], bq[M], bx[M], bs[M], blo[M], bhi[M];
#pragma unroll
      for (int i = 0; i < M * M; i++) bH[i] = __shfl_sync(FULL, H[i], src);
#pragma unroll
      for (int i = 0; i < M; i++) {
        bq[i] = __shfl_sync(FULL, q[i], src); bx[i] = __shfl_sync(FULL, x[i], src); bs[i] = __shfl_sync(FULL, search[i], src);
        blo[i] = __shfl_sync(FULL, lo[i], src); bhi[i] = __shfl_sync(FULL, hi[i], src);
      }
      const real bold = __shfl_sync(FULL, old_value, src), bsd = __shfl_sync(FULL, sdotg, src);
      for (int k0 = 1; k0 <= klast; k0 += 32) {
        const int k = k0 + lane;
        const bool vk = k <= klast;
        const real st = steps[vk ? k : klast];
        real cxc[M];
        bool moved = false;
#pragma unroll
        for (int i = 0; i < M; i++) { cxc[i] = r_clip(bx[i] + st * bs[i], blo[i], bhi[i]); moved = moved || (cxc[i] != bx[i]); }
        const real cvc = qp_value<M>(bH, bq, cxc);
        const bool cond = (cvc - bold) / (st * bsd) < armijo;
        // the sequential loop stops at k if the trial did not move (degenerate), passed the test, or was the last one (:93-95)
        const unsigned tm = __ballot_sync(FULL, vk && (!moved || !cond || k == klast));
        if (tm) {
          const int f = __ffs(tm) - 1;
          const bool f_moved = __shfl_sync(FULL, moved, f);
          const real f_vc = __shfl_sync(FULL, cvc, f);
          real f_xc[M];
#pragma unroll
          for (int i = 0; i < M; i++) f_xc[i] = __shfl_sync(FULL, cxc[i], f);
          if (lane == src) {
            if (!f_moved) {
#pragma unroll
              for (int i = 0; i < M; i++) xc[i] = x[i];
              vc = old_value;
            } else {
#pragma unroll
              for (int i = 0; i < M; i++) xc[i] = f_xc[i];
              vc = f_vc;
            }
          }
          break;
        }
      }
    }
    if (act) {
#pragma unroll
      for (int i = 0; i < M; i++) x[i] = xc[i];
      value = vc;
    }
  }
  return status;
}
#endif  // __CUDACC__

// ------------------------------------------------------------------ iLQR backward, one timestep
// Structural zeros of the analytic linearisation, known at compile time per environment.  Skipping a
// product with a structural zero (or the addition of one) is exact, so the result equals the dense
// product the reference forms -- only the instruction count changes.
template <int KIND>
struct Traits {
  static constexpr bool fu_diag = true;                       // f_u = I (NavigationLQR) or lambda I (Navigation)
  static constexpr bool fx_diag = (KIND == TFMPC_ENV_NAVLQR);  // f_x = I
  static constexpr bool lxx_diag = true, luu_diag = true, lxu_zero = true;
};

// The Q block of one timestep (ilqr.py:122-134) as one flat array, so that the warp-cooperative ("solo") backward of
// queue_core.cuh can hand it from the lanes that assemble it to the lanes that consume it through shared memory.
template <int N, int M>
struct QB {
  static constexpr int OX = 0, OU = N, OXX = N + M, OUU = OXX + N * N, OUX = OUU + M * M, OUUR = OUX + M * N, OUXR = OUUR + M * M,
                       SIZE = OUXR + M * N;
  real v[SIZE];
  HD real *Q_x() { return v + OX; }
  HD real *Q_u() { return v + OU; }
  HD real *Q_xx() { return v + OXX; }
  HD real *Q_uu() { return v + OUU; }
  HD real *Q_ux() { return v + OUX; }
  HD real *Q_uu_reg() { return v + OUUR; }
  HD real *Q_ux_reg() { return v + OUXR; }
};

// ---- stage 1: Q_x, Q_u (:122-123), Q_xx, Q_uu, Q_ux (:129-131) and the state-regularised Q_uu_reg, Q_ux_reg (:127,133-134).
// The sums are explicit fma chains (r_fma): the solo engine evaluates the same chains entry by entry on different lanes, and
// the two code shapes must round identically in the fp32 build for results to be independent of the schedule.
template <int KIND, int N, int M>
HD void assemble_q(const Lin<N, M> &L, real mu, const real *V_x, const real *V_xx, QB<N, M> &q) {
  typedef Traits<KIND> TR;
  real *Q_x = q.Q_x(), *Q_u = q.Q_u(), *Q_xx = q.Q_xx(), *Q_uu = q.Q_uu(), *Q_ux = q.Q_ux(), *Q_uu_reg = q.Q_uu_reg(), *Q_ux_reg = q.Q_ux_reg();
  real fxTV[N * N], fuTV[M * N], fuTVr[M * N];
#pragma unroll
  for (int i = 0; i < N; i++) {  // :122
    real s = 0;
#pragma unroll
    for (int p = 0; p < N; p++) {
      if (TR::fx_diag && p != i) continue;
      s = r_fma(L.f_x[p * N + i], V_x[p], s);
    }
    Q_x[i] = L.l_x[i] + s;
  }
#pragma unroll
  for (int i = 0; i < M; i++) {  // :123
    real s = 0;
#pragma unroll
    for (int p = 0; p < N; p++) {
      if (TR::fu_diag && p != i) continue;
      s = r_fma(L.f_u[p * M + i], V_x[p], s);
    }
    Q_u[i] = L.l_u[i] + s;
  }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {  // :125
      real s = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        if (TR::fx_diag && p != i) continue;
        s = r_fma(L.f_x[p * N + i], V_xx[p * N + j], s);
      }
      fxTV[i * N + j] = s;
    }
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {  // :126-127
      real s = 0, sr = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        if (TR::fu_diag && p != i) continue;
        s = r_fma(L.f_u[p * M + i], V_xx[p * N + j], s);
        sr = r_fma(L.f_u[p * M + i], (p == j ? V_xx[p * N + j] + mu * (real)1 : V_xx[p * N + j]), sr);
      }
      fuTV[i * N + j] = s;
      fuTVr[i * N + j] = sr;
    }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) {  // :129
      real s = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        if (TR::fx_diag && p != j) continue;
        s = r_fma(fxTV[i * N + p], L.f_x[p * N + j], s);
      }
      Q_xx[i * N + j] = (TR::lxx_diag && i != j) ? s : L.l_xx[i * N + j] + s;
    }
#pragma unroll
  for (int i = 0; i < M; i++) {
#pragma unroll
    for (int j = 0; j < M; j++) {  // :130, :133
      real s = 0, sr = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        if (TR::fu_diag && p != j) continue;
        s = r_fma(fuTV[i * N + p], L.f_u[p * M + j], s);
        sr = r_fma(fuTVr[i * N + p], L.f_u[p * M + j], sr);
      }
      Q_uu[i * M + j] = (TR::luu_diag && i != j) ? s : L.l_uu[i * M + j] + s;
      Q_uu_reg[i * M + j] = (TR::luu_diag && i != j) ? sr : L.l_uu[i * M + j] + sr;
    }
#pragma unroll
    for (int j = 0; j < N; j++) {  // :131, :134 (l_xu^T)
      real s = 0, sr = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        if (TR::fx_diag && p != j) continue;
        s = r_fma(fuTV[i * N + p], L.f_x[p * N + j], s);
        sr = r_fma(fuTVr[i * N + p], L.f_x[p * N + j], sr);
      }
      Q_ux[i * N + j] = TR::lxu_zero ? s : L.l_xu[j * M + i] + s;
      Q_ux_reg[i * N + j] = TR::lxu_zero ? sr : L.l_xu[j * M + i] + sr;
    }
  }
}

// Closed-form constrained controller for M <= 2 (QP_CLOSED): box-QP solution k AND feedback gain K in one go.
// NOT the reference's iteration: k is the exact minimiser of the strictly convex QP over the box, which is what
// optimization.py:6-101 converges to (the projected-Newton loop stops on a relative decrease < 1e-8 or a free-gradient
// norm < 1e-6), without the data-dependent loop and without a Cholesky factor.
// Why two candidates suffice for M = 2: with xu the unconstrained minimiser, the box minimiser lies on an edge whose
// bound xu violates (KKT on a non-violated edge's interior forces x = xu there), i.e. at (b0, clip(argmin_x1 f(b0, .)))
// or (clip(argmin_x0 f(., b1)), b1) with b = clip(xu); when both bounds are violated the lower objective wins.
// The free / clamped flags follow optimization.py:121-127 evaluated at the solution; the factorisation-failure rule of
// :37-51 (a non-positive Cholesky pivot of the full H) is the equivalent "h00 <= 0 or det <= 0";
// K[free rows] = -H_ff^-1 Q_ux_reg[free] (ilqr.py:375-383) uses the adjugate inverse when both rows are free and a plain
// reciprocal when one is.  The three reciprocals are independent of each other, so the whole controller is ~70
// instructions with a dependent chain of one reciprocal + ~25 FMA-class operations (the Cholesky + triangular solves of
// the generic path chain four divisions and two square roots per solve; the iteration costs ~1,400 instructions).
// Returns 0, or 2 when H is not positive definite (k stays at the start point, K = 0, as in the reference).
// Parity: same iteration count as the fp64 oracle on >= 99 % of C3 problems, fp32 inside the fp32-vs-fp64 noise band
// (tests/test_device_logic_emulation.py::test_closed_form_qp_*).
template <int N, int M>
HD int controller_closed(const real *H, const real *q, const real *lo, const real *hi, const real *B, real *k, real *K) {
  static_assert(M <= 2, "closed-form controller: M <= 2");
  const real eps = (real)1e-6;
  if (M == 1) {
    const real h = H[0];
    if (!(h > 0)) {
#pragma unroll
      for (int j = 0; j < N; j++) K[j] = 0;
      return 2;
    }
    const real r = (real)1 / h;
    const real x = r_clip(-(q[0] * r), lo[0], hi[0]);
    const real g = q[0] + h * x;
    const bool c = (r_abs(x - lo[0]) < eps && g > 0) || (r_abs(hi[0] - x) < eps && g < 0);
    k[0] = x;
#pragma unroll
    for (int j = 0; j < N; j++) K[j] = c ? (real)0 : -(B[j] * r);
    return 0;
  } else {
    const real h00 = H[0], h01 = H[1], h10 = H[2], h11 = H[3];
    const real det = h00 * h11 - h10 * h10;
    if (!(h00 > 0) || !(det > 0)) {
#pragma unroll
      for (int i = 0; i < M * N; i++) K[i] = 0;
      return 2;
    }
    const real rd = (real)1 / det, r0 = (real)1 / h00, r1 = (real)1 / h11;
    const real xu0 = -((h11 * q[0] - h10 * q[1]) * rd), xu1 = -((h00 * q[1] - h10 * q[0]) * rd);
    const real b0 = r_clip(xu0, lo[0], hi[0]), b1 = r_clip(xu1, lo[1], hi[1]);
    const bool v0 = xu0 != b0, v1 = xu1 != b1;
    real xa[2], xb[2];
    xa[0] = b0; xa[1] = r_clip(-((q[1] + h10 * b0) * r1), lo[1], hi[1]);
    xb[1] = b1; xb[0] = r_clip(-((q[0] + h01 * b1) * r0), lo[0], hi[0]);
    bool takeA = v0;
    if (v0 && v1) takeA = qp_value<2>(H, q, xa) <= qp_value<2>(H, q, xb);
    const bool inside = !v0 && !v1;
    const real x0 = inside ? xu0 : (takeA ? xa[0] : xb[0]), x1 = inside ? xu1 : (takeA ? xa[1] : xb[1]);
    const real g0 = q[0] + (h00 * x0 + h01 * x1), g1 = q[1] + (h10 * x0 + h11 * x1);
    const bool c0 = (r_abs(x0 - lo[0]) < eps && g0 > 0) || (r_abs(hi[0] - x0) < eps && g0 < 0);   // optimization.py:121-127 at the solution
    const bool c1 = (r_abs(x1 - lo[1]) < eps && g1 > 0) || (r_abs(hi[1] - x1) < eps && g1 < 0);
    k[0] = x0; k[1] = x1;
#pragma unroll
    for (int j = 0; j < N; j++) {
      const real B0 = B[j], B1 = B[N + j];
      const real f0 = -((h11 * B0 - h10 * B1) * rd), f1 = -((h00 * B1 - h10 * B0) * rd);   // both rows free
      K[j] = c0 ? (real)0 : (c1 ? -(B0 * r0) : f0);
      K[N + j] = c1 ? (real)0 : (c0 ? -(B1 * r1) : f1);
    }
    return 0;
  }
}

// ---- stage 2: the controller, ilqr.py:136-143 -> :357-362 (unconstrained) / :364-387 (constrained) / :139-141 (bang-bang).
// Returns 0, 1 (unconstrained Cholesky failed) or 2 (box-QP failed).  V_xx is the value function BEFORE this step's update.
// QP = QP_COOP: called by all 32 lanes of a warp in lock step (device only); the box-QP then runs warp-cooperatively.
// QP = QP_CLOSED (M <= 2): the closed-form controller above instead of the reference's iteration.
template <int KIND, int N, int M, int QP>
HD int controller(const EnvSmall &e, QB<N, M> &q, const real *V_xx, const real *u, real *K, real *k) {
  real *Q_u = q.Q_u(), *Q_uu_reg = q.Q_uu_reg(), *Q_ux_reg = q.Q_ux_reg();
  int status = 0;
  if (e.bounded) {  // :136
    bool any_nz = false;
#pragma unroll
    for (int i = 0; i < N * N; i++) any_nz = any_nz || (V_xx[i] != 0);
    bool enter_qp = any_nz;
#ifdef __CUDA_ARCH__
    if (QP == QP_COOP) enter_qp = __any_sync(0xffffffffu, any_nz);  // warp-uniform: every lane enters, lanes without a QP idle inside
#endif
    real lo[M], hi[M], Lf[M * M];
    bool fr[M];
    int st = 0;
    if constexpr (QP == QP_CLOSED && M <= 2) {
      if (any_nz) {  // :137-138 -> _get_constrained_controller :364-387
#pragma unroll
        for (int i = 0; i < M; i++) { lo[i] = e.low[i] - u[i]; hi[i] = e.high[i] - u[i]; k[i] = (lo[i] + hi[i]) / (real)2; }
        if (controller_closed<N, M>(Q_uu_reg, Q_u, lo, hi, Q_ux_reg, k, K)) status = 2;
        return status;
      }
    } else {
      if (enter_qp) {  // :137-138 -> _get_constrained_controller :364-387
#pragma unroll
        for (int i = 0; i < M; i++) { lo[i] = e.low[i] - u[i]; hi[i] = e.high[i] - u[i]; k[i] = (lo[i] + hi[i]) / (real)2; }
#ifdef __CUDA_ARCH__
        if (QP == QP_COOP) st = boxqp_warp<M>(any_nz, e.qp_steps, e.qp_klast, Q_uu_reg, Q_u, lo, hi, k, Lf, fr);
        else
#endif
        st = boxqp<M>(Q_uu_reg, Q_u, lo, hi, k, Lf, fr);
      }
    }
    if (any_nz) {
      if (st) status = 2;
#pragma unroll
      for (int j = 0; j < N; j++) {  // K[free] = -cholesky_solve(Hfree, Q_ux_reg[free]); clamped rows 0
        real col[M];
#pragma unroll
        for (int i = 0; i < M; i++) col[i] = (fr[i] && !st) ? Q_ux_reg[i * N + j] : (real)0;
        chol_solve<M>(Lf, col);
#pragma unroll
        for (int i = 0; i < M; i++) K[i * N + j] = (fr[i] && !st) ? -col[i] : (real)0;
      }
    } else {  // :139-141 bang-bang
#pragma unroll
      for (int i = 0; i < M * N; i++) K[i] = 0;
#pragma unroll
      for (int i = 0; i < M; i++) k[i] = (Q_u[i] >= 0) ? e.low[i] - u[i] : e.high[i] - u[i];
    }
  } else {  // :143 -> _get_unconstrained_controller :357-362
    real R[M * M];
    bool all[M];
#pragma unroll
    for (int i = 0; i < M; i++) all[i] = true;
    if (chol_masked<M>(Q_uu_reg, all, R)) return 1;
#pragma unroll
    for (int i = 0; i < M; i++) k[i] = Q_u[i];
    chol_solve<M>(R, k);
#pragma unroll
    for (int i = 0; i < M; i++) k[i] = -k[i];
#pragma unroll
    for (int j = 0; j < N; j++) {
      real col[M];
#pragma unroll
      for (int i = 0; i < M; i++) col[i] = Q_ux_reg[i * N + j];
      chol_solve<M>(R, col);
#pragma unroll
      for (int i = 0; i < M; i++) K[i * N + j] = -col[i];
    }
  }
  return status;
}

// ---- stage 3: value update with the UNregularised Q (:145-162), J (:164), dV1, dV2 (:166-167)
template <int N, int M>
HD void value_update(QB<N, M> &q, const real *K, const real *k, real l, real *V_x, real *V_xx, real &J, real &dV1, real &dV2) {
  const real *Q_x = q.Q_x(), *Q_u = q.Q_u(), *Q_xx = q.Q_xx(), *Q_uu = q.Q_uu(), *Q_ux = q.Q_ux();
  real KtQuu[N * M];
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < M; j++) {
      real s = 0;
#pragma unroll
      for (int p = 0; p < M; p++) s += K[p * N + i] * Q_uu[p * M + j];
      KtQuu[i * M + j] = s;
    }
  real Vn[N * N];
#pragma unroll
  for (int i = 0; i < N; i++) {
    real a1 = 0, a2 = 0, a3 = 0;
#pragma unroll
    for (int p = 0; p < M; p++) { a1 += Q_ux[p * N + i] * k[p]; a2 += K[p * N + i] * Q_u[p]; a3 += KtQuu[i * M + p] * k[p]; }
    V_x[i] = Q_x[i] + a1 + a2 + a3;
#pragma unroll
    for (int j = 0; j < N; j++) {
      real b1 = 0, b2 = 0, b3 = 0;
#pragma unroll
      for (int p = 0; p < M; p++) { b1 += Q_ux[p * N + i] * K[p * N + j]; b2 += K[p * N + i] * Q_ux[p * N + j]; b3 += KtQuu[i * M + p] * K[p * N + j]; }
      Vn[i * N + j] = Q_xx[i * N + j] + b1 + b2 + b3;
    }
  }
#pragma unroll
  for (int i = 0; i < N; i++)
#pragma unroll
    for (int j = 0; j < N; j++) V_xx[i * N + j] = (real)0.5 * (Vn[i * N + j] + Vn[j * N + i]);  // :162
  J += l;  // :164
  real d1 = 0, d2 = 0;
#pragma unroll
  for (int i = 0; i < M; i++) d1 += k[i] * Q_u[i];
  dV1 += d1;  // :166
#pragma unroll
  for (int j = 0; j < M; j++) {
    real s = 0;
#pragma unroll
    for (int i = 0; i < M; i++) s += k[i] * Q_uu[i * M + j];
    d2 += s * k[j];
  }
  dV2 += (real)0.5 * d2;  // :167
}

// tfmpc/solvers/ilqr.py:108-170 with the controllers of :357-387: the three stages above for one timestep in one thread.
// V_x, V_xx, J, dV1, dV2 are carried across timesteps.  Returns 0, 1 (unconstrained Cholesky failed) or 2 (box-QP failed).
template <int KIND, int N, int M, int QP = QP_NEWTON>
HD int backward_step(const EnvSmall &e, const Lin<N, M> &L, const real *u, real mu, real *V_x, real *V_xx, real &J, real &dV1,
                     real &dV2, real *K, real *k) {
  QB<N, M> q;
  assemble_q<KIND, N, M>(L, mu, V_x, V_xx, q);
  const int status = controller<KIND, N, M, QP>(e, q, V_xx, u, K, k);
  if (status == 1) return 1;
  value_update<N, M>(q, K, k, L.l, V_x, V_xx, J, dV1, dV2);
  return status;
}

// ------------------------------------------------------------------ trajectory / gain accessors
// The passes below read and write whole per-timestep RECORDS through two accessor concepts:
//   trajectory:  load_xu(t, x, u)  load_x(t, x)  store_xu(t, x, u)  store_x(t, x)      (t = T holds x only)
//   gains:       load(t, K, k)  store(t, K, k)
// Two implementations:
//  * Strided*  -- plain arrays addressed as base[row * stride]; stride 1 = the reference's dense [T, n] layout of one
//                 problem (stage kernels, host emulation).
//  * Vec*      -- the solve workspace: 16-byte (4 x real) chunks addressed as base[t * ts + c * cs].  One timestep of one
//                 problem moves with CH 128-bit accesses (LDG.128 / STG.128).  The kernels choose the strides so that what
//                 one access pattern touches together is contiguous: nominal trajectories [t][slot] (a warp of adjacent
//                 problems reads 512 contiguous bytes), gains [t][slot][chunk] (one problem-step = one 32-byte sector),
//                 line-search candidates [t][slot][lane] (the 4 candidates of a problem-step = two full sectors).
struct alignas(4 * sizeof(real)) R4 { real v[4]; };

template <int N, int M>
struct StridedTraj {
  real *X, *U;       // X[(t * N + i) * stride], U[(t * M + i) * stride]
  int64_t stride;
  HD void load_xu(int t, real *x, real *u) const {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = X[(int64_t)(t * N + i) * stride];
#pragma unroll
    for (int i = 0; i < M; i++) u[i] = U[(int64_t)(t * M + i) * stride];
  }
  HD void load_x(int t, real *x) const {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = X[(int64_t)(t * N + i) * stride];
  }
  HD void store_xu(int t, const real *x, const real *u) const {
#pragma unroll
    for (int i = 0; i < N; i++) X[(int64_t)(t * N + i) * stride] = x[i];
#pragma unroll
    for (int i = 0; i < M; i++) U[(int64_t)(t * M + i) * stride] = u[i];
  }
  HD void store_x(int t, const real *x) const {
#pragma unroll
    for (int i = 0; i < N; i++) X[(int64_t)(t * N + i) * stride] = x[i];
  }
};

template <int N, int M>
struct StridedGain {
  real *K, *k;       // K[(t * M * N + i) * stride], k[(t * M + i) * stride]
  int64_t stride;
  HD void load(int t, real *Kt, real *kt) const {
#pragma unroll
    for (int i = 0; i < M * N; i++) Kt[i] = K[(int64_t)(t * M * N + i) * stride];
#pragma unroll
    for (int i = 0; i < M; i++) kt[i] = k[(int64_t)(t * M + i) * stride];
  }
  HD void store(int t, const real *Kt, const real *kt) const {
#pragma unroll
    for (int i = 0; i < M * N; i++) K[(int64_t)(t * M * N + i) * stride] = Kt[i];
#pragma unroll
    for (int i = 0; i < M; i++) k[(int64_t)(t * M + i) * stride] = kt[i];
  }
};

// record t, chunk c lives at base[t * ts + c * cs] (ts, cs in R4 units)
template <int CNT>
HD void vec_load(const R4 *base, int64_t ts, int64_t cs, int t, real *out) {  // CNT reals of record t
  constexpr int CH = (CNT + 3) / 4;
#pragma unroll
  for (int c = 0; c < CH; c++) {
    const R4 r = base[(int64_t)t * ts + c * cs];
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (c * 4 + j < CNT) out[c * 4 + j] = r.v[j];
  }
}
template <int CNT>
HD void vec_store(R4 *base, int64_t ts, int64_t cs, int t, const real *in) {
  constexpr int CH = (CNT + 3) / 4;
#pragma unroll
  for (int c = 0; c < CH; c++) {
    R4 r;
#pragma unroll
    for (int j = 0; j < 4; j++) r.v[j] = (c * 4 + j < CNT) ? in[c * 4 + j] : (real)0;
    base[(int64_t)t * ts + c * cs] = r;
  }
}

template <int N, int M>
struct VecTraj {     // record t = [x (N), u (M)]
  R4 *base;          // already offset to this problem's slot (and candidate lane)
  int64_t ts, cs;    // strides between timesteps / between the chunks of one record
  static constexpr int CH = (N + M + 3) / 4;
  HD void load_xu(int t, real *x, real *u) const {
    real r[N + M];
    vec_load<N + M>(base, ts, cs, t, r);
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = r[i];
#pragma unroll
    for (int i = 0; i < M; i++) u[i] = r[N + i];
  }
  HD void load_x(int t, real *x) const {
    real r[N + M];
    vec_load<N + M>(base, ts, cs, t, r);
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = r[i];
  }
  HD void store_xu(int t, const real *x, const real *u) const {
    real r[N + M];
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = x[i];
#pragma unroll
    for (int i = 0; i < M; i++) r[N + i] = u[i];
    vec_store<N + M>(base, ts, cs, t, r);
  }
  HD void store_x(int t, const real *x) const {
    real r[N + M];
#pragma unroll
    for (int i = 0; i < N; i++) r[i] = x[i];
#pragma unroll
    for (int i = 0; i < M; i++) r[N + i] = 0;
    vec_store<N + M>(base, ts, cs, t, r);
  }
};

template <int N, int M>
struct VecGain {     // record t = [K (M*N), k (M)]
  R4 *base;
  int64_t ts, cs;
  static constexpr int CH = (M * N + M + 3) / 4;
  HD void load(int t, real *Kt, real *kt) const {
    real r[M * N + M];
    vec_load<M * N + M>(base, ts, cs, t, r);
#pragma unroll
    for (int i = 0; i < M * N; i++) Kt[i] = r[i];
#pragma unroll
    for (int i = 0; i < M; i++) kt[i] = r[M * N + i];
  }
  HD void store(int t, const real *Kt, const real *kt) const {
    real r[M * N + M];
#pragma unroll
    for (int i = 0; i < M * N; i++) r[i] = Kt[i];
#pragma unroll
    for (int i = 0; i < M; i++) r[M * N + i] = kt[i];
    vec_store<M * N + M>(base, ts, cs, t, r);
  }
};

// optional per-timestep cost sink (stage API); p == nullptr discards
struct CostSink {
  real *p;
  int64_t stride;
  HD void put(int t, real c) const { if (p) p[(int64_t)t * stride] = c; }
};

// iLQR.backward over the whole horizon (ilqr.py:94-172), linearisation fused (ilqr.py:84-92).
// Also accumulates sum_t max_i |k|/(|u|+1) for the g_norm test of ilqr.py:243.
template <int KIND, int N, int M, int QP = QP_NEWTON, class TJ, class GN>
HD int backward_pass(const EnvSmall &e, int T, const TJ &nom, real mu, const GN &gain, real &J, real &dV1, real &dV2, real &gsum) {
  real V_x[N], V_xx[N * N], x[N], u[M];
  nom.load_x(T, x);
  env_final_quad<KIND, N, M>(e, x, J, V_x, V_xx);  // :101-104
  dV1 = 0; dV2 = 0; gsum = 0;
  int status = 0;
  real xn[N], un[M];
  nom.load_xu(T - 1, xn, un);
  for (int t = T - 1; t >= 0; t--) {
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = xn[i];
#pragma unroll
    for (int i = 0; i < M; i++) u[i] = un[i];
    nom.load_xu(t > 0 ? t - 1 : 0, xn, un);  // software prefetch of the next (earlier) timestep
    Lin<N, M> L;
    env_linearize<KIND, N, M>(e, x, u, L);
    real K[M * N], k[M];
    int st = backward_step<KIND, N, M, QP>(e, L, u, mu, V_x, V_xx, J, dV1, dV2, K, k);
    if (st == 1 && QP != QP_COOP) return 1;  // (QP_COOP is used for bounded envs only: the unconstrained failure cannot occur)
    if (st) status = st;
    real mx = 0;
#pragma unroll
    for (int i = 0; i < M; i++) {
      real v = r_abs(k[i]) / (r_abs(u[i]) + (real)1.0);
      mx = (i == 0 || v > mx) ? v : mx;
    }
    gsum += mx;
    gain.store(t, K, k);
  }
  return status;
}

// iLQR.forward (ilqr.py:174-212).
template <int N, int M>
struct NomRec { real xh[N], uh[M], K[M * N], k[M]; };  // nominal state/action and gains of one timestep

template <int KIND, int N, int M, class TO>
HD void forward_step(const EnvSmall &e, real alpha, const NomRec<N, M> &r, int t, real *x, const TO &out, const CostSink &Co, real &J,
                     real &residual) {
  real u[M], xn[N];
#pragma unroll
  for (int i = 0; i < M; i++) {
    real s = 0;
#pragma unroll
    for (int j = 0; j < N; j++) s += r.K[i * N + j] * (x[j] - r.xh[j]);
    real du = alpha * r.k[i] + s;                             // :194
    u[i] = r_clip(r.uh[i] + du, e.low[i], e.high[i]);         // :196-197
    residual = r_max(residual, r_abs(du));                    // :206 (pre-clip)
  }
  real c = env_cost<KIND, N, M>(e, x, u);
  env_step<KIND, N, M>(e, x, u, xn);
  out.store_xu(t, x, u);
  Co.put(t, c);
  J += c;
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = xn[i];
}

// (Stage API and host emulation; the solve's line search uses rollout_staged() in ilqr_small.cu instead.)
// The records are loaded TWO steps ahead of their use through a 3-slot register ring; the loop is unrolled by 3 so
// that the ring needs no register-to-register rotation (a rotating MOV would wait on the load it copies from and
// collapse the prefetch distance to one step -- seen as 73% of the stall samples of the line-search kernel in ncu).
// One step of arithmetic is shorter than a DRAM / far-L2 round trip, hence the distance of two.
template <int KIND, int N, int M, class TJ, class GN, class TO>
HD void forward_pass(const EnvSmall &e, int T, const TJ &nom, const GN &gain, real alpha, const TO &out, const CostSink &Co, real &J,
                     real &residual) {
  real x[N];
  NomRec<N, M> r0, r1, r2;
  J = 0; residual = 0;
  const int last = T - 1;
  nom.load_xu(0, r0.xh, r0.uh); gain.load(0, r0.K, r0.k);
  nom.load_xu(1 < last ? 1 : last, r1.xh, r1.uh); gain.load(1 < last ? 1 : last, r1.K, r1.k);
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = r0.xh[i];
  for (int t = 0; t < T; t += 3) {
    { const int tl = t + 2 < last ? t + 2 : last; nom.load_xu(tl, r2.xh, r2.uh); gain.load(tl, r2.K, r2.k); }
    forward_step<KIND, N, M>(e, alpha, r0, t, x, out, Co, J, residual);
    if (t + 1 < T) {
      { const int tl = t + 3 < last ? t + 3 : last; nom.load_xu(tl, r0.xh, r0.uh); gain.load(tl, r0.K, r0.k); }
      forward_step<KIND, N, M>(e, alpha, r1, t + 1, x, out, Co, J, residual);
    }
    if (t + 2 < T) {
      { const int tl = t + 4 < last ? t + 4 : last; nom.load_xu(tl, r1.xh, r1.uh); gain.load(tl, r1.K, r1.k); }
      forward_step<KIND, N, M>(e, alpha, r2, t + 2, x, out, Co, J, residual);
    }
  }
  out.store_x(T, x);
  real cf = env_final_cost<KIND, N, M>(e, x);
  Co.put(T, cf);
  J += cf;
}

// iLQR.start with pinned actions (ilqr.py:53-82); u_init[t * M + i] is the reference's dense layout
template <int KIND, int N, int M, class TO>
HD void start_pass(const EnvSmall &e, int T, const real *x0, const real *u_init, const TO &out, const CostSink &Co) {
  real x[N], u[M], xn[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = x0[i];
  for (int t = 0; t < T; t++) {
#pragma unroll
    for (int i = 0; i < M; i++) u[i] = u_init[t * M + i];
    Co.put(t, env_cost<KIND, N, M>(e, x, u));
    env_step<KIND, N, M>(e, x, u, xn);
    out.store_xu(t, x, u);
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = xn[i];
  }
  out.store_x(T, x);
  Co.put(T, env_final_cost<KIND, N, M>(e, x));
}

// ------------------------------------------------------------------ the solve, as per-problem TICKS
// iLQR.solve (ilqr.py:214-283) + _backward (:285-315) + _forward (:317-355) cut into the pieces the
// kernels run once per "tick": one backward pass, one line search, one schedule update.  A tick is one
// turn of the reference's inner `while True` (:238); the outer iteration index advances only when a
// step is accepted, exactly as in the reference.  The same functions are composed sequentially by
// solve_one() below (used by the host-emulation harness) and in parallel by the kernels.
enum { PH_SEARCH = 0, PH_DONE = 1 };

struct Prob {            // per-problem solver state
  double mu, delta;      // python floats in the reference (ilqr.py:215-216)
  int iteration, n_bwd, n_fwd, status, cur, phase, guard;
  real J_hat, dV1, dV2;
};

HD void prob_init(Prob &p) {
  p.mu = 0.0; p.delta = 1.0; p.iteration = 0; p.n_bwd = 0; p.n_fwd = 0; p.status = TFMPC_ST_MAXITER; p.cur = 0; p.phase = PH_SEARCH;
  p.guard = 0; p.J_hat = 0; p.dV1 = 0; p.dV2 = 0;
}

// _backward (:285-315) + the g_norm test (:243-248).  Leaves p.phase = PH_SEARCH if a line search must follow.
template <int KIND, int N, int M, int QP = QP_NEWTON, class TJ, class GN>
HD void tick_backward(const EnvSmall &e, const IlqrOpts &o, int T, const TJ &nom, const GN &gain, Prob &p) {
  real gsum;
  double mu_l = p.mu, delta_l = p.delta;  // the retry bump is local, ilqr.py:308-309,315
  int bst, tries = 0;
  for (;;) {
    bst = backward_pass<KIND, N, M, QP>(e, T, nom, (real)mu_l, gain, p.J_hat, p.dV1, p.dV2, gsum);
    p.n_bwd++;
    if (bst != 1 || ++tries > 200) break;
    delta_l = fmax(o.delta_0, delta_l * o.delta_0);
    mu_l = fmax(o.mu_min, mu_l * delta_l);
  }
  p.phase = PH_SEARCH;
  if (bst) { p.status = TFMPC_ST_NONPD; p.phase = PH_DONE; return; }
  real g = gsum / (real)T;  // :243
  if (!(g == g)) { p.status = TFMPC_ST_NAN; p.phase = PH_DONE; return; }
  if (g < o.atol) { p.status = TFMPC_ST_CONVERGED; p.phase = PH_DONE; }  // :245-248 (nominal kept)
}

// acceptance test of one candidate, _forward :339-351
HD bool ls_accepts(const IlqrOpts &o, real alpha, real J_hat, real dV1, real dV2, real J) {
  real delta_J = -alpha * (dV1 + alpha * dV2);
  real dcost = J_hat - J;
  real z = (delta_J > 0) ? dcost / delta_J : r_sgn(dcost);
  return z >= o.c1;
}

// What solve() does with the outcome of the line search (:253-270).  `rollouts` = candidates the
// reference would have evaluated; `residual` belongs to the last of them.  Returns true when that last
// candidate becomes the nominal.
HD bool tick_finish(const IlqrOpts &o, bool accept, real residual, int rollouts, Prob &p) {
  p.n_fwd += rollouts;
  if (residual < o.atol) {  // :253-257 -- taken even if the line search rejected it
    p.status = TFMPC_ST_CONVERGED; p.phase = PH_DONE;
    return true;
  }
  if (accept) {  // :259-266
    p.delta = fmin(1.0 / o.delta_0, p.delta / o.delta_0);
    p.mu = p.mu * p.delta * (double)(p.mu * p.delta > o.mu_min);
    p.guard = 0;
    if (p.iteration + 1 >= o.max_iterations) { p.status = TFMPC_ST_MAXITER; p.phase = PH_DONE; }  // `for iteration in t` exhausted (:227)
    else p.iteration++;
    return true;
  }
  p.delta = fmax(o.delta_0, p.delta * o.delta_0);  // :267-270, then redo the backward pass on the same linearisation
  p.mu = fmax(o.mu_min, p.mu * p.delta);
  if (++p.guard > 200) { p.status = TFMPC_ST_REGLOOP; p.phase = PH_DONE; }
  return false;
}

// Sequential composition (host emulation / documentation of the control flow): traj[0..1] ping-pong.
template <int KIND, int N, int M, int QP = QP_NEWTON, class TJ, class GN>
HD int solve_one(const EnvSmall &e, const IlqrOpts &o, int T, const TJ traj[2], const GN &gain, int32_t *stats) {
  Prob p;
  prob_init(p);
  const CostSink none = {nullptr, 0};
  while (p.phase != PH_DONE) {
    tick_backward<KIND, N, M, QP>(e, o, T, traj[p.cur], gain, p);
    if (p.phase == PH_DONE) break;
    bool accept = false;
    real residual = 0;
    int rollouts = 0;
    for (int ai = 0; ai < N_ALPHA && !accept; ai++) {  // :322 first-accept backtracking
      real J;
      forward_pass<KIND, N, M>(e, T, traj[p.cur], gain, o.alphas[ai], traj[p.cur ^ 1], none, J, residual);
      rollouts++;
      accept = ls_accepts(o, o.alphas[ai], p.J_hat, p.dV1, p.dV2, J);
    }
    if (tick_finish(o, accept, residual, rollouts, p)) p.cur ^= 1;
  }
  stats[0] = p.iteration; stats[1] = p.n_bwd; stats[2] = p.n_fwd; stats[3] = p.status;
  return p.cur;
}
