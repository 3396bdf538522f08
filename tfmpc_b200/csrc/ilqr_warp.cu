// Lane-per-state iLQR kernels for the LARGE environments (Reservoir, HVAC; n = m <= 32).
//
// Mapping: one problem per group of G lanes (G = 4, 8, 16 or 32, the padded state dimension;
// a warp holds 32/G problems), lane i owns state/action component i.  The reference layouts
// states[b][t][i] / actions[b][t][i] are already lane-contiguous, so the nominal trajectory is
// read and written in place with coalesced accesses and no staging copy.
//
// Algorithmic specialisation (what the reference itself computes for these two envs): their
// costs are piecewise linear, so l_xx = l_uu = l_xu = 0 and the terminal V_xx = 0
// (reference tests/test_env_reservoir.py:223-233, tests/test_env_hvac.py:207-211).  Then in
// iLQR.backward (tfmpc/solvers/ilqr.py:119-167) Q_xx = Q_uu = Q_ux = 0, count_nonzero(V_xx) == 0
// at every step, so the bounded branch :139-141 fires 100% of the time: K = 0,
// k = where(Q_u >= 0, low - u, high - u), V_x <- Q_x, V_xx <- 0, dV2 = 0.  The dense products
// with zero matrices are skipped; every value the reference would produce is produced.
// (SURVEY.md section 0 finding 3 and Appendix C, Q9.)  mu therefore has no effect here.
#include "common.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int kWarpsPerBlock = 4;

template <int G>
__device__ __forceinline__ real group_sum(real v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o, G);
  return v;
}
template <int G>
__device__ __forceinline__ real group_max(real v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v = r_max(v, __shfl_xor_sync(FULL, v, o, G));
  return v;
}
// One row of a coupling matrix, held by the lane that owns that row.  NZ == 0: dense (G registers, G shuffles per
// matvec).  NZ > 0: the row has at most NZ non-zeros (reservoir chains: 1, room grids: 4) and is kept as (column,
// value) pairs in ascending column order -- NZ shuffles per matvec.  Skipping structural zeros is exact and the
// remaining terms are accumulated in the same (ascending) order as the dense loop.
template <int G, int NZ>
struct Row {
  real val[NZ > 0 ? NZ : G];
  int col[NZ > 0 ? NZ : 1];

  __device__ void load(const real *__restrict__ src) {  // src = 32 padded entries of this lane's row
    if (NZ == 0) {
#pragma unroll
      for (int j = 0; j < G; j++) val[j] = src[j];
    } else {
      int c = 0;
#pragma unroll
      for (int q = 0; q < NZ; q++) { val[q] = 0; col[q] = 0; }
      for (int j = 0; j < G; j++) {
        real v = src[j];
        if (v != (real)0 && c < NZ) {
#pragma unroll
          for (int q = 0; q < NZ; q++) if (q == c) { val[q] = v; col[q] = j; }
          c++;
        }
      }
    }
  }
  // sum_j row[j] * v_j
  __device__ __forceinline__ real dot(real v) const {
    real acc = 0;
    if (NZ == 0) {
#pragma unroll
      for (int j = 0; j < G; j++) acc += val[j] * __shfl_sync(FULL, v, j, G);
    } else {
#pragma unroll
      for (int q = 0; q < NZ; q++) acc += val[q] * __shfl_sync(FULL, v, col[q], G);
    }
    return acc;
  }
  // sum_j -row[j] * (x_i - x_j)
  __device__ __forceinline__ real diffuse(real x) const {
    real acc = 0;
    if (NZ == 0) {
#pragma unroll
      for (int j = 0; j < G; j++) acc += -val[j] * (x - __shfl_sync(FULL, x, j, G));
    } else {
#pragma unroll
      for (int q = 0; q < NZ; q++) acc += -val[q] * (x - __shfl_sync(FULL, x, col[q], G));
    }
    return acc;
  }
};

// Per-lane environment: parameters of component i in registers.
template <int KIND, int G, int NZ = 0>
struct LaneEnv {
  real p[9];
  Row<G, NZ> rowF, rowB;
  bool in_range;  // lane index < n

  __device__ void load(const EnvLarge &e, int i) {
    in_range = i < e.n;
#pragma unroll
    for (int r = 0; r < 9; r++) p[r] = e.vec[r * 32 + i];
    rowF.load(e.matF + i * 32);
    rowB.load(e.matB + i * 32);
  }

  // per-lane term of cost(state, action) / final_cost(state); the caller sums over the group
  __device__ __forceinline__ real cost_term(bool act, real x, real u, bool final) const {
    if (!act) return (real)0;
    if (KIND == TFMPC_ENV_RESERVOIR) {  // reference reservoir/__init__.py:64-84; p = cap lb ub lowpen highpen sppen rain
      real c1 = -p[3] * r_max((real)0, p[1] - x);
      real c2 = -p[4] * r_max((real)0, x - p[2]);
      real c3 = -p[5] * r_abs((p[1] + p[2]) / (real)2.0 - x);
      return c1 + c2 + c3;
    } else {  // reference hvac/__init__.py:93-126; p = lb ub s air_max g_out g_hall t_out t_hall rowsum
      real oob = (real)20000 * (r_max((real)0, p[0] - x) + r_max((real)0, x - p[1]));
      real sp = (real)10.0 * r_abs((p[0] + p[1]) / (real)2 - x);
      return final ? oob + sp : (real)1.0 * (u * p[3]) + oob + sp;
    }
  }

  // transition(state, action), deterministic (cec=True)
  __device__ __forceinline__ real step(bool act, real x, real u) const {
    real xn;
    if (KIND == TFMPC_ENV_RESERVOIR) {  // reservoir/__init__.py:47-62
      real out = u * x;
      real inflow = rowF.dot(out);
      real cap = act ? p[0] : (real)1;
      real vap = (real)0.5 * r_sin(x / cap) * x;
      xn = x + p[6] + inflow - vap - out;
    } else {  // hvac/__init__.py:69-91,128-149
      real air = u * p[3];
      real heating = air * (real)1.006 * ((real)40.0 - x);
      real cbr = rowF.diffuse(x);
      real cwo = p[4] * (p[6] - x);
      real cwh = p[5] * (p[7] - x);
      xn = x + p[2] * (heating + cbr + cwo + cwh);
    }
    return act ? xn : (real)0;
  }

  // l_x component (cost and final cost share it for both envs up to the action term)
  __device__ __forceinline__ real l_x(bool act, real x) const {
    if (!act) return (real)0;
    if (KIND == TFMPC_ENV_RESERVOIR) {  // tests/test_env_reservoir.py:189-221
      real mid = (p[1] + p[2]) / (real)2.0;
      return p[3] * (real)(p[1] - x > 0) - p[4] * (real)(x - p[2] > 0) + p[5] * r_sgn(mid - x);
    } else {  // tests/test_env_hvac.py:170-205
      real mid = (p[0] + p[1]) / (real)2;
      return (real)20000 * ((real)(x - p[1] > 0) - (real)(p[0] - x > 0)) - (real)10.0 * r_sgn(mid - x);
    }
  }

  // Q_x = l_x + f_x^T V_x, Q_u = l_u + f_u^T V_x for this lane (ilqr.py:122-123), analytic f_x, f_u
  __device__ __forceinline__ void adjoint(bool act, real x, real u, real V, real lx, real &Q_x, real &Q_u) const {
    real BV = rowB.dot(V);
    if (KIND == TFMPC_ENV_RESERVOIR) {  // f_x = I - diag(dvap) - diag(u) + D^T diag(u); f_u = -diag(x) + D^T diag(x)
      real cap = act ? p[0] : (real)1;
      real a = x / cap;
      real dvap = (real)0.5 * (r_cos(a) * a + r_sin(a));
      Q_x = lx + (u * BV + ((real)1 - dvap - u) * V);
      Q_u = x * (BV - V);  // factored so that an exact tie V_i == (D V)_i gives exactly 0
    } else {  // f_x = I + diag(s)(A - diag(u amax c_air + A1 + g_out + g_hall)); f_u = diag(s amax c_air (40 - x))
      real diag = (real)1 + p[2] * (-(u * p[3]) * (real)1.006 - p[8] - p[4] - p[5]);
      Q_x = lx + (BV + diag * V);
      Q_u = p[3] + (p[2] * (p[3] * (real)1.006 * ((real)40.0 - x))) * V;
    }
    if (!act) { Q_x = 0; Q_u = 0; }
  }
};

// ---- one problem, one group --------------------------------------------------------
// backward sweep (K == 0): writes k[t][i]; returns J, dV1 (dV2 == 0) and sum_t max_i |k|/(|u|+1)
template <int KIND, int G, int NZ>
__device__ __forceinline__ void group_backward(const LaneEnv<KIND, G, NZ> &E, bool act, int n, int T, int i, const real *__restrict__ X,
                                               const real *__restrict__ U, real *__restrict__ kout, real lo, real hi, real &J,
                                               real &dV1, real &gsum) {
  real xT = act ? X[(int64_t)T * n + i] : (real)0;
  J = group_sum<G>(E.cost_term(act, xT, (real)0, true));  // final l (ilqr.py:104)
  real V = E.l_x(act, xT);                                 // V_x = l_x^f (:101)
  real d1 = 0;
  gsum = 0;
  real xn = act ? X[(int64_t)(T - 1) * n + i] : (real)0, un = act ? U[(int64_t)(T - 1) * n + i] : (real)0;
  for (int t = T - 1; t >= 0; t--) {
    real x = xn, u = un;
    if (t > 0) { xn = act ? X[(int64_t)(t - 1) * n + i] : (real)0; un = act ? U[(int64_t)(t - 1) * n + i] : (real)0; }
    real Q_x, Q_u;
    E.adjoint(act, x, u, V, E.l_x(act, x), Q_x, Q_u);
    real k = (Q_u >= 0) ? lo - u : hi - u;  // ilqr.py:141
    if (!act) k = 0;
    V = Q_x;                                // :149-154 with K = 0, Q_ux = 0, Q_uu = 0
    J += group_sum<G>(E.cost_term(act, x, u, false));  // :164
    d1 += k * Q_u;                          // :166 (summed over lanes at the end)
    gsum += group_max<G>(act ? r_abs(k) / (r_abs(u) + (real)1.0) : (real)0);
    if (act) kout[(int64_t)t * n + i] = k;
  }
  dV1 = group_sum<G>(d1);
}

// forward rollout with K == 0 (ilqr.py:174-212): u = clip(u_hat + alpha k)
template <int KIND, int G, int NZ, bool WRITE_C>
__device__ __forceinline__ void group_forward(const LaneEnv<KIND, G, NZ> &E, bool act, int n, int T, int i, const real *__restrict__ Xh,
                                              const real *__restrict__ Uh, const real *__restrict__ kin, real alpha, real lo, real hi,
                                              real *__restrict__ Xo, real *__restrict__ Uo, real *__restrict__ Co, real &J, real &residual) {
  real x = act ? Xh[i] : (real)0;
  if (act) Xo[i] = x;
  real res = 0;
  J = 0;
  real uh = act ? Uh[i] : (real)0, kk = act ? kin[i] : (real)0;
  for (int t = 0; t < T; t++) {
    real du = alpha * kk;                       // :194 (K = 0)
    real u = r_clip(uh + du, lo, hi);           // :196-197
    res = r_max(res, r_abs(du));                // :206
    if (t + 1 < T) { uh = act ? Uh[(int64_t)(t + 1) * n + i] : (real)0; kk = act ? kin[(int64_t)(t + 1) * n + i] : (real)0; }
    real c = group_sum<G>(E.cost_term(act, x, u, false));
    real xn = E.step(act, x, u);
    if (act) { Uo[(int64_t)t * n + i] = u; Xo[(int64_t)(t + 1) * n + i] = xn; }
    if (WRITE_C && act && (i == 0)) Co[t] = c;
    J += c;
    x = xn;
  }
  real cf = group_sum<G>(E.cost_term(act, x, (real)0, true));
  if (WRITE_C && act && (i == 0)) Co[T] = cf;
  J += cf;
  residual = group_max<G>(res);
}

template <int KIND, int G, int NZ, bool WRITE_C>
__device__ __forceinline__ void group_start(const LaneEnv<KIND, G, NZ> &E, bool act, int n, int T, int i, const real *__restrict__ x0,
                                            const real *__restrict__ Ui, real *__restrict__ Xo, real *__restrict__ Uo, real *__restrict__ Co) {
  real x = act ? x0[i] : (real)0;
  if (act) Xo[i] = x;
  for (int t = 0; t < T; t++) {
    real u = act ? Ui[(int64_t)t * n + i] : (real)0;
    if (WRITE_C) { real c = group_sum<G>(E.cost_term(act, x, u, false)); if (act && i == 0) Co[t] = c; }
    real xn = E.step(act, x, u);
    if (act) { Uo[(int64_t)t * n + i] = u; Xo[(int64_t)(t + 1) * n + i] = xn; }
    x = xn;
  }
  if (WRITE_C) { real cf = group_sum<G>(E.cost_term(act, x, (real)0, true)); if (act && i == 0) Co[T] = cf; }
}

// ---- stage kernels (one group per problem, grid-stride) -----------------------------------
template <int KIND, int G, int NZ>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) kw_start(EnvLarge e, int64_t B, int T, const real *__restrict__ x0,
                                                                const real *__restrict__ u_init, real *__restrict__ states,
                                                                real *__restrict__ actions, real *__restrict__ costs) {
  const int i = threadIdx.x % G, n = e.n;
  LaneEnv<KIND, G, NZ> E;
  E.load(e, i);
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ngroups = (int64_t)gridDim.x * blockDim.x / G;
  // whole warps iterate together so that the group shuffles stay convergent
  int64_t warp_first = gid - (threadIdx.x % 32) / G;
  for (int64_t base = warp_first; base < B; base += ngroups) {
    int64_t b = base + (threadIdx.x % 32) / G;
    bool valid = b < B;
    int64_t bb = valid ? b : B - 1;
    real *Xo = states + bb * (T + 1) * n, *Uo = actions + bb * T * n, *Co = costs + bb * (T + 1);
    group_start<KIND, G, NZ, true>(E, E.in_range && valid, n, T, i, x0 + bb * n, u_init + bb * T * n, Xo, Uo, Co);
  }
}

template <int KIND, int G, int NZ>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) kw_backward(EnvLarge e, int64_t B, int T, const real *__restrict__ states,
                                                                   const real *__restrict__ actions, real *__restrict__ K,
                                                                   real *__restrict__ k, real *__restrict__ J, real *__restrict__ dV1,
                                                                   real *__restrict__ dV2, int32_t *__restrict__ status, real lo, real hi) {
  const int i = threadIdx.x % G, n = e.n;
  LaneEnv<KIND, G, NZ> E;
  E.load(e, i);
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t warp_first = gid - (threadIdx.x % 32) / G;
  for (int64_t base = warp_first; base < B; base += ngroups) {
    int64_t b = base + (threadIdx.x % 32) / G;
    bool valid = b < B;
    int64_t bb = valid ? b : B - 1;
    const bool act = E.in_range && valid;
    real Jb, d1, g;
    group_backward<KIND, G, NZ>(E, act, n, T, i, states + bb * (T + 1) * n, actions + bb * T * n, k + bb * T * n, lo, hi, Jb, d1, g);
    if (act) {
      for (int t = 0; t < T; t++)
        for (int j = 0; j < n; j++) K[((bb * T + t) * n + i) * n + j] = 0;
    }
    if (valid && i == 0) { J[b] = Jb; dV1[b] = d1; dV2[b] = 0; if (status) status[b] = 0; }
  }
}

template <int KIND, int G, int NZ>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) kw_forward(EnvLarge e, int64_t B, int T, const real *__restrict__ states,
                                                                  const real *__restrict__ actions, const real *__restrict__ k, real alpha,
                                                                  real *__restrict__ xs, real *__restrict__ us, real *__restrict__ cs,
                                                                  real *__restrict__ J, real *__restrict__ residual, real lo, real hi) {
  const int i = threadIdx.x % G, n = e.n;
  LaneEnv<KIND, G, NZ> E;
  E.load(e, i);
  int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ngroups = (int64_t)gridDim.x * blockDim.x / G;
  int64_t warp_first = gid - (threadIdx.x % 32) / G;
  for (int64_t base = warp_first; base < B; base += ngroups) {
    int64_t b = base + (threadIdx.x % 32) / G;
    bool valid = b < B;
    int64_t bb = valid ? b : B - 1;
    real Jb, res;
    group_forward<KIND, G, NZ, true>(E, E.in_range && valid, n, T, i, states + bb * (T + 1) * n, actions + bb * T * n, k + bb * T * n, alpha,
                                 lo, hi, xs + bb * (T + 1) * n, us + bb * T * n, cs + bb * (T + 1), Jb, res);
    if (valid && i == 0) { J[b] = Jb; residual[b] = res; }
  }
}

// ---- fused solve: persistent kernel, work queue ---------------------------------------------
// Each warp repeatedly claims 32/G problems from an atomic counter and runs the whole
// iLQR.solve for them: schedule, convergence tests and first-accept line search on the device.
// When a warp holds several problems (G < 32) their control flow is made warp-uniform by
// iterating until every group in the warp is done (finished groups idle with active = false).
template <int KIND, int G, int NZ>
__global__ void __launch_bounds__(32 * kWarpsPerBlock) kw_solve(EnvLarge e, IlqrOpts o, int64_t B, int T, const real *__restrict__ x0,
                                                                const real *__restrict__ u_init, real *__restrict__ states,
                                                                real *__restrict__ actions, real *__restrict__ costs,
                                                                int32_t *__restrict__ stats, unsigned long long *__restrict__ counter,
                                                                real *__restrict__ ws, real lo, real hi) {
  constexpr int PPW = 32 / G;  // problems per warp
  const int lane = threadIdx.x % 32, i = lane % G, sub = lane / G, n = e.n;
  LaneEnv<KIND, G, NZ> E;
  E.load(e, i);
  const int64_t per = (int64_t)(3 * T + 1) * n;  // workspace reals per problem: cand X, cand U, k
  for (;;) {
    unsigned long long first = 0;
    if (lane == 0) first = atomicAdd(counter, (unsigned long long)PPW);
    first = __shfl_sync(FULL, first, 0);
    if ((int64_t)first >= B) break;
    const int64_t b = (int64_t)first + sub;
    const bool valid = b < B;
    const int64_t bb = valid ? b : B - 1;
    const bool live = E.in_range && valid;  // this lane owns a component of a real problem
    // buffer pair 0 = the output arrays, pair 1 = workspace
    real *Xb[2] = {states + bb * (T + 1) * n, ws + bb * per};
    real *Ub[2] = {actions + bb * T * n, ws + bb * per + (int64_t)(T + 1) * n};
    real *kb = ws + bb * per + (int64_t)(2 * T + 1) * n;
    group_start<KIND, G, NZ, false>(E, live, n, T, i, x0 + bb * n, u_init + bb * T * n, Xb[0], Ub[0], nullptr);
    double mu = 0.0, delta = 1.0;  // kept for fidelity with ilqr.py:215-216,261-270; mu is inert when V_xx == 0
    int cur = 0, n_bwd = 0, n_fwd = 0, status = TFMPC_ST_MAXITER, iteration = 0;
    bool done = !valid;
    // `done` differs between the groups of a warp; the loops below run while ANY group is live
    for (int it = 0; it < o.max_iterations; it++) {
      if (__all_sync(FULL, done)) break;
      if (!done) iteration = it;
      int guard = 0;
      bool iter_open = !done;  // this group still has to finish outer iteration `it`
      while (__any_sync(FULL, iter_open)) {
        real J_hat, dV1, g;
        group_backward<KIND, G, NZ>(E, live && iter_open, n, T, i, Xb[cur], Ub[cur], kb, lo, hi, J_hat, dV1, g);
        if (iter_open) n_bwd++;
        g = g / (real)T;  // ilqr.py:243
        bool stop = false;
        if (iter_open) {
          if (!(g == g)) { status = TFMPC_ST_NAN; stop = true; }
          else if (g < o.atol) { status = TFMPC_ST_CONVERGED; stop = true; }  // :245-248
        }
        bool searching = iter_open && !stop;
        bool accept = false;
        real residual = 0;
        for (int ai = 0; ai < N_ALPHA; ai++) {  // :322 first-accept backtracking
          if (!__any_sync(FULL, searching)) break;
          real alpha = o.alphas[ai], J, res;
          group_forward<KIND, G, NZ, false>(E, live && searching, n, T, i, Xb[cur], Ub[cur], kb, alpha, lo, hi, Xb[cur ^ 1], Ub[cur ^ 1],
                                        nullptr, J, res);
          if (searching) {
            n_fwd++;
            residual = res;
            real delta_J = -alpha * (dV1 + alpha * (real)0);  // :339, dV2 == 0
            real dcost = J_hat - J;
            real z = (delta_J > 0) ? dcost / delta_J : r_sgn(dcost);
            if (z >= o.c1) { accept = true; searching = false; }  // :351
          }
        }
        if (iter_open && !stop) {
          if (residual < o.atol) {  // :253-257
            status = TFMPC_ST_CONVERGED; stop = true; cur ^= 1;
          } else if (accept) {  // :259-266
            delta = fmin(1.0 / o.delta_0, delta / o.delta_0);
            mu = mu * delta * (double)(mu * delta > o.mu_min);
            cur ^= 1;
            iter_open = false;
          } else {  // :267-270
            delta = fmax(o.delta_0, delta * o.delta_0);
            mu = fmax(o.mu_min, mu * delta);
            if (++guard > 200) { status = TFMPC_ST_REGLOOP; stop = true; }
          }
        }
        if (stop) { done = true; iter_open = false; }
      }
    }
    // results: nominal -> output arrays (copy if it ended in the workspace pair), costs recomputed
    if (live && cur == 1) {
      for (int t = 0; t <= T; t++) Xb[0][(int64_t)t * n + i] = Xb[1][(int64_t)t * n + i];
      for (int t = 0; t < T; t++) Ub[0][(int64_t)t * n + i] = Ub[1][(int64_t)t * n + i];
    }
    __syncwarp();
    {
      real *Co = costs + bb * (T + 1);
      for (int t = 0; t <= T; t++) {
        real x = live ? Xb[cur][(int64_t)t * n + i] : (real)0;
        real u = (live && t < T) ? Ub[cur][(int64_t)t * n + i] : (real)0;
        real c = group_sum<G>(E.cost_term(live, x, u, t == T));
        if (valid && i == 0) Co[t] = c;
      }
    }
    if (valid && i == 0) { stats[b * 4] = iteration; stats[b * 4 + 1] = n_bwd; stats[b * 4 + 2] = n_fwd; stats[b * 4 + 3] = status; }
  }
}

int pick_group(int n) { return n <= 4 ? 4 : (n <= 8 ? 8 : (n <= 16 ? 16 : 32)); }

int check_large(const tfmpc_env *e) {
  if (e->kind != TFMPC_ENV_RESERVOIR && e->kind != TFMPC_ENV_HVAC)
    return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "no lane-per-state kernel for environment kind %d with n=%d (dense path not built yet)", e->kind, e->n);
  return TFMPC_OK;
}

int sm_count(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

}  // namespace

#define WARP_DISPATCH_G(e, KD, NZ, CALL)                                \
  do {                                                                   \
    int g_ = pick_group((e)->n);                                         \
    if (g_ == 4) { CALL(KD, 4, NZ); } else if (g_ == 8) { CALL(KD, 8, NZ); } \
    else if (g_ == 16) { CALL(KD, 16, NZ); } else { CALL(KD, 32, NZ); }  \
  } while (0)
// sparse rows (<= 4 non-zeros per row of both coupling matrices: reservoir chains, room grids) or dense rows
#define WARP_DISPATCH(e, CALL)                                           \
  do {                                                                   \
    const bool sp_ = (e)->max_row_nnz <= 4;                              \
    if ((e)->kind == TFMPC_ENV_RESERVOIR) {                              \
      if (sp_) WARP_DISPATCH_G(e, TFMPC_ENV_RESERVOIR, 4, CALL); else WARP_DISPATCH_G(e, TFMPC_ENV_RESERVOIR, 0, CALL); \
    } else {                                                             \
      if (sp_) WARP_DISPATCH_G(e, TFMPC_ENV_HVAC, 4, CALL); else WARP_DISPATCH_G(e, TFMPC_ENV_HVAC, 0, CALL); \
    }                                                                    \
  } while (0)

static unsigned stage_grid(const tfmpc_env *e, int64_t B) {
  int g = pick_group(e->n);
  int64_t groups_per_block = 32 * kWarpsPerBlock / g;
  int64_t blocks = (B + groups_per_block - 1) / groups_per_block;
  int64_t cap = (int64_t)sm_count(e->device) * 8;
  return (unsigned)(blocks < cap ? blocks : cap);
}

int warp_ilqr_start(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions, real *costs,
                    cudaStream_t s) {
  int rc = check_large(e);
  if (rc) return rc;
#define CALL(K, G, NZ) kw_start<K, G, NZ><<<stage_grid(e, B), 32 * kWarpsPerBlock, 0, s>>>(e->el, B, T, x0, u_init, states, actions, costs)
  WARP_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int warp_ilqr_backward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, double mu, real *K, real *k, real *J,
                       real *dV1, real *dV2, int32_t *status, cudaStream_t s) {
  int rc = check_large(e);
  if (rc) return rc;
  (void)mu;  // inert: V_xx == 0 for these environments (see the header comment)
  real lo = (real)e->low[0], hi = (real)e->high[0];
#define CALL(KD, G, NZ) kw_backward<KD, G, NZ><<<stage_grid(e, B), 32 * kWarpsPerBlock, 0, s>>>(e->el, B, T, states, actions, K, k, J, dV1, dV2, status, lo, hi)
  WARP_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int warp_ilqr_forward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, const real *K, const real *k, double alpha,
                      real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s) {
  int rc = check_large(e);
  if (rc) return rc;
  (void)K;  // K == 0 for these environments
  real lo = (real)e->low[0], hi = (real)e->high[0];
#define CALL(KD, G, NZ) kw_forward<KD, G, NZ><<<stage_grid(e, B), 32 * kWarpsPerBlock, 0, s>>>(e->el, B, T, states, actions, k, (real)alpha, xs, us, cs, J, residual, lo, hi)
  WARP_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int64_t warp_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T) {
  return 256 + (int64_t)(3 * T + 1) * e->n * B * (int64_t)sizeof(real);
}

int warp_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states, real *actions,
                    real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s) {
  int rc = check_large(e);
  if (rc) return rc;
  if (ws_bytes < warp_ilqr_workspace_bytes(e, B, T)) return tfmpc_set_error(TFMPC_E_WORKSPACE, "workspace too small");
  unsigned long long *counter = (unsigned long long *)ws;
  real *wsr = (real *)((char *)ws + 256);
  CUDA_TRY(cudaMemsetAsync(counter, 0, 256, s));
  real lo = (real)e->low[0], hi = (real)e->high[0];
  int g = pick_group(e->n);
  int64_t warps_needed = (B + (32 / g) - 1) / (32 / g);
  int64_t blocks_needed = (warps_needed + kWarpsPerBlock - 1) / kWarpsPerBlock;
  int64_t resident = (int64_t)sm_count(e->device) * 4;  // persistent: 4 blocks x 4 warps per SM
  unsigned grid = (unsigned)(blocks_needed < resident ? blocks_needed : resident);
#define CALL(KD, G, NZ) kw_solve<KD, G, NZ><<<grid, 32 * kWarpsPerBlock, 0, s>>>(e->el, o, B, T, x0, u_init, states, actions, costs, stats, counter, wsr, lo, hi)
  WARP_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}
