// TEST INFRASTRUCTURE ONLY -- executes the __host__ __device__ per-problem code of
// tfmpc_b200/csrc/small_core.cuh on the CPU, one "thread" at a time, with the same
// struct-of-arrays workspace addressing the CUDA kernel uses.  It exists because the build
// container has no GPU: it lets the CPU test-suite exercise the exact device logic (branching,
// box-QP, schedule) before a GPU run.  It is NOT part of the product and is never loaded by
// tfmpc_b200/.
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../tfmpc_b200/csrc/small_core.cuh"

static void fill_env(EnvSmall &s, int kind, int n, int nz, const double *p) {
  memset(&s, 0, sizeof(s));
  s.kind = kind; s.n = n; s.m = n; s.nz = nz;
  if (kind == TFMPC_ENV_NAVLQR) {
    for (int i = 0; i < n; i++) { s.goal[i] = (real)p[i]; s.low[i] = (real)p[n + 1 + i]; s.high[i] = (real)p[n + 1 + n + i]; }
    s.beta = (real)p[n];
  } else {
    for (int i = 0; i < 2; i++) { s.goal[i] = (real)p[i]; s.low[i] = (real)p[2 + i]; s.high[i] = (real)p[4 + i]; }
    for (int z = 0; z < nz; z++) { s.center[z][0] = (real)p[6 + 2 * z]; s.center[z][1] = (real)p[6 + 2 * z + 1]; s.decay[z] = (real)p[6 + 2 * nz + z]; }
  }
  static real tab[QP_MAX_STEPS];
  s.qp_klast = qp_step_table(tab);
  s.qp_steps = tab;
  s.bounded = 1;
  for (int i = 0; i < n; i++) if (std::isinf((double)s.low[i]) || std::isinf((double)s.high[i])) s.bounded = 0;
}

template <int KIND, int N, int M, int QP>
static void run(const EnvSmall &e, const IlqrOpts &o, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions,
                real *costs, int32_t *stats) {
  // same vector-record workspace addressing as the CUDA solve kernels (chunk-major, slot fastest)
  const int64_t S = (B + 31) / 32 * 32;
  const int64_t nomch = (int64_t)(T + 1) * VecTraj<N, M>::CH, gch = (int64_t)T * VecGain<N, M>::CH;
  std::vector<R4> ws((size_t)(2 * nomch + gch) * S);
  const int64_t nx = (int64_t)(T + 1) * N, nu = (int64_t)T * M;
  for (int64_t b = 0; b < B; b++) {
    const int64_t chn = VecTraj<N, M>::CH, chg = VecGain<N, M>::CH;
    VecTraj<N, M> traj[2] = {{ws.data() + b, chn * S, S}, {ws.data() + nomch * S + b, chn * S, S}};   // [t][chunk][slot]
    VecGain<N, M> gain = {ws.data() + 2 * nomch * S + b * chg, S * chg, 1};                           // [t][slot][chunk]
    const CostSink none = {nullptr, 0};
    start_pass<KIND, N, M>(e, T, x0 + b * N, u_init + b * nu, traj[0], none);
    int cur = solve_one<KIND, N, M, QP>(e, o, T, traj, gain, stats + b * 4);
    real x[N], u[M];
    for (int t = 0; t < T; t++) {
      traj[cur].load_xu(t, x, u);
      for (int i = 0; i < N; i++) states[b * nx + t * N + i] = x[i];
      for (int i = 0; i < M; i++) actions[b * nu + t * M + i] = u[i];
      costs[b * (T + 1) + t] = env_cost<KIND, N, M>(e, x, u);
    }
    traj[cur].load_x(T, x);
    for (int i = 0; i < N; i++) states[b * nx + T * N + i] = x[i];
    costs[b * (T + 1) + T] = env_final_cost<KIND, N, M>(e, x);
  }
}

// qp_mode: 0 = the reference's projected-Newton box-QP (QP_NEWTON), 2 = closed form for m <= 2 (QP_CLOSED)
extern "C" int emul_ilqr_solve_qp(int kind, int n, int nz, const double *params, double atol, int max_iterations, double mu_min, double delta_0,
                                  double c1, const double *alphas, int64_t B, int T, const real *x0, const real *u_init, real *states,
                                  real *actions, real *costs, int32_t *stats, int qp_mode) {
  EnvSmall e;
  fill_env(e, kind, n, nz, params);
  IlqrOpts o;
  o.atol = (real)atol; o.c1 = (real)c1; o.max_iterations = max_iterations; o.mu_min = mu_min; o.delta_0 = delta_0;
  for (int i = 0; i < N_ALPHA; i++) o.alphas[i] = (real)alphas[i];
  const bool closed = qp_mode == QP_CLOSED;
  if (kind == TFMPC_ENV_NAVIGATION && n == 2) {
    if (closed) run<TFMPC_ENV_NAVIGATION, 2, 2, QP_CLOSED>(e, o, B, T, x0, u_init, states, actions, costs, stats);
    else run<TFMPC_ENV_NAVIGATION, 2, 2, QP_NEWTON>(e, o, B, T, x0, u_init, states, actions, costs, stats);
  } else if (kind == TFMPC_ENV_NAVLQR && n == 2) {
    if (closed) run<TFMPC_ENV_NAVLQR, 2, 2, QP_CLOSED>(e, o, B, T, x0, u_init, states, actions, costs, stats);
    else run<TFMPC_ENV_NAVLQR, 2, 2, QP_NEWTON>(e, o, B, T, x0, u_init, states, actions, costs, stats);
  } else if (kind == TFMPC_ENV_NAVLQR && n == 3) run<TFMPC_ENV_NAVLQR, 3, 3, QP_NEWTON>(e, o, B, T, x0, u_init, states, actions, costs, stats);
  else return -2;
  return 0;
}
extern "C" int emul_ilqr_solve(int kind, int n, int nz, const double *params, double atol, int max_iterations, double mu_min, double delta_0,
                               double c1, const double *alphas, int64_t B, int T, const real *x0, const real *u_init, real *states,
                               real *actions, real *costs, int32_t *stats) {
  return emul_ilqr_solve_qp(kind, n, nz, params, atol, max_iterations, mu_min, delta_0, c1, alphas, B, T, x0, u_init, states, actions, costs, stats, 0);
}
extern "C" int emul_real_bytes(void) { return (int)sizeof(real); }
