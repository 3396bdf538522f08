// TEST INFRASTRUCTURE ONLY -- executes the persistent work-queue solver of tfmpc_b200/csrc/queue_core.cuh on the CPU:
// every lane of every emulated warp is an OS thread, every warp collective a barrier (warp_rt.cuh, host branch).  What
// this exercises without a GPU: the ticket queue (acquire / re-queue / retire), the line-search rounds (lane -> (problem,
// step size) assignment, first-accept selection, in-place candidate store and replay), the cooperative 128-byte line
// staging, and the per-problem results, which must equal the sequential composition solve_one() bit for bit.
// Never loaded by tfmpc_b200/.
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "../../tfmpc_b200/csrc/queue_core.cuh"

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
static int run(const EnvSmall &e, const IlqrOpts &o, int B, int T, const real *x0, const real *u_init, real *states, real *actions,
               real *costs, int32_t *stats, int nwarps, int w_target, int patience, int solo_max, int *ctrl_out) {
  using namespace tq;
  constexpr int CHn = VecTraj<N, M>::CH;
  const int NH = ((T + 1) * CHn + CPH - 1) / CPH, row_r4 = NH * CPH;
  unsigned cap = 1024;
  while (cap < 2u * (unsigned)B) cap <<= 1;
  std::vector<int> ctrl(C_INTS, 0);
  std::vector<unsigned long long> ring(cap, 0ull);
  std::vector<QProb> prob(B);
  std::vector<R4> traj((size_t)2 * B * row_r4);
  std::vector<R2> gain((size_t)nwarps * T * Gain2<N, M>::CH2 * 32);
  memset(traj.data(), 0xff, traj.size() * sizeof(R4));   // NaN pattern: any read of an unwritten record shows up
  ctrl[C_TAIL] = B; ctrl[C_COUNT] = B; ctrl[C_ALIVE] = nwarps;
  QParams q;
  q.ctrl = ctrl.data(); q.ring = ring.data(); q.ring_mask = cap - 1; q.prob = prob.data(); q.traj = traj.data(); q.gain = gain.data();
  q.B = B; q.T = T; q.row_r4 = row_r4; q.w_target = w_target; q.patience = patience; q.solo_max = solo_max; q.w_solo = solo_max > 0 ? w_target : 0; q.watchdog_ns = 120ull * 1000000000ull;
  q.trace = nullptr; q.trace_cap = 0;
  static int bulk_counter = 0;
  q.bulk = solo_max == 0 && w_target > 1 ? &bulk_counter : nullptr; q.bulk_thr = 4 * w_target;   // (solo_max = 0: the draining-pipeline rule decides)
  q.x0 = x0; q.u_init = u_init; q.states = states; q.actions = actions; q.costs = costs; q.stats = stats;
  std::vector<WarpShared> shared(nwarps);
  std::vector<WarpSmem<N, M>> smem(nwarps);
  for (auto &w : shared) pthread_barrier_init(&w.bar, nullptr, 32);
  std::vector<std::thread> th;
  for (int w = 0; w < nwarps; w++)
    for (int l = 0; l < 32; l++)
      th.emplace_back([&, w, l] {
        WarpRT rt(l, &shared[w]);
        queue_warp_main<KIND, N, M, QP>(rt, e, o, q, smem[w], w);
      });
  for (auto &t : th) t.join();
  for (auto &w : shared) pthread_barrier_destroy(&w.bar);
  if (ctrl_out) memcpy(ctrl_out, ctrl.data(), sizeof(int) * C_INTS);
  return ctrl[C_ERR];
}

extern "C" int emul_queue_solve(int kind, int n, int nz, const double *params, double atol, int max_iterations, double mu_min, double delta_0,
                                double c1, const double *alphas, int B, int T, const real *x0, const real *u_init, real *states,
                                real *actions, real *costs, int32_t *stats, int qp_mode, int nwarps, int w_target, int patience, int solo_max, int *ctrl_out) {
  EnvSmall e;
  fill_env(e, kind, n, nz, params);
  IlqrOpts o;
  o.atol = (real)atol; o.c1 = (real)c1; o.max_iterations = max_iterations; o.mu_min = mu_min; o.delta_0 = delta_0;
  for (int i = 0; i < N_ALPHA; i++) o.alphas[i] = (real)alphas[i];
  const bool closed = qp_mode == QP_CLOSED;
#define RUN(K, NN, QP) return run<K, NN, NN, QP>(e, o, B, T, x0, u_init, states, actions, costs, stats, nwarps, w_target, patience, solo_max, ctrl_out)
  if (kind == TFMPC_ENV_NAVIGATION && n == 2 && nz <= 2) { if (closed) RUN(TFMPC_ENV_NAVIGATION_Z2, 2, QP_CLOSED); else RUN(TFMPC_ENV_NAVIGATION_Z2, 2, QP_NEWTON); }   // same dispatch as ilqr_queue.cu
  if (kind == TFMPC_ENV_NAVIGATION && n == 2) { if (closed) RUN(TFMPC_ENV_NAVIGATION, 2, QP_CLOSED); else RUN(TFMPC_ENV_NAVIGATION, 2, QP_NEWTON); }
  if (kind == TFMPC_ENV_NAVLQR && n == 1) { if (closed) RUN(TFMPC_ENV_NAVLQR, 1, QP_CLOSED); else RUN(TFMPC_ENV_NAVLQR, 1, QP_NEWTON); }
  if (kind == TFMPC_ENV_NAVLQR && n == 2) { if (closed) RUN(TFMPC_ENV_NAVLQR, 2, QP_CLOSED); else RUN(TFMPC_ENV_NAVLQR, 2, QP_NEWTON); }
  if (kind == TFMPC_ENV_NAVLQR && n == 3) RUN(TFMPC_ENV_NAVLQR, 3, QP_NEWTON);
  if (kind == TFMPC_ENV_NAVLQR && n == 4) RUN(TFMPC_ENV_NAVLQR, 4, QP_NEWTON);
#undef RUN
  return -2;
}
extern "C" int emul_queue_ctrl_ints(void) { return tq::C_INTS; }
