// Shared definitions for the tfmpc_b200 CUDA sources (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "../../include/tfmpc_b200.h"

typedef tfmpc_real real;

#define MAXD TFMPC_MAX_DIM
#define MAXZ TFMPC_MAX_ZONES
#define N_ALPHA 11  // np.geomspace(1.0, alpha_min, 11), reference ilqr.py:322

#ifdef __CUDACC__
#define HD __host__ __device__ __forceinline__
#else
#define HD inline
#endif

// ---- scalar math in the build's precision.  The fp32 product library is compiled with --use_fast_math (build.py:
// MUFU-based division / exp / sqrt, FTZ), the fp64 verification build is IEEE without FMA contraction, and lib_ieee/ holds an IEEE
// fp32 build of the same sources for the A/B of tests/test_gpu_fullsize.py::test_c3_fast_math_does_not_create_failures
// (status codes and iteration counts of the whole C3 batch, both builds against the oracle); DESIGN.md "numerics" has the data.
HD real r_abs(real v) { return v < 0 ? -v : v; }
HD real r_max(real a, real b) { return a > b ? a : b; }
HD real r_min(real a, real b) { return a < b ? a : b; }
HD real r_sgn(real v) { return (real)((v > 0) - (v < 0)); }
HD real r_clip(real v, real lo, real hi) { return v < lo ? lo : (v > hi ? hi : v); }
// Explicitly fused / explicitly unfused products for the few places where TWO code shapes must round identically in the fp32
// product build (the one-thread Q assembly and its lane-parallel twin in the solo engine): an explicit fma chain cannot be
// re-contracted by the compiler, and __fmul_rn is never fused into a following add.  The fp64 verification build is compiled
// without FMA contraction and keeps separate multiply / add (what the gcc-built oracle does); the host emulation likewise.
HD real r_fma(real a, real b, real c) {
#if defined(__CUDA_ARCH__) && !defined(TFMPC_F64)
  return __fmaf_rn(a, b, c);
#else
  return a * b + c;
#endif
}
HD real r_mul(real a, real b) {
#if defined(__CUDA_ARCH__) && !defined(TFMPC_F64)
  return __fmul_rn(a, b);
#else
  return a * b;
#endif
}
#ifdef TFMPC_F64
HD real r_sqrt(real v) { return sqrt(v); }
HD real r_exp(real v) { return exp(v); }
HD real r_sin(real v) { return sin(v); }
HD real r_cos(real v) { return cos(v); }
#else
HD real r_sqrt(real v) { return sqrtf(v); }
HD real r_exp(real v) { return expf(v); }
HD real r_sin(real v) { return sinf(v); }
HD real r_cos(real v) { return cosf(v); }
#endif

// ---- environment parameters -------------------------------------------------------
// Small environments (thread-per-problem kernels): passed by value as a kernel argument,
// so every field is a constant-bank operand.
#define QP_MAX_STEPS 128
struct EnvSmall {
  int kind, n, m, nz, bounded;
  real goal[4], low[4], high[4], beta;
  real center[MAXZ][2], decay[MAXZ];
  // box-QP backtracking step sizes (real)(0.6^k), k = 0..qp_klast, produced with the reference's own double
  // recurrence `step *= 0.6` (utils/optimization.py:88); qp_klast is the first k with 0.6^k < 1e-22 (:93).
  // Device (or, in the host emulation, host) pointer; used by the warp-cooperative backtracking.
  const real *qp_steps;
  int qp_klast;
};

// fills tab[0..QP_MAX_STEPS) and returns qp_klast
inline int qp_step_table(real *tab) {
  double step = 1.0;
  int klast = -1;
  for (int k = 0; k < QP_MAX_STEPS; k++) {
    tab[k] = (real)step;
    if (klast < 0 && step < 1e-22) klast = k;
    step *= 0.6;
  }
  return klast;
}

// Large environments (warp-per-problem kernels): one device blob, one row per lane.
// Reservoir rows: cap lb ub lowpen highpen sppen rain | D[n][n] | Dt[n][n]
// HVAC rows:      lb ub s(=1/cap) air_max g_out(=adj/R) g_hall t_out t_hall rowsum(A) | A[n][n] | Bt[n][n] (Bt[i][j] = s_j A[j][i])
struct EnvLarge {
  int kind, n, m;
  const real *vec;   // [nvec][32] zero-padded
  const real *matF;  // [32][32] row i = forward-matvec row of lane i, zero-padded
  const real *matB;  // [32][32] row i = backward-matvec row of lane i, zero-padded
};

struct IlqrOpts {
  real atol, c1;
  int max_iterations;
  double mu_min, delta_0;
  real alphas[N_ALPHA];
};

struct tfmpc_env {
  int kind, n, m, nz, bounded, small;
  double low[MAXD], high[MAXD];
  EnvSmall es;
  EnvLarge el;
  real *dblob;       // device storage behind el
  real *dsteps;      // device copy of the box-QP step table (es.qp_steps)
  double goal[MAXD], beta;  // NavigationLQR (host copy; the dense path takes them as kernel arguments)
  int max_row_nnz;   // large envs: max non-zeros per row over the forward and backward coupling matrices
  int device;
  unsigned long long uid;  // unique per created environment (keys the CUDA-graph cache of the tick solve)
  // cached device scratch for the *_host entry points
  void *h_scratch;
  int64_t h_scratch_bytes;
};

// ---- error plumbing (api.cu) --------------------------------------------------------
int tfmpc_set_error(int code, const char *fmt, ...);
void tfmpc_count_launch(int n);
#define CUDA_TRY(expr)                                                                         \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess) return tfmpc_set_error(TFMPC_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); \
  } while (0)
#define LAUNCH_CHECK()                                                                         \
  do {                                                                                         \
    tfmpc_count_launch(1);                                                                     \
    cudaError_t _e = cudaGetLastError();                                                       \
    if (_e != cudaSuccess) return tfmpc_set_error(TFMPC_E_CUDA, "kernel launch: %s", cudaGetErrorString(_e)); \
  } while (0)

// host launchers implemented per translation unit
int small_ilqr_start(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions,
                     real *costs, cudaStream_t s);
int small_ilqr_backward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, double mu, real *K, real *k,
                        real *J, real *dV1, real *dV2, int32_t *status, cudaStream_t s);
int small_ilqr_forward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, const real *K, const real *k,
                       double alpha, real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s);
int64_t small_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T);
int small_ilqr_graph_mode(int on);             // returns the previous mode
void small_ilqr_forget(const tfmpc_env *e);   // drop the cached CUDA graphs of a destroyed environment
// done != nullptr: asynchronous form -- `s` does not wait for the straggler ticks, `done` is recorded behind the results
int small_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s, cudaEvent_t done = nullptr);
// persistent work-queue solve (ilqr_queue.cu); qp = QP_NEWTON / QP_CLOSED (small_core.cuh)
int64_t queue_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T);
int queue_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, int qp, cudaStream_t s);
int queue_ilqr_option(const char *name, int value, int *previous);
int queue_ilqr_counters(const void *ws, int *out, int n, cudaStream_t s);
int64_t queue_ilqr_trace(const tfmpc_env *e, int64_t B, int T, const void *ws, unsigned *out, int64_t max_records, cudaStream_t s);
int small_boxqp(int64_t B, int m, const real *H, const real *q, const real *lo, const real *hi, real *x, real *Hfree, int32_t *isfree,
                int32_t *nfree, int32_t *status, cudaStream_t s);

int warp_ilqr_start(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions,
                    real *costs, cudaStream_t s);
int warp_ilqr_backward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, double mu, real *K, real *k,
                       real *J, real *dV1, real *dV2, int32_t *status, cudaStream_t s);
int warp_ilqr_forward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, const real *K, const real *k,
                      double alpha, real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s);
int64_t warp_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T);
int warp_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                    real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s);

int env_ops_step(const tfmpc_env *e, int64_t R, const real *x, const real *u, real *xn, real *cost, cudaStream_t s);
int env_ops_plant_noise(const tfmpc_env *e, int64_t R, real *xn, unsigned long long seed, unsigned long long offset, cudaStream_t s);
int env_ops_initial_actions(const tfmpc_env *e, int64_t B, int T, unsigned long long seed, real *u_init, cudaStream_t s);
int env_ops_final_cost(const tfmpc_env *e, int64_t R, const real *x, real *cost, cudaStream_t s);
int env_ops_linearize(const tfmpc_env *e, int64_t R, const real *x, const real *u, real *f_x, real *f_u, real *l, real *l_x,
                      real *l_u, real *l_xx, real *l_uu, real *l_ux, real *l_xu, cudaStream_t s);
int env_ops_final_quad(const tfmpc_env *e, int64_t R, const real *x, real *l, real *l_x, real *l_xx, cudaStream_t s);

int lqr_solve_launch(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                     const real *c, int64_t sc, const real *x0, int terminal_zero, real *states, real *actions, real *costs,
                     real *K, real *k, real *V, real *v, real *cst, int32_t *status, cudaStream_t s);
int lqr_forward_launch(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                       const real *c, int64_t sc, const real *K, const real *k, const real *x0, real *states, real *actions, real *costs,
                       cudaStream_t s);
int lqr_step_launch(int64_t R, int n, int m, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC, const real *c,
                    int64_t sc, const real *x, const real *u, real *xn, real *cost, real *fcost, cudaStream_t s);
int dense_backward_launch(int64_t B, int T, int n, int m, int bounded, const double *low, const double *high, const real *actions,
                          const real *f_x, const real *f_u, const real *l, const real *l_x, const real *l_u, const real *l_xx,
                          const real *l_uu, const real *l_xu, const real *fl, const real *fl_x, const real *fl_xx, double mu, real *K,
                          real *k, real *J, real *dV1, real *dV2, int32_t *status, cudaStream_t s);
int64_t dense_navlqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T);
int dense_navlqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                       real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s);
int dense_navlqr_forward(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *xh, const real *uh, const real *K, const real *k,
                         double alpha, real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s);
