// Thread-per-problem iLQR kernels for small environments (NavigationLQR n<=4, Navigation n=2).
//
// Why thread-per-problem: the headline workload (BASELINE config C3) has n = m = 2, so one
// problem's Riccati state (V_xx: 3 unique values, V_x: 2, K: 4, k: 2) fits in a handful of
// registers and a warp-per-problem mapping would idle 30 of 32 lanes (SURVEY.md section 7,
// "hard parts").  Data layout: the per-problem trajectories live in a struct-of-arrays
// workspace ws[row][slot] with the problem slot fastest, so the 32 lanes of a warp read and
// write 32 consecutive words (one 128-byte line) for every row they touch.
#include "small_core.cuh"

namespace {

constexpr int kThreads = 128;

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_start(EnvSmall e, int64_t B, int T, const real *__restrict__ x0,
                                                    const real *__restrict__ u_init, real *__restrict__ states,
                                                    real *__restrict__ actions, real *__restrict__ costs) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  real x[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = x0[b * N + i];
  CView Ui = {u_init + b * T * M, 1};
  View Xo = {states + b * (T + 1) * N, 1}, Uo = {actions + b * T * M, 1}, Co = {costs + b * (T + 1), 1};
  start_pass<KIND, N, M>(e, T, x, Ui, Xo, Uo, Co);
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_backward(EnvSmall e, int64_t B, int T, const real *__restrict__ states,
                                                       const real *__restrict__ actions, real mu, real *__restrict__ K,
                                                       real *__restrict__ k, real *__restrict__ J, real *__restrict__ dV1,
                                                       real *__restrict__ dV2, int32_t *__restrict__ status) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  CView X = {states + b * (T + 1) * N, 1}, U = {actions + b * T * M, 1};
  View Kv = {K + b * T * M * N, 1}, kv = {k + b * T * M, 1};
  real Jb, d1, d2, g;
  int st = backward_pass<KIND, N, M>(e, T, X, U, mu, Kv, kv, Jb, d1, d2, g);
  J[b] = Jb; dV1[b] = d1; dV2[b] = d2;
  if (status) status[b] = st;
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_forward(EnvSmall e, int64_t B, int T, const real *__restrict__ states,
                                                      const real *__restrict__ actions, const real *__restrict__ K,
                                                      const real *__restrict__ k, real alpha, real *__restrict__ xs,
                                                      real *__restrict__ us, real *__restrict__ cs, real *__restrict__ J,
                                                      real *__restrict__ residual) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  CView X = {states + b * (T + 1) * N, 1}, U = {actions + b * T * M, 1}, Kv = {K + b * T * M * N, 1}, kv = {k + b * T * M, 1};
  View Xo = {xs + b * (T + 1) * N, 1}, Uo = {us + b * T * M, 1}, Co = {cs + b * (T + 1), 1};
  real Jb, res;
  forward_pass<KIND, N, M>(e, T, X, U, Kv, kv, alpha, Xo, Uo, Co, Jb, res);
  J[b] = Jb; residual[b] = res;
}

// rows of the struct-of-arrays workspace, per problem slot
template <int N, int M>
__host__ __device__ constexpr int64_t ws_rows(int T) {
  return 2 * ((int64_t)(T + 1) * N + (int64_t)T * M) + (int64_t)T * M * N + (int64_t)T * M;
}

// The whole iLQR.solve for one problem per thread: start rollout, then the outer loop with the
// mu/delta schedule, convergence tests and line search all on the device (no host round trip).
template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_solve(EnvSmall e, IlqrOpts o, int64_t B, int64_t S, int T,
                                                    const real *__restrict__ x0, const real *__restrict__ u_init,
                                                    real *__restrict__ ws, real *__restrict__ states, real *__restrict__ actions,
                                                    real *__restrict__ costs, int32_t *__restrict__ stats) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int64_t nx = (int64_t)(T + 1) * N, nu = (int64_t)T * M;
  real *base = ws + b;
  View X[2] = {{base, S}, {base + (nx + nu) * S, S}};
  View U[2] = {{base + nx * S, S}, {base + (2 * nx + nu) * S, S}};
  View Kv = {base + 2 * (nx + nu) * S, S};
  View kv = {base + (2 * (nx + nu) + nu * N) * S, S};
  const View none = {nullptr, 0};
  {
    real x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = x0[b * N + i];
    CView Ui = {u_init + b * nu, 1};
    start_pass<KIND, N, M>(e, T, x, Ui, X[0], U[0], none);
  }
  int32_t st[4];
  int cur = solve_one<KIND, N, M>(e, o, T, X, U, Kv, kv, st);
  // emit the converged nominal in the reference's layouts; costs are cost(x_t, u_t) of that
  // nominal, i.e. exactly what the accepting forward pass (or start) computed (ilqr.py:199,208)
  real *so = states + b * nx, *ao = actions + b * nu, *co = costs + b * (T + 1);
  real x[N], u[M];
  for (int t = 0; t < T; t++) {
#pragma unroll
    for (int i = 0; i < N; i++) { x[i] = X[cur](t * N + i); so[t * N + i] = x[i]; }
#pragma unroll
    for (int i = 0; i < M; i++) { u[i] = U[cur](t * M + i); ao[t * M + i] = u[i]; }
    co[t] = env_cost<KIND, N, M>(e, x, u);
  }
#pragma unroll
  for (int i = 0; i < N; i++) { x[i] = X[cur](T * N + i); so[T * N + i] = x[i]; }
  co[T] = env_final_cost<KIND, N, M>(e, x);
#pragma unroll
  for (int i = 0; i < 4; i++) stats[b * 4 + i] = st[i];
}

template <int M>
__global__ void __launch_bounds__(kThreads) k_boxqp(int64_t B, const real *__restrict__ H, const real *__restrict__ q,
                                                    const real *__restrict__ lo, const real *__restrict__ hi, real *__restrict__ x,
                                                    real *__restrict__ Hfree, int32_t *__restrict__ isfree, int32_t *__restrict__ nfree,
                                                    int32_t *__restrict__ status) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  real Hl[M * M], ql[M], lol[M], hil[M], xl[M], L[M * M];
  bool fr[M];
#pragma unroll
  for (int i = 0; i < M * M; i++) Hl[i] = H[b * M * M + i];
#pragma unroll
  for (int i = 0; i < M; i++) { ql[i] = q[b * M + i]; lol[i] = lo[b * M + i]; hil[i] = hi[b * M + i]; xl[i] = x[b * M + i]; }
  int st = boxqp<M>(Hl, ql, lol, hil, xl, L, fr);
  // compact the masked factor into the leading nfree x nfree block, as the reference returns it
  int nf = 0;
  int map[M];
#pragma unroll
  for (int i = 0; i < M; i++) { map[i] = nf; nf += fr[i] ? 1 : 0; }
  for (int i = 0; i < M * M; i++) Hfree[b * M * M + i] = 0;
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < M; j++)
      if (fr[i] && fr[j]) Hfree[b * M * M + map[i] * M + map[j]] = L[i * M + j];
#pragma unroll
  for (int i = 0; i < M; i++) { x[b * M + i] = xl[i]; isfree[b * M + i] = fr[i] ? 1 : 0; }
  nfree[b] = nf;
  status[b] = st;
}

inline unsigned grid_for(int64_t B) { return (unsigned)((B + kThreads - 1) / kThreads); }

}  // namespace

// (kind, n) -> template instantiation
#define SMALL_DISPATCH(e, CALL)                                                               \
  do {                                                                                        \
    if ((e)->kind == TFMPC_ENV_NAVIGATION && (e)->n == 2) { CALL(TFMPC_ENV_NAVIGATION, 2, 2); } \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 1) { CALL(TFMPC_ENV_NAVLQR, 1, 1); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 2) { CALL(TFMPC_ENV_NAVLQR, 2, 2); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 3) { CALL(TFMPC_ENV_NAVLQR, 3, 3); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 4) { CALL(TFMPC_ENV_NAVLQR, 4, 4); }   \
    else return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "no thread-per-problem kernel for kind=%d n=%d", (e)->kind, (e)->n); \
  } while (0)

int small_ilqr_start(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions,
                     real *costs, cudaStream_t s) {
#define CALL(K, N, M) k_start<K, N, M><<<grid_for(B), kThreads, 0, s>>>(e->es, B, T, x0, u_init, states, actions, costs)
  SMALL_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int small_ilqr_backward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, double mu, real *K, real *k,
                        real *J, real *dV1, real *dV2, int32_t *status, cudaStream_t s) {
#define CALL(KD, N, M) k_backward<KD, N, M><<<grid_for(B), kThreads, 0, s>>>(e->es, B, T, states, actions, (real)mu, K, k, J, dV1, dV2, status)
  SMALL_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int small_ilqr_forward(const tfmpc_env *e, int64_t B, int T, const real *states, const real *actions, const real *K, const real *k,
                       double alpha, real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s) {
#define CALL(KD, N, M) k_forward<KD, N, M><<<grid_for(B), kThreads, 0, s>>>(e->es, B, T, states, actions, K, k, (real)alpha, xs, us, cs, J, residual)
  SMALL_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

static int64_t padded_slots(int64_t B) { return (B + 31) / 32 * 32; }

int64_t small_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T) {
  int64_t N = e->n, M = e->m;
  int64_t rows = 2 * ((int64_t)(T + 1) * N + (int64_t)T * M) + (int64_t)T * M * N + (int64_t)T * M;
  return rows * padded_slots(B) * (int64_t)sizeof(real);
}

int small_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s) {
  if (ws_bytes < small_ilqr_workspace_bytes(e, B, T)) return tfmpc_set_error(TFMPC_E_WORKSPACE, "workspace too small");
  int64_t S = padded_slots(B);
#define CALL(KD, N, M) k_solve<KD, N, M><<<grid_for(B), kThreads, 0, s>>>(e->es, o, B, S, T, x0, u_init, (real *)ws, states, actions, costs, stats)
  SMALL_DISPATCH(e, CALL);
#undef CALL
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int small_boxqp(int64_t B, int m, const real *H, const real *q, const real *lo, const real *hi, real *x, real *Hfree, int32_t *isfree,
                int32_t *nfree, int32_t *status, cudaStream_t s) {
  switch (m) {
#define CASE(MM) case MM: k_boxqp<MM><<<grid_for(B), kThreads, 0, s>>>(B, H, q, lo, hi, x, Hfree, isfree, nfree, status); break;
    CASE(1) CASE(2) CASE(3) CASE(4) CASE(5) CASE(6) CASE(7) CASE(8)
#undef CASE
    default: return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "box-QP stage kernel supports m <= 8, got %d", m);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}
