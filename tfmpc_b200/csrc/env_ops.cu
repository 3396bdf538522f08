// Batched environment operators behind the DiffEnv surface of the reference
// (tfmpc/envs/diffenv.py:13-101 and the four env classes): transition, cost, final_cost and the
// analytic linearisation, for R independent (state, action) rows.
//
// Small environments (n <= 4): one thread per row; for the headline n = m = 2 case every output
// block (f_x, f_u, l_xx, l_uu, l_ux, l_xu: 4 values each) is a single 16-byte store per thread,
// so a warp writes 512 contiguous bytes per matrix.
// Large environments (Reservoir, HVAC): one warp per row, lane j owns column j, so every
// [n x n] matrix row is written as one contiguous segment.
#include "rng.cuh"
#include "small_core.cuh"

namespace {

constexpr int kThreads = 128;
constexpr unsigned FULL = 0xffffffffu;

template <int CNT>
__device__ __forceinline__ void store_block(real *dst, const real *v) {
#ifndef TFMPC_F64
  if (CNT % 4 == 0) {  // dst + row * CNT is 16-byte aligned whenever the base pointer is
#pragma unroll
    for (int i = 0; i < CNT / 4; i++) reinterpret_cast<float4 *>(dst)[i] = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    return;
  }
  if (CNT % 2 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 2; i++) reinterpret_cast<float2 *>(dst)[i] = make_float2(v[2 * i], v[2 * i + 1]);
    return;
  }
#else
  if (CNT % 2 == 0) {
#pragma unroll
    for (int i = 0; i < CNT / 2; i++) reinterpret_cast<double2 *>(dst)[i] = make_double2(v[2 * i], v[2 * i + 1]);
    return;
  }
#endif
#pragma unroll
  for (int i = 0; i < CNT; i++) dst[i] = v[i];
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) ks_step(EnvSmall e, int64_t R, const real *__restrict__ x, const real *__restrict__ u,
                                                    real *__restrict__ xn, real *__restrict__ cost) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  real xl[N], ul[M], o[N];
#pragma unroll
  for (int i = 0; i < N; i++) xl[i] = x[r * N + i];
#pragma unroll
  for (int i = 0; i < M; i++) ul[i] = u[r * M + i];
  if (xn) { env_step<KIND, N, M>(e, xl, ul, o); store_block<N>(xn + r * N, o); }
  if (cost) cost[r] = env_cost<KIND, N, M>(e, xl, ul);
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) ks_final(EnvSmall e, int64_t R, const real *__restrict__ x, real *__restrict__ l,
                                                     real *__restrict__ l_x, real *__restrict__ l_xx) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  real xl[N], lv, lx[N], lxx[N * N];
#pragma unroll
  for (int i = 0; i < N; i++) xl[i] = x[r * N + i];
  env_final_quad<KIND, N, M>(e, xl, lv, lx, lxx);
  if (l) l[r] = lv;
  if (l_x) store_block<N>(l_x + r * N, lx);
  if (l_xx) store_block<N * N>(l_xx + r * N * N, lxx);
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) ks_linearize(EnvSmall e, int64_t R, const real *__restrict__ x, const real *__restrict__ u,
                                                         real *__restrict__ f_x, real *__restrict__ f_u, real *__restrict__ l,
                                                         real *__restrict__ l_x, real *__restrict__ l_u, real *__restrict__ l_xx,
                                                         real *__restrict__ l_uu, real *__restrict__ l_ux, real *__restrict__ l_xu) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  real xl[N], ul[M];
#pragma unroll
  for (int i = 0; i < N; i++) xl[i] = x[r * N + i];
#pragma unroll
  for (int i = 0; i < M; i++) ul[i] = u[r * M + i];
  Lin<N, M> L;
  env_linearize<KIND, N, M>(e, xl, ul, L);
  if (f_x) store_block<N * N>(f_x + r * N * N, L.f_x);
  if (f_u) store_block<N * M>(f_u + r * N * M, L.f_u);
  if (l) l[r] = L.l;
  if (l_x) store_block<N>(l_x + r * N, L.l_x);
  if (l_u) store_block<M>(l_u + r * M, L.l_u);
  if (l_xx) store_block<N * N>(l_xx + r * N * N, L.l_xx);
  if (l_uu) store_block<M * M>(l_uu + r * M * M, L.l_uu);
  if (l_xu) store_block<N * M>(l_xu + r * N * M, L.l_xu);
  if (l_ux) {
    real t[M * N];
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
      for (int j = 0; j < N; j++) t[i * N + j] = L.l_xu[j * M + i];
    store_block<M * N>(l_ux + r * M * N, t);
  }
}

// ---------------------------------------------------------------- large envs, warp per row
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}

struct LargeLane {  // parameters of component j of a Reservoir / HVAC env
  real p[9];
};

// large NavigationLQR (n > 4): vec rows = goal, beta (lane 0), low, high
__device__ __forceinline__ real large_cost_term(int kind, const real *p, real x, real u, bool final) {
  if (kind == TFMPC_ENV_NAVLQR) {
    const real beta = __shfl_sync(FULL, p[1], 0);
    const real c1 = (x - p[0]) * (x - p[0]);
    return final ? c1 : c1 + beta * (u * u);   // per-lane share of c1 + beta c2 (lqr/navigation/__init__.py:34-41)
  }
  if (kind == TFMPC_ENV_RESERVOIR) {
    real c1 = -p[3] * r_max((real)0, p[1] - x);
    real c2 = -p[4] * r_max((real)0, x - p[2]);
    real c3 = -p[5] * r_abs((p[1] + p[2]) / (real)2.0 - x);
    return c1 + c2 + c3;
  }
  real oob = (real)20000 * (r_max((real)0, p[0] - x) + r_max((real)0, x - p[1]));
  real sp = (real)10.0 * r_abs((p[0] + p[1]) / (real)2 - x);
  return final ? oob + sp : (real)1.0 * (u * p[3]) + oob + sp;
}

__device__ __forceinline__ real large_l_x(int kind, const real *p, real x) {
  if (kind == TFMPC_ENV_NAVLQR) return (real)2 * (x - p[0]);
  if (kind == TFMPC_ENV_RESERVOIR) {
    real mid = (p[1] + p[2]) / (real)2.0;
    return p[3] * (real)(p[1] - x > 0) - p[4] * (real)(x - p[2] > 0) + p[5] * r_sgn(mid - x);
  }
  real mid = (p[0] + p[1]) / (real)2;
  return (real)20000 * ((real)(x - p[1] > 0) - (real)(p[0] - x > 0)) - (real)10.0 * r_sgn(mid - x);
}

// x_next, cost, final cost for one row per warp
__global__ void __launch_bounds__(kThreads) kl_step(EnvLarge e, int64_t R, const real *__restrict__ x, const real *__restrict__ u,
                                                    real *__restrict__ xn, real *__restrict__ cost, real *__restrict__ fcost) {
  const int lane = threadIdx.x % 32, n = e.n;
  const bool act = lane < n;
  real p[9];
#pragma unroll
  for (int r = 0; r < 9; r++) p[r] = e.vec[r * 32 + lane];
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32, nw = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = w; r < R; r += nw) {
    real xv = act ? x[r * n + lane] : (real)0;
    real uv = (act && u) ? u[r * n + lane] : (real)0;
    if (xn) {
      real o;
      if (e.kind == TFMPC_ENV_NAVLQR) {
        o = xv + uv;
      } else if (e.kind == TFMPC_ENV_RESERVOIR) {
        real out = uv * xv, inflow = 0;
        for (int j = 0; j < n; j++) inflow += e.matF[lane * 32 + j] * __shfl_sync(FULL, out, j);
        real cap = act ? p[0] : (real)1;
        real vap = (real)0.5 * r_sin(xv / cap) * xv;
        o = xv + p[6] + inflow - vap - out;
      } else {
        real air = uv * p[3];
        real heating = air * (real)1.006 * ((real)40.0 - xv);
        real cbr = 0;
        for (int j = 0; j < n; j++) cbr += -e.matF[lane * 32 + j] * (xv - __shfl_sync(FULL, xv, j));
        real cwo = p[4] * (p[6] - xv), cwh = p[5] * (p[7] - xv);
        o = xv + p[2] * (heating + cbr + cwo + cwh);
      }
      if (act) xn[r * n + lane] = o;
    }
    if (cost) {
      const real ct = large_cost_term(e.kind, p, xv, uv, false);
      real c = warp_sum(act ? ct : (real)0);
      if (lane == 0) cost[r] = c;
    }
    if (fcost) {
      const real ct = large_cost_term(e.kind, p, xv, (real)0, true);
      real c = warp_sum(act ? ct : (real)0);
      if (lane == 0) fcost[r] = c;
    }
  }
}

// analytic f_x, f_u, l, l_x, l_u and the (identically zero) second-order blocks; closed forms of
// SURVEY.md Appendix B, pinned by reference tests/test_env_reservoir.py:151-233, test_env_hvac.py:92-98,170-211
__global__ void __launch_bounds__(kThreads) kl_linearize(EnvLarge e, int64_t R, const real *__restrict__ x, const real *__restrict__ u,
                                                         real *__restrict__ f_x, real *__restrict__ f_u, real *__restrict__ l,
                                                         real *__restrict__ l_x, real *__restrict__ l_u, real *__restrict__ l_xx,
                                                         real *__restrict__ l_uu, real *__restrict__ l_ux, real *__restrict__ l_xu,
                                                         int final_only) {
  const int lane = threadIdx.x % 32, n = e.n;
  const bool act = lane < n;
  real p[9];
#pragma unroll
  for (int r = 0; r < 9; r++) p[r] = e.vec[r * 32 + lane];
  int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 32, nw = (int64_t)gridDim.x * blockDim.x / 32;
  for (int64_t r = w; r < R; r += nw) {
    real xv = act ? x[r * n + lane] : (real)0;
    real uv = (act && u) ? u[r * n + lane] : (real)0;
    if (l) {
      const real ct = large_cost_term(e.kind, p, xv, uv, final_only != 0);
      real c = warp_sum(act ? ct : (real)0);
      if (lane == 0) l[r] = c;
    }
    if (l_x && act) l_x[r * n + lane] = large_l_x(e.kind, p, xv);
    const real beta_nl = __shfl_sync(FULL, p[1], 0);  // NavigationLQR: beta lives in lane 0 of row 1
    if (l_u && act) l_u[r * n + lane] = (e.kind == TFMPC_ENV_HVAC) ? p[3] : (e.kind == TFMPC_ENV_NAVLQR ? (real)2 * beta_nl * uv : (real)0);
    // lane j computes column j of row i
    real diag;  // f_x[j][j] extra term of this lane
    if (e.kind == TFMPC_ENV_NAVLQR) {
      diag = 1;
    } else if (e.kind == TFMPC_ENV_RESERVOIR) {
      real cap = act ? p[0] : (real)1;
      real a = xv / cap;
      diag = (real)1 - (real)0.5 * (r_cos(a) * a + r_sin(a)) - uv;
    } else {
      diag = (real)1 + p[2] * (-(uv * p[3]) * (real)1.006 - p[8] - p[4] - p[5]);
    }
    for (int i = 0; i < n; i++) {
      real fx, fu;
      real hxx = 0, huu = 0;  // second-order cost blocks (non-zero for NavigationLQR only)
      if (e.kind == TFMPC_ENV_NAVLQR) {  // f_x = f_u = I, l_xx = 2I, l_uu = 2 beta I (lqr/navigation/__init__.py:30-47)
        fx = (i == lane) ? (real)1 : (real)0;
        fu = fx;
        hxx = (i == lane) ? (real)2 : (real)0;
        huu = (i == lane) ? (real)2 * beta_nl : (real)0;
      } else if (e.kind == TFMPC_ENV_RESERVOIR) {  // f_x[i][j] = D[j][i] u_j (+diag), f_u[i][j] = D[j][i] x_j (- x_i on the diagonal)
        real d = act ? e.matB[lane * 32 + i] : (real)0;
        fx = d * uv + (i == lane ? diag : (real)0);
        fu = d * xv + (i == lane ? -xv : (real)0);
      } else {  // f_x[i][j] = s_i A[i][j] (+diag), f_u diagonal
        real s_i = __shfl_sync(FULL, p[2], i);
        fx = s_i * (act ? e.matF[i * 32 + lane] : (real)0) + (i == lane ? diag : (real)0);
        fu = (i == lane) ? p[2] * (p[3] * (real)1.006 * ((real)40.0 - xv)) : (real)0;
      }
      if (act) {
        int64_t o = (r * n + i) * n + lane;
        if (f_x) f_x[o] = fx;
        if (f_u) f_u[o] = fu;
        if (l_xx) l_xx[o] = hxx;
        if (l_uu) l_uu[o] = final_only ? (real)0 : huu;
        if (l_ux) l_ux[o] = 0;
        if (l_xu) l_xu[o] = 0;
      }
    }
  }
}

inline unsigned grid_rows(int64_t R) { return (unsigned)((R + kThreads - 1) / kThreads); }
inline unsigned grid_warps(int64_t R) {
  int64_t blocks = (R + kThreads / 32 - 1) / (kThreads / 32);
  return (unsigned)(blocks < 148 * 16 ? blocks : 148 * 16);
}

int check_large(const tfmpc_env *e) {
  if (e->kind != TFMPC_ENV_RESERVOIR && e->kind != TFMPC_ENV_HVAC && e->kind != TFMPC_ENV_NAVLQR)
    return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "no batched operator for environment kind %d with n=%d", e->kind, e->n);
  return TFMPC_OK;
}

}  // namespace

#define SMALL_DISPATCH(e, CALL)                                                               \
  do {                                                                                        \
    if ((e)->kind == TFMPC_ENV_NAVIGATION && (e)->n == 2) { CALL(TFMPC_ENV_NAVIGATION, 2, 2); } \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 1) { CALL(TFMPC_ENV_NAVLQR, 1, 1); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 2) { CALL(TFMPC_ENV_NAVLQR, 2, 2); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 3) { CALL(TFMPC_ENV_NAVLQR, 3, 3); }   \
    else if ((e)->kind == TFMPC_ENV_NAVLQR && (e)->n == 4) { CALL(TFMPC_ENV_NAVLQR, 4, 4); }   \
    else return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "no thread-per-row kernel for kind=%d n=%d", (e)->kind, (e)->n); \
  } while (0)

int env_ops_step(const tfmpc_env *e, int64_t R, const real *x, const real *u, real *xn, real *cost, cudaStream_t s) {
  if (e->small) {
#define CALL(K, N, M) ks_step<K, N, M><<<grid_rows(R), kThreads, 0, s>>>(e->es, R, x, u, xn, cost)
    SMALL_DISPATCH(e, CALL);
#undef CALL
  } else {
    int rc = check_large(e);
    if (rc) return rc;
    kl_step<<<grid_warps(R), kThreads, 0, s>>>(e->el, R, x, u, xn, cost, nullptr);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}

// ---- stochastic plant (GymEnv.step -> transition(cec=False), reference tfmpc/envs/gymenv.py:18)
// One thread per state component.  Navigation: x' += truncated normal(0, 0.2) (navigation/__init__.py:45).  Reservoir: the
// rainfall term of x' (reservoir/__init__.py:57) becomes a Gamma(rain_shape, rain_scale) draw instead of its mean (:98-105),
// i.e. x' += rain - shape * scale.  The other environments have no noise model in the reference (their transition() takes
// no `cec` argument).
__global__ void __launch_bounds__(256) k_plant_noise(int kind, int n, int64_t R, real *xn, const real *vec, unsigned long long seed,
                                                     unsigned long long offset) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= R * n) return;
  const int64_t r = idx / n;
  const int i = (int)(idx - r * n);
  Philox g(seed ^ (offset * 0x9E3779B97F4A7C15ull), (uint32_t)r, (uint32_t)((uint64_t)r >> 32), (uint32_t)i, 0x504cu);   // tag "PL"ant
  if (kind == TFMPC_ENV_NAVIGATION) {
    uint32_t w[4];
    g.next(w);
    xn[idx] += (real)(0.2 * trunc_normal2(u01(w[0], w[1])));
  } else if (kind == TFMPC_ENV_RESERVOIR) {
    const double shape = (double)vec[7 * 32 + i], scale = (double)vec[8 * 32 + i];
    xn[idx] += (real)(gamma_draw(g, shape, scale) - shape * scale);
  }
}

// ---- iLQR.start's random initial actions (ilqr.py:59-70): ONE uniform scalar per (problem, step), broadcast over the action
// dimensions, scaled to [low, high] with infinite bounds replaced by -1 / +1
struct ActionBounds { real lo[MAXD], hi[MAXD]; };
__global__ void __launch_bounds__(256) k_initial_actions(int m, int64_t BT, ActionBounds b, real *u, unsigned long long seed) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= BT) return;
  Philox g(seed, (uint32_t)idx, (uint32_t)((uint64_t)idx >> 32), 0u, 0x5354u);   // tag "ST"art
  uint32_t w[4];
  g.next(w);
  const real r = (real)u01(w[0], w[1]);
  for (int i = 0; i < m; i++) u[idx * m + i] = b.lo[i] + r * (b.hi[i] - b.lo[i]);
}

int env_ops_plant_noise(const tfmpc_env *e, int64_t R, real *xn, unsigned long long seed, unsigned long long offset, cudaStream_t s) {
  if (e->kind != TFMPC_ENV_NAVIGATION && e->kind != TFMPC_ENV_RESERVOIR) return TFMPC_OK;   // no noise model in the reference
  const int64_t items = R * e->n;
  k_plant_noise<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(e->kind, e->n, R, xn, e->el.vec, seed, offset);
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int env_ops_initial_actions(const tfmpc_env *e, int64_t B, int T, unsigned long long seed, real *u_init, cudaStream_t s) {
  ActionBounds b;
  for (int i = 0; i < e->m; i++) {
    b.lo[i] = (real)(std::isinf(e->low[i]) ? -1.0 : e->low[i]);
    b.hi[i] = (real)(std::isinf(e->high[i]) ? 1.0 : e->high[i]);
  }
  const int64_t items = B * T;
  k_initial_actions<<<(unsigned)((items + 255) / 256), 256, 0, s>>>(e->m, items, b, u_init, seed);
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int env_ops_final_cost(const tfmpc_env *e, int64_t R, const real *x, real *cost, cudaStream_t s) {
  if (e->small) {
#define CALL(K, N, M) ks_final<K, N, M><<<grid_rows(R), kThreads, 0, s>>>(e->es, R, x, cost, nullptr, nullptr)
    SMALL_DISPATCH(e, CALL);
#undef CALL
  } else {
    int rc = check_large(e);
    if (rc) return rc;
    kl_step<<<grid_warps(R), kThreads, 0, s>>>(e->el, R, x, nullptr, nullptr, nullptr, cost);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int env_ops_linearize(const tfmpc_env *e, int64_t R, const real *x, const real *u, real *f_x, real *f_u, real *l, real *l_x, real *l_u,
                      real *l_xx, real *l_uu, real *l_ux, real *l_xu, cudaStream_t s) {
  if (e->small) {
#define CALL(K, N, M) ks_linearize<K, N, M><<<grid_rows(R), kThreads, 0, s>>>(e->es, R, x, u, f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_ux, l_xu)
    SMALL_DISPATCH(e, CALL);
#undef CALL
  } else {
    int rc = check_large(e);
    if (rc) return rc;
    kl_linearize<<<grid_warps(R), kThreads, 0, s>>>(e->el, R, x, u, f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_ux, l_xu, 0);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int env_ops_final_quad(const tfmpc_env *e, int64_t R, const real *x, real *l, real *l_x, real *l_xx, cudaStream_t s) {
  if (e->small) {
#define CALL(K, N, M) ks_final<K, N, M><<<grid_rows(R), kThreads, 0, s>>>(e->es, R, x, l, l_x, l_xx)
    SMALL_DISPATCH(e, CALL);
#undef CALL
  } else {
    int rc = check_large(e);
    if (rc) return rc;
    // final cost of both large envs has l_xx = 0 and the same l_x as the stage cost
    kl_linearize<<<grid_warps(R), kThreads, 0, s>>>(e->el, R, x, nullptr, nullptr, nullptr, l, l_x, nullptr, l_xx, nullptr, nullptr, nullptr, 1);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}
