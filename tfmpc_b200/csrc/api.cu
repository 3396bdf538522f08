// extern "C" boundary of libtfmpc_b200 (see include/tfmpc_b200.h for the contract).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "common.cuh"
#include <mutex>

// ------------------------------------------------------------------ errors / counters
static thread_local char g_err[512] = "";
static std::atomic<int64_t> g_launches{0};

int tfmpc_set_error(int code, const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
void tfmpc_count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ------------------------------------------------------------------ DLPack (ABI v0.x) -- own declaration of the public standard layout
namespace {
struct DLDevice { int32_t device_type; int32_t device_id; };
struct DLDataType { uint8_t code; uint8_t bits; uint16_t lanes; };
struct DLTensor {
  void *data; DLDevice device; int32_t ndim; DLDataType dtype; int64_t *shape; int64_t *strides; uint64_t byte_offset;
};
struct DLManagedTensor { DLTensor dl_tensor; void *manager_ctx; void (*deleter)(DLManagedTensor *); };
enum { kDLCPU = 1, kDLCUDA = 2, kDLCUDAHost = 3, kDLCUDAManaged = 13 };
enum { kDLInt = 0, kDLUInt = 1, kDLFloat = 2 };

void line_search_alphas(double alpha_min, real *out) {
  // np.geomspace(1.0, alpha_min, 11) (reference ilqr.py:322); the default table is numpy's own output
  static const double dflt[N_ALPHA] = {1.0, 0.5011872336272722, 0.251188643150958, 0.12589254117941676, 0.06309573444801933,
                                       0.03162277660168379, 0.01584893192461114, 0.007943282347242814, 0.003981071705534973,
                                       0.0019952623149688807, 0.001};
  double a[N_ALPHA];
  if (alpha_min == 1e-3) {
    memcpy(a, dflt, sizeof(dflt));
  } else {
    double l = log10(alpha_min);
    for (int i = 0; i < N_ALPHA; i++) a[i] = pow(10.0, l * i / (N_ALPHA - 1.0));
    a[0] = 1.0; a[N_ALPHA - 1] = alpha_min;
  }
  for (int i = 0; i < N_ALPHA; i++) out[i] = (real)a[i];
}

int make_opts(const tfmpc_ilqr_opts_t *in, IlqrOpts *o) {
  tfmpc_ilqr_opts_t d;
  tfmpc_ilqr_default_opts(&d);
  if (in) d = *in;
  if (d.max_iterations < 1 || !(d.atol >= 0) || !(d.delta_0 > 0) || !(d.alpha_min > 0) || !(d.alpha_min <= 1))
    return tfmpc_set_error(TFMPC_E_INVALID, "bad iLQR options");
  o->atol = (real)d.atol; o->c1 = (real)d.c1; o->max_iterations = d.max_iterations; o->mu_min = d.mu_min; o->delta_0 = d.delta_0;
  line_search_alphas(d.alpha_min, o->alphas);
  return TFMPC_OK;
}

bool use_small(const tfmpc_env *e) { return e->small != 0; }

int env_choice(const char *name, const char *a, int va, const char *b, int vb, int dflt) {
  const char *v = getenv(name);
  if (!v || !*v) return dflt;
  if (!strcmp(v, a)) return va;
  if (!strcmp(v, b)) return vb;
  return atoi(v);
}
// solver: 1 = persistent work-queue kernel, 0 = per-tick launch sequence
std::atomic<int> g_solver{env_choice("TFMPC_SOLVER", "queue", 1, "ticks", 0, 1)};
// qp: 2 = closed form (m <= 2), 0 = the reference's projected-Newton iteration.  The fp64 verification build defaults to
// the reference's iteration so that it reproduces the fp64 oracle exactly.
#ifdef TFMPC_F64
std::atomic<int> g_qp{env_choice("TFMPC_QP", "closed", 2, "newton", 0, 0)};
#else
std::atomic<int> g_qp{env_choice("TFMPC_QP", "closed", 2, "newton", 0, 2)};
#endif
bool use_queue(const tfmpc_env *e) { return use_small(e) && g_solver.load() != 0; }
}  // namespace

extern "C" {

int tfmpc_abi_version(void) { return TFMPC_ABI_VERSION; }
int tfmpc_real_bytes(void) { return (int)sizeof(real); }
const char *tfmpc_last_error(void) { return g_err; }
int64_t tfmpc_kernel_launch_count(void) { return g_launches.load(); }

void tfmpc_ilqr_default_opts(tfmpc_ilqr_opts_t *o) {  // reference ilqr.py:27-37
  o->atol = 5e-3; o->max_iterations = 100; o->mu_min = 1e-6; o->delta_0 = 2.0; o->c1 = 0.0; o->alpha_min = 1e-3;
}

int tfmpc_dl_unpack(const void *p, int want_device, int want_code, void **data, int32_t *ndim, int64_t *shape, int32_t *device_id) {
  if (!p || !data) return tfmpc_set_error(TFMPC_E_DLPACK, "null DLManagedTensor");
  const DLTensor &t = ((const DLManagedTensor *)p)->dl_tensor;
  bool is_cuda = t.device.device_type == kDLCUDA || t.device.device_type == kDLCUDAManaged;
  bool is_cpu = t.device.device_type == kDLCPU || t.device.device_type == kDLCUDAHost;
  if (want_device == 2 && !is_cuda) return tfmpc_set_error(TFMPC_E_DLPACK, "tensor is not on a CUDA device (device_type=%d)", t.device.device_type);
  if (want_device == 1 && !is_cpu) return tfmpc_set_error(TFMPC_E_DLPACK, "tensor is not in host memory (device_type=%d)", t.device.device_type);
  if (t.dtype.lanes != 1) return tfmpc_set_error(TFMPC_E_DLPACK, "vector dtypes are not supported");
  if (want_code == 0) {
    if (t.dtype.code != kDLFloat || t.dtype.bits != 8 * sizeof(real))
      return tfmpc_set_error(TFMPC_E_DLPACK, "expected float%d, got code=%d bits=%d", (int)(8 * sizeof(real)), t.dtype.code, t.dtype.bits);
  } else {
    if (t.dtype.code != kDLInt || t.dtype.bits != 32) return tfmpc_set_error(TFMPC_E_DLPACK, "expected int32, got code=%d bits=%d", t.dtype.code, t.dtype.bits);
  }
  if (t.ndim < 0 || t.ndim > 8) return tfmpc_set_error(TFMPC_E_DLPACK, "rank %d out of range", t.ndim);
  if (t.strides) {  // must be compact row-major (size-1 dims may carry any stride)
    int64_t expect = 1;
    for (int i = t.ndim - 1; i >= 0; i--) {
      if (t.shape[i] != 1 && t.strides[i] != expect) return tfmpc_set_error(TFMPC_E_DLPACK, "tensor is not contiguous (dim %d stride %lld)", i, (long long)t.strides[i]);
      expect *= t.shape[i];
    }
  }
  if (ndim) *ndim = t.ndim;
  if (shape) for (int i = 0; i < t.ndim; i++) shape[i] = t.shape[i];
  if (device_id) *device_id = t.device.device_id;
  *data = (char *)t.data + t.byte_offset;
  return TFMPC_OK;
}

// ------------------------------------------------------------------ environments
int tfmpc_env_create(int kind, int n, int m, int nz, const double *p, int64_t nparams, tfmpc_env_t **out) {
  if (!p || !out) return tfmpc_set_error(TFMPC_E_INVALID, "null argument");
  if (n < 1 || n > MAXD || m < 1 || m > MAXD || nz < 0 || nz > MAXZ) return tfmpc_set_error(TFMPC_E_INVALID, "dimension out of range (n=%d m=%d nz=%d)", n, m, nz);
  int64_t need = 0;
  switch (kind) {
    case TFMPC_ENV_NAVLQR: need = n + 1 + 2 * m; if (n != m) return tfmpc_set_error(TFMPC_E_INVALID, "NavigationLQR needs n == m"); break;
    case TFMPC_ENV_NAVIGATION: need = 6 + 3 * nz; if (n != 2 || m != 2) return tfmpc_set_error(TFMPC_E_INVALID, "Navigation is 2-D"); break;
    case TFMPC_ENV_RESERVOIR: need = 8 * n + (int64_t)n * n; if (n != m) return tfmpc_set_error(TFMPC_E_INVALID, "Reservoir needs n == m"); break;
    case TFMPC_ENV_HVAC: need = 10 * n + 2 * (int64_t)n * n; if (n != m) return tfmpc_set_error(TFMPC_E_INVALID, "HVAC needs n == m"); break;
    default: return tfmpc_set_error(TFMPC_E_INVALID, "unknown environment kind %d", kind);
  }
  if (nparams != need) return tfmpc_set_error(TFMPC_E_INVALID, "kind %d expects %lld parameters, got %lld", kind, (long long)need, (long long)nparams);
  tfmpc_env *e = (tfmpc_env *)calloc(1, sizeof(tfmpc_env));
  static std::atomic<unsigned long long> next_uid{1};
  if (e) e->uid = next_uid.fetch_add(1);
  if (!e) return tfmpc_set_error(TFMPC_E_INVALID, "out of host memory");
  e->kind = kind; e->n = n; e->m = m; e->nz = nz;
  EnvSmall &s = e->es;
  s.kind = kind; s.n = n; s.m = m; s.nz = nz;
  if (kind == TFMPC_ENV_NAVLQR) {
    for (int i = 0; i < m; i++) { e->low[i] = p[n + 1 + i]; e->high[i] = p[n + 1 + m + i]; }
    for (int i = 0; i < n; i++) e->goal[i] = p[i];
    e->beta = p[n];
    e->small = n <= 4;
    if (e->small) {
      for (int i = 0; i < n; i++) { s.goal[i] = (real)p[i]; s.low[i] = (real)e->low[i]; s.high[i] = (real)e->high[i]; }
      s.beta = (real)p[n];
    }
  } else if (kind == TFMPC_ENV_NAVIGATION) {
    for (int i = 0; i < 2; i++) { s.goal[i] = (real)p[i]; e->low[i] = p[2 + i]; e->high[i] = p[4 + i]; s.low[i] = (real)p[2 + i]; s.high[i] = (real)p[4 + i]; }
    for (int z = 0; z < nz; z++) { s.center[z][0] = (real)p[6 + 2 * z]; s.center[z][1] = (real)p[6 + 2 * z + 1]; s.decay[z] = (real)p[6 + 2 * nz + z]; }
    e->small = 1;
  } else {
    for (int i = 0; i < m; i++) { e->low[i] = 0.0; e->high[i] = 1.0; }
    e->small = 0;
  }
  e->bounded = 1;
  for (int i = 0; i < m; i++) if (std::isinf(e->low[i]) || std::isinf(e->high[i])) e->bounded = 0;
  s.bounded = e->bounded;

  if (cudaGetDevice(&e->device) != cudaSuccess) { free(e); return tfmpc_set_error(TFMPC_E_CUDA, "no CUDA device: %s", cudaGetErrorString(cudaGetLastError())); }

  if (e->small) {
    real tab[QP_MAX_STEPS];
    s.qp_klast = qp_step_table(tab);
    if (cudaMalloc((void **)&e->dsteps, sizeof(tab)) != cudaSuccess || cudaMemcpy(e->dsteps, tab, sizeof(tab), cudaMemcpyHostToDevice) != cudaSuccess) {
      int rc = tfmpc_set_error(TFMPC_E_CUDA, "box-QP step table upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      if (e->dsteps) cudaFree(e->dsteps);
      free(e);
      return rc;
    }
    s.qp_steps = e->dsteps;
  }

  if (kind == TFMPC_ENV_RESERVOIR || kind == TFMPC_ENV_HVAC || (kind == TFMPC_ENV_NAVLQR && !e->small)) {
    // device blob for the warp-per-problem kernels: vec[16][32] | matF[32][32] | matB[32][32], zero padded
    const int NV = 16;
    std::vector<real> h((size_t)NV * 32 + 2 * 32 * 32, (real)0);
    real *vec = h.data(), *mF = vec + NV * 32, *mB = mF + 32 * 32;
    if (kind == TFMPC_ENV_RESERVOIR) {  // reference tfmpc/envs/reservoir/__init__.py:11-37
      for (int i = 0; i < n; i++) {
        for (int r = 0; r < 6; r++) vec[r * 32 + i] = (real)p[r * n + i];  // cap lb ub lowpen highpen sppen
        vec[6 * 32 + i] = (real)p[6 * n + i] * (real)p[7 * n + i];          // rain = shape * scale (:98-100)
        vec[7 * 32 + i] = (real)p[6 * n + i];                               // gamma rainfall of the stochastic plant (:102-104)
        vec[8 * 32 + i] = (real)p[7 * n + i];
        for (int j = 0; j < n; j++) {
          real d_ij = (real)p[8 * n + i * n + j];
          mB[i * 32 + j] = d_ij;  // (D V)_i
          mF[j * 32 + i] = d_ij;  // (D^T o)_j
        }
      }
    } else if (kind == TFMPC_ENV_HVAC) {  // reference tfmpc/envs/hvac/__init__.py:17-58
      const double *rw = p + 10 * n, *adj = p + 10 * n + (int64_t)n * n;
      std::vector<real> A((size_t)n * n);
      for (int i = 0; i < n; i++)
        for (int j = 0; j < n; j++) {
          real a = (adj[i * n + j] != 0.0 || adj[j * n + i] != 0.0) ? (real)1 : (real)0;  // :130-131
          A[i * n + j] = a / (real)rw[i * n + j];
        }
      for (int i = 0; i < n; i++) {
        real s_i = (real)1.0 / (real)p[6 * n + i];  // TIME_DELTA / capacity
        real rowsum = 0;
        for (int j = 0; j < n; j++) rowsum += A[i * n + j];
        vec[0 * 32 + i] = (real)p[2 * n + i];                       // lb
        vec[1 * 32 + i] = (real)p[3 * n + i];                       // ub
        vec[2 * 32 + i] = s_i;
        vec[3 * 32 + i] = (real)p[7 * n + i];                       // air_max
        vec[4 * 32 + i] = (real)p[8 * n + i] / (real)p[4 * n + i];  // adj_outside / R_outside
        vec[5 * 32 + i] = (real)p[9 * n + i] / (real)p[5 * n + i];  // adj_hall / R_hall
        vec[6 * 32 + i] = (real)p[0 * n + i];                       // temp_outside
        vec[7 * 32 + i] = (real)p[1 * n + i];                       // temp_hall
        vec[8 * 32 + i] = rowsum;
        for (int j = 0; j < n; j++) {
          mF[i * 32 + j] = A[i * n + j];
          mB[i * 32 + j] = ((real)1.0 / (real)p[6 * n + j]) * A[j * n + i];
        }
      }
    } else {  // large NavigationLQR: goal, beta, low, high rows
      for (int i = 0; i < n; i++) { vec[0 * 32 + i] = (real)p[i]; vec[2 * 32 + i] = (real)e->low[i]; vec[3 * 32 + i] = (real)e->high[i]; }
      vec[1 * 32] = (real)p[n];
    }
    e->max_row_nnz = 0;
    for (int i = 0; i < 32; i++) {
      int cf = 0, cb = 0;
      for (int j = 0; j < 32; j++) { cf += mF[i * 32 + j] != (real)0; cb += mB[i * 32 + j] != (real)0; }
      e->max_row_nnz = std::max(e->max_row_nnz, std::max(cf, cb));
    }
    size_t bytes = h.size() * sizeof(real);
    if (cudaMalloc((void **)&e->dblob, bytes) != cudaSuccess || cudaMemcpy(e->dblob, h.data(), bytes, cudaMemcpyHostToDevice) != cudaSuccess) {
      int rc = tfmpc_set_error(TFMPC_E_CUDA, "env parameter upload failed: %s", cudaGetErrorString(cudaGetLastError()));
      if (e->dblob) cudaFree(e->dblob);
      free(e);
      return rc;
    }
    e->el.kind = kind; e->el.n = n; e->el.m = m;
    e->el.vec = e->dblob; e->el.matF = e->dblob + NV * 32; e->el.matB = e->dblob + NV * 32 + 32 * 32;
  }
  *out = e;
  return TFMPC_OK;
}

int tfmpc_env_destroy(tfmpc_env_t *e) {
  if (!e) return TFMPC_OK;
  small_ilqr_forget(e);
  if (e->dblob) cudaFree(e->dblob);
  if (e->dsteps) cudaFree(e->dsteps);
  if (e->h_scratch) cudaFree(e->h_scratch);
  free(e);
  return TFMPC_OK;
}

int tfmpc_env_info(const tfmpc_env_t *e, int32_t *n, int32_t *m, int32_t *bounded, double *low, double *high) {
  if (!e) return tfmpc_set_error(TFMPC_E_INVALID, "null env");
  if (n) *n = e->n;
  if (m) *m = e->m;
  if (bounded) *bounded = e->bounded;
  for (int i = 0; i < e->m; i++) { if (low) low[i] = e->low[i]; if (high) high[i] = e->high[i]; }
  return TFMPC_OK;
}

#define REQ(cond, msg) do { if (!(cond)) return tfmpc_set_error(TFMPC_E_INVALID, msg); } while (0)

int tfmpc_env_step(const tfmpc_env_t *e, int64_t R, const real *x, const real *u, real *xn, real *cost, void *stream) {
  REQ(e && x && u && R >= 0, "tfmpc_env_step: bad argument");
  if (R == 0) return TFMPC_OK;
  return env_ops_step(e, R, x, u, xn, cost, (cudaStream_t)stream);
}
int tfmpc_env_step_noisy(const tfmpc_env_t *e, int64_t R, const real *x, const real *u, real *xn, real *cost, uint64_t seed, uint64_t offset,
                         void *stream) {
  REQ(e && x && u && xn && R >= 0, "tfmpc_env_step_noisy: bad argument");
  if (R == 0) return TFMPC_OK;
  int rc = env_ops_step(e, R, x, u, xn, cost, (cudaStream_t)stream);
  if (rc) return rc;
  return env_ops_plant_noise(e, R, xn, seed, offset, (cudaStream_t)stream);
}
int tfmpc_env_has_noise_model(const tfmpc_env_t *e) { return e && (e->kind == TFMPC_ENV_NAVIGATION || e->kind == TFMPC_ENV_RESERVOIR) ? 1 : 0; }
int tfmpc_ilqr_initial_actions(const tfmpc_env_t *e, int64_t B, int T, uint64_t seed, real *u_init, void *stream) {
  REQ(e && u_init && B >= 0 && T >= 1, "tfmpc_ilqr_initial_actions: bad argument");
  if (B == 0) return TFMPC_OK;
  return env_ops_initial_actions(e, B, T, seed, u_init, (cudaStream_t)stream);
}
int tfmpc_env_final_cost(const tfmpc_env_t *e, int64_t R, const real *x, real *cost, void *stream) {
  REQ(e && x && cost && R >= 0, "tfmpc_env_final_cost: bad argument");
  if (R == 0) return TFMPC_OK;
  return env_ops_final_cost(e, R, x, cost, (cudaStream_t)stream);
}
int tfmpc_env_linearize(const tfmpc_env_t *e, int64_t R, const real *x, const real *u, real *f_x, real *f_u, real *l, real *l_x, real *l_u,
                        real *l_xx, real *l_uu, real *l_ux, real *l_xu, void *stream) {
  REQ(e && x && u && R >= 0, "tfmpc_env_linearize: bad argument");
  if (R == 0) return TFMPC_OK;
  return env_ops_linearize(e, R, x, u, f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_ux, l_xu, (cudaStream_t)stream);
}
int tfmpc_env_final_quad(const tfmpc_env_t *e, int64_t R, const real *x, real *l, real *l_x, real *l_xx, void *stream) {
  REQ(e && x && R >= 0, "tfmpc_env_final_quad: bad argument");
  if (R == 0) return TFMPC_OK;
  return env_ops_final_quad(e, R, x, l, l_x, l_xx, (cudaStream_t)stream);
}

int tfmpc_boxqp(int64_t B, int m, const real *H, const real *q, const real *low, const real *high, real *x, real *Hfree, int32_t *isfree,
                int32_t *nfree, int32_t *status, void *stream) {
  REQ(H && q && low && high && x && Hfree && isfree && nfree && status && B >= 0 && m >= 1, "tfmpc_boxqp: bad argument");
  if (B == 0) return TFMPC_OK;
  return small_boxqp(B, m, H, q, low, high, x, Hfree, isfree, nfree, status, (cudaStream_t)stream);
}

// ------------------------------------------------------------------ iLQR
int tfmpc_ilqr_start(const tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, real *states, real *actions, real *costs, void *stream) {
  REQ(e && x0 && u_init && states && actions && costs && B >= 0 && T >= 1, "tfmpc_ilqr_start: bad argument");
  if (B == 0) return TFMPC_OK;
  if (!use_small(e) && e->kind == TFMPC_ENV_NAVLQR)
    return dense_navlqr_forward(e, B, T, x0, nullptr, u_init, nullptr, nullptr, 0.0, states, actions, costs, nullptr, nullptr, (cudaStream_t)stream);
  return use_small(e) ? small_ilqr_start(e, B, T, x0, u_init, states, actions, costs, (cudaStream_t)stream)
                      : warp_ilqr_start(e, B, T, x0, u_init, states, actions, costs, (cudaStream_t)stream);
}
int tfmpc_ilqr_backward(const tfmpc_env_t *e, int64_t B, int T, const real *states, const real *actions, double mu, real *K, real *k, real *J,
                        real *dV1, real *dV2, int32_t *status, void *stream) {
  REQ(e && states && actions && K && k && J && dV1 && dV2 && B >= 0 && T >= 1, "tfmpc_ilqr_backward: bad argument");
  if (B == 0) return TFMPC_OK;
  return use_small(e) ? small_ilqr_backward(e, B, T, states, actions, mu, K, k, J, dV1, dV2, status, (cudaStream_t)stream)
                      : warp_ilqr_backward(e, B, T, states, actions, mu, K, k, J, dV1, dV2, status, (cudaStream_t)stream);
}
int tfmpc_ilqr_forward(const tfmpc_env_t *e, int64_t B, int T, const real *states, const real *actions, const real *K, const real *k, double alpha,
                       real *xs, real *us, real *cs, real *J, real *residual, void *stream) {
  REQ(e && states && actions && K && k && xs && us && cs && J && residual && B >= 0 && T >= 1, "tfmpc_ilqr_forward: bad argument");
  if (B == 0) return TFMPC_OK;
  if (!use_small(e) && e->kind == TFMPC_ENV_NAVLQR)
    return dense_navlqr_forward(e, B, T, nullptr, states, actions, K, k, alpha, xs, us, cs, J, residual, (cudaStream_t)stream);
  return use_small(e) ? small_ilqr_forward(e, B, T, states, actions, K, k, alpha, xs, us, cs, J, residual, (cudaStream_t)stream)
                      : warp_ilqr_forward(e, B, T, states, actions, K, k, alpha, xs, us, cs, J, residual, (cudaStream_t)stream);
}

int tfmpc_ilqr_backward_staged(int64_t B, int T, int n, int m, const double *low, const double *high, const real *actions, const real *f_x,
                               const real *f_u, const real *l, const real *l_x, const real *l_u, const real *l_xx, const real *l_uu,
                               const real *l_xu, const real *fl, const real *fl_x, const real *fl_xx, double mu, real *K, real *k, real *J,
                               real *dV1, real *dV2, int32_t *status, void *stream) {
  REQ(low && high && actions && f_x && f_u && l && l_x && l_u && l_xx && l_uu && l_xu && fl && fl_x && fl_xx && K && k && J && dV1 && dV2,
      "tfmpc_ilqr_backward_staged: null argument");
  REQ(B >= 0 && T >= 1 && n >= 1 && m >= 1 && n <= MAXD && m <= MAXD, "tfmpc_ilqr_backward_staged: size out of range");
  if (B == 0) return TFMPC_OK;
  int bounded = 1;  // gym Box.is_bounded(): every bound finite (ilqr.py:136)
  for (int i = 0; i < m; i++) if (std::isinf(low[i]) || std::isinf(high[i])) bounded = 0;
  return dense_backward_launch(B, T, n, m, bounded, low, high, actions, f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_xu, fl, fl_x, fl_xx, mu, K, k, J, dV1,
                               dV2, status, (cudaStream_t)stream);
}

int tfmpc_set_option(const char *name, int value) {
  if (!name) return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_set_option: null name");
  if (!strcmp(name, "solver")) return g_solver.exchange(value ? 1 : 0);
  if (!strcmp(name, "qp")) {
    if (value != 0 && value != 2) return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_set_option: qp must be 0 (newton) or 2 (closed form)");
    return g_qp.exchange(value);
  }
  int prev = 0;
  if (queue_ilqr_option(name, value, &prev)) return prev;
  return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_set_option: unknown option '%s'", name);
}

int tfmpc_ilqr_queue_counters(const void *workspace, int32_t *out, void *stream) {
  REQ(workspace && out, "tfmpc_ilqr_queue_counters: null argument");
  int raw[352];
  int rc = queue_ilqr_counters(workspace, raw, 352, (cudaStream_t)stream);
  if (rc) return rc;
  out[0] = raw[160]; out[1] = raw[192]; out[2] = raw[224]; out[3] = raw[256]; out[4] = raw[128];
  return TFMPC_OK;
}

int64_t tfmpc_ilqr_queue_trace(const tfmpc_env_t *e, int64_t B, int T, const void *workspace, uint32_t *out, int64_t max_records, void *stream) {
  if (!e || !workspace || !out || !use_small(e)) return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_ilqr_queue_trace: bad argument");
  return queue_ilqr_trace(e, B, T, workspace, out, max_records, (cudaStream_t)stream);
}

int64_t tfmpc_ilqr_workspace_bytes(const tfmpc_env_t *e, int64_t B, int T) {
  if (!e || B < 0 || T < 1) return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_ilqr_workspace_bytes: bad argument");
  int64_t b = use_queue(e) ? queue_ilqr_workspace_bytes(e, std::max<int64_t>(B, 1), T) : use_small(e) ? small_ilqr_workspace_bytes(e, B, T)
                           : (e->kind == TFMPC_ENV_NAVLQR ? dense_navlqr_workspace_bytes(e, B, T) : warp_ilqr_workspace_bytes(e, B, T));
  return (b + 255) / 256 * 256;
}

static int ilqr_solve_impl(const tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, const tfmpc_ilqr_opts_t *opts,
                           real *states, real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, void *stream, cudaEvent_t done) {
  REQ(e && x0 && u_init && states && actions && costs && stats && B >= 0 && T >= 1, "tfmpc_ilqr_solve: bad argument");
  cudaStream_t s = (cudaStream_t)stream;
  if (B == 0) {
    if (done) CUDA_TRY(cudaEventRecord(done, s));
    return TFMPC_OK;
  }
  REQ(ws, "tfmpc_ilqr_solve: null workspace");
  IlqrOpts o;
  int rc = make_opts(opts, &o);
  if (rc) return rc;
  if (use_queue(e)) {   // one persistent kernel, wholly in stream order
    rc = queue_ilqr_solve(e, B, T, x0, u_init, o, states, actions, costs, stats, ws, ws_bytes, g_qp.load(), s);
    if (!rc && done) CUDA_TRY(cudaEventRecord(done, s));
    return rc;
  }
  if (use_small(e)) return small_ilqr_solve(e, B, T, x0, u_init, o, states, actions, costs, stats, ws, ws_bytes, s, done);
  // the persistent kernels run wholly in stream order
  rc = e->kind == TFMPC_ENV_NAVLQR ? dense_navlqr_solve(e, B, T, x0, u_init, o, states, actions, costs, stats, ws, ws_bytes, s)
                                   : warp_ilqr_solve(e, B, T, x0, u_init, o, states, actions, costs, stats, ws, ws_bytes, s);
  if (!rc && done) CUDA_TRY(cudaEventRecord(done, s));
  return rc;
}

int tfmpc_ilqr_solve(const tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, const tfmpc_ilqr_opts_t *opts, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, void *stream) {
  return ilqr_solve_impl(e, B, T, x0, u_init, opts, states, actions, costs, stats, ws, ws_bytes, stream, nullptr);
}

int tfmpc_set_graph_mode(int on) { return small_ilqr_graph_mode(on); }

int tfmpc_ilqr_solve_async(const tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, const tfmpc_ilqr_opts_t *opts,
                           real *states, real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, void *stream, void *done_event) {
  REQ(done_event, "tfmpc_ilqr_solve_async: null completion event");
  return ilqr_solve_impl(e, B, T, x0, u_init, opts, states, actions, costs, stats, ws, ws_bytes, stream, (cudaEvent_t)done_event);
}

static int host_scratch(tfmpc_env *e, int64_t bytes) {
  if (e->h_scratch_bytes >= bytes) return TFMPC_OK;
  if (e->h_scratch) { cudaFree(e->h_scratch); e->h_scratch = nullptr; e->h_scratch_bytes = 0; }
  CUDA_TRY(cudaMalloc(&e->h_scratch, (size_t)bytes));
  e->h_scratch_bytes = bytes;
  return TFMPC_OK;
}
static int64_t al(int64_t b) { return (b + 255) / 256 * 256; }

int tfmpc_ilqr_solve_host(tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, const tfmpc_ilqr_opts_t *opts, real *states,
                          real *actions, real *costs, int32_t *stats, void *stream) {
  REQ(e && x0 && u_init && states && actions && costs && stats && B >= 0 && T >= 1, "tfmpc_ilqr_solve_host: bad argument");
  if (B == 0) return TFMPC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = e->n, m = e->m;
  int64_t b_x0 = al(B * n * sizeof(real)), b_u = al(B * T * m * sizeof(real)), b_s = al(B * (T + 1) * n * sizeof(real));
  int64_t b_c = al(B * (T + 1) * sizeof(real)), b_st = al(B * 4 * sizeof(int32_t)), b_ws = tfmpc_ilqr_workspace_bytes(e, B, T);
  if (b_ws < 0) return (int)b_ws;
  int rc = host_scratch(e, b_x0 + 2 * b_u + b_s + b_c + b_st + b_ws);
  if (rc) return rc;
  char *p = (char *)e->h_scratch;
  real *d_x0 = (real *)p; p += b_x0;
  real *d_ui = (real *)p; p += b_u;
  real *d_a = (real *)p; p += b_u;
  real *d_s = (real *)p; p += b_s;
  real *d_c = (real *)p; p += b_c;
  int32_t *d_st = (int32_t *)p; p += b_st;
  void *d_ws = p;
  CUDA_TRY(cudaMemcpyAsync(d_x0, x0, B * n * sizeof(real), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d_ui, u_init, B * T * m * sizeof(real), cudaMemcpyHostToDevice, s));
  rc = tfmpc_ilqr_solve(e, B, T, d_x0, d_ui, opts, d_s, d_a, d_c, d_st, d_ws, b_ws, stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(states, d_s, B * (T + 1) * n * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(actions, d_a, B * T * m * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(costs, d_c, B * (T + 1) * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(stats, d_st, B * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TFMPC_OK;
}

int64_t tfmpc_ilqr_solve_host_scratch_bytes(const tfmpc_env_t *e, int64_t B, int T) {
  if (!e || B < 0 || T < 1) return tfmpc_set_error(TFMPC_E_INVALID, "tfmpc_ilqr_solve_host_scratch_bytes: bad argument");
  const int64_t n = e->n, m = e->m;
  const int64_t ws = tfmpc_ilqr_workspace_bytes(e, B, T);
  if (ws < 0) return ws;
  return al(B * n * sizeof(real)) + 2 * al(B * T * m * sizeof(real)) + al(B * (T + 1) * n * sizeof(real)) + al(B * (T + 1) * sizeof(real)) +
         al(B * 4 * sizeof(int32_t)) + ws;
}

int tfmpc_ilqr_solve_host_async(const tfmpc_env_t *e, int64_t B, int T, const real *x0, const real *u_init, const tfmpc_ilqr_opts_t *opts,
                                real *states, real *actions, real *costs, int32_t *stats, void *scratch, int64_t scratch_bytes, void *stream) {
  REQ(e && x0 && u_init && states && actions && costs && stats && B >= 0 && T >= 1, "tfmpc_ilqr_solve_host_async: bad argument");
  if (B == 0) return TFMPC_OK;
  REQ(scratch, "tfmpc_ilqr_solve_host_async: null scratch");
  const int64_t need = tfmpc_ilqr_solve_host_scratch_bytes(e, B, T);
  if (need < 0) return (int)need;
  if (scratch_bytes < need) return tfmpc_set_error(TFMPC_E_WORKSPACE, "tfmpc_ilqr_solve_host_async: scratch too small (%lld < %lld)", (long long)scratch_bytes, (long long)need);
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t n = e->n, m = e->m;
  const int64_t b_x0 = al(B * n * sizeof(real)), b_u = al(B * T * m * sizeof(real)), b_s = al(B * (T + 1) * n * sizeof(real));
  const int64_t b_c = al(B * (T + 1) * sizeof(real)), b_st = al(B * 4 * sizeof(int32_t));
  char *p = (char *)scratch;
  real *d_x0 = (real *)p; p += b_x0;
  real *d_ui = (real *)p; p += b_u;
  real *d_a = (real *)p; p += b_u;
  real *d_s = (real *)p; p += b_s;
  real *d_c = (real *)p; p += b_c;
  int32_t *d_st = (int32_t *)p; p += b_st;
  void *d_ws = p;
  const int64_t b_ws = scratch_bytes - (p - (char *)scratch);
  CUDA_TRY(cudaMemcpyAsync(d_x0, x0, B * n * sizeof(real), cudaMemcpyHostToDevice, s));
  CUDA_TRY(cudaMemcpyAsync(d_ui, u_init, B * T * m * sizeof(real), cudaMemcpyHostToDevice, s));
  int rc = tfmpc_ilqr_solve(e, B, T, d_x0, d_ui, opts, d_s, d_a, d_c, d_st, d_ws, b_ws, stream);
  if (rc) return rc;
  CUDA_TRY(cudaMemcpyAsync(states, d_s, B * (T + 1) * n * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(actions, d_a, B * T * m * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(costs, d_c, B * (T + 1) * sizeof(real), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaMemcpyAsync(stats, d_st, B * 4 * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
  return TFMPC_OK;
}

// ------------------------------------------------------------------ LQR
int tfmpc_lqr_solve(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC, const real *c,
                    int64_t sc, const real *x0, int terminal_zero, real *states, real *actions, real *costs, real *K, real *k, real *V, real *v,
                    real *cst, int32_t *status, void *stream) {
  REQ(F && f && C && c && x0 && states && actions && costs && B >= 0 && T >= 1 && n >= 1 && m >= 1, "tfmpc_lqr_solve: bad argument");
  REQ(n <= MAXD && m <= MAXD, "tfmpc_lqr_solve: n, m must be <= 32");
  if (B == 0) return TFMPC_OK;
  return lqr_solve_launch(B, n, m, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, K, k, V, v, cst, status,
                          (cudaStream_t)stream);
}

int tfmpc_lqr_forward(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                      const real *c, int64_t sc, const real *K, const real *k, const real *x0, real *states, real *actions, real *costs,
                      void *stream) {
  REQ(F && f && C && c && K && k && x0 && states && actions && costs && B >= 0 && T >= 1 && n >= 1 && m >= 1, "tfmpc_lqr_forward: bad argument");
  REQ(n <= MAXD && m <= MAXD, "tfmpc_lqr_forward: n, m must be <= 32");
  if (B == 0) return TFMPC_OK;
  return lqr_forward_launch(B, n, m, T, F, sF, f, sf, C, sC, c, sc, K, k, x0, states, actions, costs, (cudaStream_t)stream);
}

int tfmpc_lqr_step(int64_t R, int n, int m, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC, const real *c,
                   int64_t sc, const real *x, const real *u, real *x_next, real *cost, real *final_cost, void *stream) {
  REQ(F && f && C && c && x && R >= 0 && n >= 1 && m >= 1, "tfmpc_lqr_step: bad argument");
  REQ(n <= MAXD && m <= MAXD, "tfmpc_lqr_step: n, m must be <= 32");
  REQ(u || (!x_next && !cost), "tfmpc_lqr_step: transition and cost need an action");
  if (R == 0) return TFMPC_OK;
  return lqr_step_launch(R, n, m, F, sF, f, sf, C, sC, c, sc, x, u, x_next, cost, final_cost, (cudaStream_t)stream);
}

int tfmpc_lqr_solve_host(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                         const real *c, int64_t sc, const real *x0, int terminal_zero, real *states, real *actions, real *costs, int32_t *status,
                         void *stream) {
  REQ(F && f && C && c && x0 && states && actions && costs && B >= 0 && T >= 1 && n >= 1 && m >= 1, "tfmpc_lqr_solve_host: bad argument");
  if (B == 0) return TFMPC_OK;
  cudaStream_t s = (cudaStream_t)stream;
  const int64_t N = n + m;
  int64_t eF = (sF ? B : 1) * n * N, ef = (sf ? B : 1) * n, eC = (sC ? B : 1) * N * N, ec = (sc ? B : 1) * N;
  int64_t bF = al(eF * sizeof(real)), bf = al(ef * sizeof(real)), bC = al(eC * sizeof(real)), bc = al(ec * sizeof(real));
  int64_t bx = al(B * n * sizeof(real)), bs = al(B * (T + 1) * n * sizeof(real)), ba = al(B * T * m * sizeof(real)), bco = al(B * (T + 1) * sizeof(real));
  int64_t bst = al(B * sizeof(int32_t));
  char *base = nullptr;
  {  // keep freed stream-ordered memory in the pool across calls (the default threshold of 0 returns it to the OS at every sync)
    static std::once_flag pool_once[64];
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    std::call_once(pool_once[dev & 63], [dev] {
      cudaMemPool_t pool;
      uint64_t keep = ~0ull;
      if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      cudaGetLastError();
    });
  }
  CUDA_TRY(cudaMallocAsync((void **)&base, bF + bf + bC + bc + bx + bs + ba + bco + bst, s));
  char *p = base;
  real *dF = (real *)p; p += bF; real *df = (real *)p; p += bf; real *dC = (real *)p; p += bC; real *dc = (real *)p; p += bc;
  real *dx = (real *)p; p += bx; real *ds = (real *)p; p += bs; real *da = (real *)p; p += ba; real *dco = (real *)p; p += bco;
  int32_t *dst = (int32_t *)p;
  int rc = TFMPC_OK;
#define TRY_FREE(expr) do { cudaError_t _e = (expr); if (_e != cudaSuccess) { rc = tfmpc_set_error(TFMPC_E_CUDA, "%s: %s", #expr, cudaGetErrorString(_e)); goto done; } } while (0)
  TRY_FREE(cudaMemcpyAsync(dF, F, eF * sizeof(real), cudaMemcpyHostToDevice, s));
  TRY_FREE(cudaMemcpyAsync(df, f, ef * sizeof(real), cudaMemcpyHostToDevice, s));
  TRY_FREE(cudaMemcpyAsync(dC, C, eC * sizeof(real), cudaMemcpyHostToDevice, s));
  TRY_FREE(cudaMemcpyAsync(dc, c, ec * sizeof(real), cudaMemcpyHostToDevice, s));
  TRY_FREE(cudaMemcpyAsync(dx, x0, B * n * sizeof(real), cudaMemcpyHostToDevice, s));
  rc = tfmpc_lqr_solve(B, n, m, T, dF, sF, df, sf, dC, sC, dc, sc, dx, terminal_zero, ds, da, dco, nullptr, nullptr, nullptr, nullptr, nullptr, dst, stream);
  if (rc) goto done;
  TRY_FREE(cudaMemcpyAsync(states, ds, B * (T + 1) * n * sizeof(real), cudaMemcpyDeviceToHost, s));
  TRY_FREE(cudaMemcpyAsync(actions, da, B * T * m * sizeof(real), cudaMemcpyDeviceToHost, s));
  TRY_FREE(cudaMemcpyAsync(costs, dco, B * (T + 1) * sizeof(real), cudaMemcpyDeviceToHost, s));
  if (status) TRY_FREE(cudaMemcpyAsync(status, dst, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
done:
  cudaFreeAsync(base, s);
  cudaError_t se = cudaStreamSynchronize(s);
  if (!rc && se != cudaSuccess) rc = tfmpc_set_error(TFMPC_E_CUDA, "stream sync: %s", cudaGetErrorString(se));
  return rc;
}

}  // extern "C"
