/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement ("oracle") of the thiagopbueno/tf-mpc
 * v0.7.0 LQR / iLQR hot path.  Nothing under tfmpc_b200/ links, imports or calls this
 * file: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs use it, and there only as the checker or as the timed CPU baseline.
 *
 * Every function cites the reference lines (relative to /root/reference) it restates.
 * The arithmetic is deliberately the reference's dense, structure-agnostic math: the
 * same matrix products, the same order of accumulation where the reference's Python
 * fixes one, the same branch rules (Appendix A/C of SURVEY.md).  Environment
 * derivatives use the closed forms that the reference's own tests pin
 * (tests/test_env_*.py) instead of autodiff.
 *
 * Pinning: tests/test_oracle_golden.py checks this file against tests/golden/ *.npz,
 * which were produced by executing the unmodified reference Python under the
 * torch-backed TensorFlow API shim in oracle/tf_shim (TensorFlow is not installable
 * in the build container).  What is therefore NOT pinned is Eigen's own rounding.
 *
 * Build (oracle/Makefile): compiled twice, -DREAL=float -> liboracle_f32.so and
 * -DREAL=double -> liboracle_f64.so.  Layouts are the reference's: states [B,T+1,n],
 * actions [B,T,m], costs [B,T+1], row-major.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#ifndef REAL
#define REAL float
#endif

#define MAXD 32 /* max state / action dimension of an environment */
#define MAXZ 8  /* max deceleration zones */

enum { ENV_NAVLQR = 0, ENV_NAVIGATION = 1, ENV_RESERVOIR = 2, ENV_HVAC = 3 };
enum { ST_OK = 0, ST_MAXITER = 1, ST_NONPD = 2, ST_REGLOOP = 3, ST_NAN = 4 };

typedef struct {
  int kind, n, m, nz, bounded;
  REAL low[MAXD], high[MAXD];
  REAL goal[MAXD], beta;
  REAL center[MAXZ][2], decay[MAXZ];
  /* reservoir */
  REAL cap[MAXD], lb[MAXD], ub[MAXD], lowpen[MAXD], highpen[MAXD], sppen[MAXD], rain[MAXD];
  REAL D[MAXD * MAXD];
  /* hvac */
  REAL t_out[MAXD], t_hall[MAXD], r_out[MAXD], r_hall[MAXD], capac[MAXD], air_max[MAXD];
  REAL adj_out[MAXD], adj_hall[MAXD];
  REAL A[MAXD * MAXD]; /* (adj | adj^T) / R_wall */
} env_t;

typedef struct {
  double atol;
  int max_iterations;
  double mu_min, delta_0, c1, alpha_min;
} opts_t;

/* ------------------------------------------------------------------ env construction
 * Packed parameter layout (doubles), shared with include/tfmpc_b200.h:
 *  NAVLQR      goal[n] beta low[n] high[n]                    (+-inf = unbounded)
 *  NAVIGATION  goal[2] low[2] high[2] center[nz][2] decay[nz]
 *  RESERVOIR   max_res_cap lower_bound upper_bound low_penalty high_penalty set_point_penalty
 *              rain_shape rain_scale (each [n]) downstream[n][n]
 *  HVAC        temp_outside temp_hall temp_lower_bound temp_upper_bound R_outside R_hall capacity
 *              air_max adj_outside adj_hall (each [n]) R_wall[n][n] adj[n][n]
 */
void *oracle_env_create(int kind, int n, int m, int nz, const double *p) {
  if (n < 1 || n > MAXD || m < 1 || m > MAXD || nz < 0 || nz > MAXZ) return NULL;
  env_t *e = (env_t *)calloc(1, sizeof(env_t));
  e->kind = kind; e->n = n; e->m = m; e->nz = nz;
  int i, j;
  if (kind == ENV_NAVLQR) { /* envs/lqr/navigation/__init__.py:10-21 */
    for (i = 0; i < n; i++) e->goal[i] = (REAL)p[i];
    e->beta = (REAL)p[n];
    for (i = 0; i < m; i++) { e->low[i] = (REAL)p[n + 1 + i]; e->high[i] = (REAL)p[n + 1 + m + i]; }
  } else if (kind == ENV_NAVIGATION) { /* envs/navigation/__init__.py:11-24,86-96 */
    for (i = 0; i < 2; i++) { e->goal[i] = (REAL)p[i]; e->low[i] = (REAL)p[2 + i]; e->high[i] = (REAL)p[4 + i]; }
    for (i = 0; i < nz; i++) { e->center[i][0] = (REAL)p[6 + 2 * i]; e->center[i][1] = (REAL)p[6 + 2 * i + 1]; }
    for (i = 0; i < nz; i++) e->decay[i] = (REAL)p[6 + 2 * nz + i];
  } else if (kind == ENV_RESERVOIR) { /* envs/reservoir/__init__.py:11-37 */
    for (i = 0; i < n; i++) {
      e->cap[i] = (REAL)p[i]; e->lb[i] = (REAL)p[n + i]; e->ub[i] = (REAL)p[2 * n + i];
      e->lowpen[i] = (REAL)p[3 * n + i]; e->highpen[i] = (REAL)p[4 * n + i]; e->sppen[i] = (REAL)p[5 * n + i];
      /* _rainfall(cec=True) = rain_shape * rain_scale, reservoir/__init__.py:98-100 */
      e->rain[i] = (REAL)p[6 * n + i] * (REAL)p[7 * n + i];
      e->low[i] = (REAL)0; e->high[i] = (REAL)1;
    }
    for (i = 0; i < n * n; i++) e->D[i] = (REAL)p[8 * n + i];
  } else if (kind == ENV_HVAC) { /* envs/hvac/__init__.py:17-58 */
    for (i = 0; i < n; i++) {
      e->t_out[i] = (REAL)p[i]; e->t_hall[i] = (REAL)p[n + i]; e->lb[i] = (REAL)p[2 * n + i]; e->ub[i] = (REAL)p[3 * n + i];
      e->r_out[i] = (REAL)p[4 * n + i]; e->r_hall[i] = (REAL)p[5 * n + i]; e->capac[i] = (REAL)p[6 * n + i];
      e->air_max[i] = (REAL)p[7 * n + i]; e->adj_out[i] = (REAL)p[8 * n + i]; e->adj_hall[i] = (REAL)p[9 * n + i];
      e->low[i] = (REAL)0; e->high[i] = (REAL)1;
    }
    const double *rw = p + 10 * n, *adj = p + 10 * n + n * n;
    for (i = 0; i < n; i++)
      for (j = 0; j < n; j++) { /* hvac/__init__.py:130-133: logical_or(adj, adj^T) / R_wall */
        REAL a = (adj[i * n + j] != 0.0 || adj[j * n + i] != 0.0) ? (REAL)1 : (REAL)0;
        e->A[i * n + j] = a / (REAL)rw[i * n + j];
      }
  } else { free(e); return NULL; }
  /* gym Box.is_bounded(): every low finite and every high finite (ilqr.py:136) */
  e->bounded = 1;
  for (i = 0; i < m; i++) if (isinf((double)e->low[i]) || isinf((double)e->high[i])) e->bounded = 0;
  return e;
}
void oracle_env_destroy(void *e) { free(e); }

/* ------------------------------------------------------------------ env dynamics/costs */
static REAL nav_lambda(const env_t *e, const REAL *x, REAL *lam_z, REAL *r_z) {
  /* navigation/__init__.py:62-74: lambda = prod_z 2/(1+exp(-decay*||x-c||)) - 1 */
  REAL lam = (REAL)1;
  for (int z = 0; z < e->nz; z++) {
    REAL d0 = x[0] - e->center[z][0], d1 = x[1] - e->center[z][1];
    REAL r = (REAL)sqrt((double)(d0 * d0 + d1 * d1));
    REAL l = (REAL)2 / ((REAL)1 + (REAL)exp((double)(-e->decay[z] * r))) - (REAL)1;
    if (lam_z) { lam_z[z] = l; r_z[z] = r; }
    lam *= l;
  }
  return lam;
}

static void env_step(const env_t *e, const REAL *x, const REAL *u, REAL *xn) {
  int n = e->n, i, j;
  switch (e->kind) {
  case ENV_NAVLQR: /* lqr/navigation/__init__.py:30-32 */
    for (i = 0; i < n; i++) xn[i] = x[i] + u[i];
    break;
  case ENV_NAVIGATION: { /* navigation/__init__.py:34-48 (cec=True) */
    REAL lam = nav_lambda(e, x, NULL, NULL);
    for (i = 0; i < 2; i++) xn[i] = x[i] + lam * u[i];
    break; }
  case ENV_RESERVOIR: { /* reservoir/__init__.py:47-62: x + rain + D^T(u*x) - 1/2 sin(x/cap) x - u*x */
    REAL out[MAXD];
    for (i = 0; i < n; i++) out[i] = u[i] * x[i];
    for (i = 0; i < n; i++) {
      REAL inflow = 0;
      for (j = 0; j < n; j++) inflow += e->D[j * n + i] * out[j];
      REAL vap = (REAL)0.5 * (REAL)sin((double)(x[i] / e->cap[i])) * x[i];
      xn[i] = x[i] + e->rain[i] + inflow - vap - out[i];
    }
    break; }
  case ENV_HVAC: { /* hvac/__init__.py:69-91,128-149 */
    for (i = 0; i < n; i++) {
      REAL air = u[i] * e->air_max[i];
      REAL heating = air * (REAL)1.006 * ((REAL)40.0 - x[i]);
      REAL cbr = 0;
      for (j = 0; j < n; j++) cbr += -e->A[i * n + j] * (x[i] - x[j]);
      REAL cwo = e->adj_out[i] / e->r_out[i] * (e->t_out[i] - x[i]);
      REAL cwh = e->adj_hall[i] / e->r_hall[i] * (e->t_hall[i] - x[i]);
      xn[i] = x[i] + (REAL)1.0 / e->capac[i] * (heating + cbr + cwo + cwh);
    }
    break; }
  }
}

static REAL relu(REAL v) { return v > 0 ? v : (REAL)0; }

static REAL env_cost_impl(const env_t *e, const REAL *x, const REAL *u, int final) {
  int n = e->n, i;
  REAL c = 0;
  switch (e->kind) {
  case ENV_NAVLQR: { /* lqr/navigation/__init__.py:34-47 */
    REAL c1 = 0, c2 = 0;
    for (i = 0; i < n; i++) c1 += (x[i] - e->goal[i]) * (x[i] - e->goal[i]);
    if (final) return c1;
    for (i = 0; i < e->m; i++) c2 += u[i] * u[i];
    return c1 + e->beta * c2; }
  case ENV_NAVIGATION: /* navigation/__init__.py:50-60: no action cost */
    for (i = 0; i < 2; i++) c += (x[i] - e->goal[i]) * (x[i] - e->goal[i]);
    return c;
  case ENV_RESERVOIR: /* reservoir/__init__.py:64-84 (final_cost == cost) */
    for (i = 0; i < n; i++) {
      REAL c1 = -e->lowpen[i] * relu(e->lb[i] - x[i]);
      REAL c2 = -e->highpen[i] * relu(x[i] - e->ub[i]);
      REAL c3 = -e->sppen[i] * (REAL)fabs((double)((e->lb[i] + e->ub[i]) / (REAL)2.0 - x[i]));
      c += c1 + c2 + c3;
    }
    return c;
  case ENV_HVAC: /* hvac/__init__.py:93-126 */
    for (i = 0; i < n; i++) {
      REAL oob = (REAL)20000 * (relu(e->lb[i] - x[i]) + relu(x[i] - e->ub[i]));
      REAL sp = (REAL)10.0 * (REAL)fabs((double)((e->lb[i] + e->ub[i]) / (REAL)2 - x[i]));
      if (final) c += oob + sp;
      else c += (REAL)1.0 * (u[i] * e->air_max[i]) + oob + sp;
    }
    return c;
  }
  return c;
}
static REAL env_cost(const env_t *e, const REAL *x, const REAL *u) { return env_cost_impl(e, x, u, 0); }
static REAL env_final_cost(const env_t *e, const REAL *x) { return env_cost_impl(e, x, NULL, 1); }

static REAL sgn(REAL v) { return (REAL)((v > 0) - (v < 0)); }

/* Analytic replacement of DiffEnv.get_linear_transition / get_quadratic_cost
 * (diffenv.py:13-83); closed forms are those asserted by the reference's tests:
 * tests/test_env_navigation.py:62-176, test_env_lqr_navigation.py:28-135,
 * test_env_reservoir.py:151-233, test_env_hvac.py:92-98,170-211 (+ the -diag(A 1) term
 * the commented-out f_x assertion at test_env_hvac.py:90 misses).
 * l_xu is [n,m]; the reference's backward uses l_xu^T (ilqr.py:131). */
static void env_linearize(const env_t *e, const REAL *x, const REAL *u, REAL *f_x, REAL *f_u, REAL *l, REAL *l_x,
                          REAL *l_u, REAL *l_xx, REAL *l_uu, REAL *l_xu) {
  int n = e->n, m = e->m, i, j;
  memset(f_x, 0, sizeof(REAL) * n * n); memset(f_u, 0, sizeof(REAL) * n * m);
  memset(l_xx, 0, sizeof(REAL) * n * n); memset(l_uu, 0, sizeof(REAL) * m * m); memset(l_xu, 0, sizeof(REAL) * n * m);
  *l = env_cost(e, x, u);
  switch (e->kind) {
  case ENV_NAVLQR:
    for (i = 0; i < n; i++) { f_x[i * n + i] = 1; f_u[i * m + i] = 1; l_x[i] = (REAL)2 * (x[i] - e->goal[i]); l_xx[i * n + i] = 2; }
    for (i = 0; i < m; i++) { l_u[i] = (REAL)2 * e->beta * u[i]; l_uu[i * m + i] = (REAL)2 * e->beta; }
    break;
  case ENV_NAVIGATION: {
    REAL lam_z[MAXZ], r_z[MAXZ], g[2] = {0, 0};
    REAL lam = nav_lambda(e, x, lam_z, r_z);
    for (int z = 0; z < e->nz; z++) {
      /* d lambda_z / d r = 2 d e^{-d r} / (1 + e^{-d r})^2 */
      REAL ex = (REAL)exp((double)(-e->decay[z] * r_z[z]));
      REAL h = (REAL)2 * e->decay[z] * ex / (((REAL)1 + ex) * ((REAL)1 + ex));
      REAL others = 1;
      for (int y = 0; y < e->nz; y++) if (y != z) others *= lam_z[y];
      g[0] += h * (x[0] - e->center[z][0]) / r_z[z] * others;
      g[1] += h * (x[1] - e->center[z][1]) / r_z[z] * others;
    }
    for (i = 0; i < 2; i++) {
      for (j = 0; j < 2; j++) f_x[i * 2 + j] = (i == j ? (REAL)1 : (REAL)0) + u[i] * g[j];
      f_u[i * 2 + i] = lam;
      l_x[i] = (REAL)2 * (x[i] - e->goal[i]); l_u[i] = 0; l_xx[i * 2 + i] = 2;
    }
    break; }
  case ENV_RESERVOIR:
    for (i = 0; i < n; i++) {
      REAL a = x[i] / e->cap[i];
      REAL dvap = (REAL)0.5 * ((REAL)cos((double)a) * a + (REAL)sin((double)a));
      for (j = 0; j < n; j++) { /* D^T diag(u) and D^T diag(x) */
        f_x[i * n + j] = e->D[j * n + i] * u[j];
        f_u[i * n + j] = e->D[j * n + i] * x[j];
      }
      f_x[i * n + i] += (REAL)1 - dvap - u[i];
      f_u[i * n + i] += -x[i];
      REAL mid = (e->lb[i] + e->ub[i]) / (REAL)2.0;
      l_x[i] = e->lowpen[i] * (REAL)(e->lb[i] - x[i] > 0) - e->highpen[i] * (REAL)(x[i] - e->ub[i] > 0) +
               e->sppen[i] * sgn(mid - x[i]);
      l_u[i] = 0;
    }
    break;
  case ENV_HVAC:
    for (i = 0; i < n; i++) {
      REAL s = (REAL)1.0 / e->capac[i], rowsum = 0;
      for (j = 0; j < n; j++) { f_x[i * n + j] = s * e->A[i * n + j]; rowsum += e->A[i * n + j]; }
      f_x[i * n + i] += (REAL)1 + s * (-(u[i] * e->air_max[i]) * (REAL)1.006 - rowsum - e->adj_out[i] / e->r_out[i] -
                                       e->adj_hall[i] / e->r_hall[i]);
      f_u[i * n + i] = s * (e->air_max[i] * (REAL)1.006 * ((REAL)40.0 - x[i]));
      REAL mid = (e->lb[i] + e->ub[i]) / (REAL)2;
      l_x[i] = (REAL)20000 * ((REAL)(x[i] - e->ub[i] > 0) - (REAL)(e->lb[i] - x[i] > 0)) - (REAL)10.0 * sgn(mid - x[i]);
      l_u[i] = e->air_max[i];
    }
    break;
  }
}

/* diffenv.py:85-101 */
static void env_final_quad(const env_t *e, const REAL *x, REAL *l, REAL *l_x, REAL *l_xx) {
  int n = e->n, i;
  memset(l_xx, 0, sizeof(REAL) * n * n);
  *l = env_final_cost(e, x);
  switch (e->kind) {
  case ENV_NAVLQR:
  case ENV_NAVIGATION:
    for (i = 0; i < n; i++) { l_x[i] = (REAL)2 * (x[i] - e->goal[i]); l_xx[i * n + i] = 2; }
    break;
  case ENV_RESERVOIR:
    for (i = 0; i < n; i++) {
      REAL mid = (e->lb[i] + e->ub[i]) / (REAL)2.0;
      l_x[i] = e->lowpen[i] * (REAL)(e->lb[i] - x[i] > 0) - e->highpen[i] * (REAL)(x[i] - e->ub[i] > 0) +
               e->sppen[i] * sgn(mid - x[i]);
    }
    break;
  case ENV_HVAC:
    for (i = 0; i < n; i++) {
      REAL mid = (e->lb[i] + e->ub[i]) / (REAL)2;
      l_x[i] = (REAL)20000 * ((REAL)(x[i] - e->ub[i] > 0) - (REAL)(e->lb[i] - x[i] > 0)) - (REAL)10.0 * sgn(mid - x[i]);
    }
    break;
  }
}

/* batched env evaluation for the fixtures: x[B,n] u[B,m] */
void oracle_env_eval(const void *env, int B, const REAL *x, const REAL *u, REAL *next, REAL *cost, REAL *final_cost) {
  const env_t *e = (const env_t *)env;
  for (int b = 0; b < B; b++) {
    env_step(e, x + b * e->n, u + b * e->m, next + b * e->n);
    cost[b] = env_cost(e, x + b * e->n, u + b * e->m);
    final_cost[b] = env_final_cost(e, x + b * e->n);
  }
}
void oracle_env_linearize(const void *env, int B, const REAL *x, const REAL *u, REAL *f_x, REAL *f_u, REAL *l, REAL *l_x,
                          REAL *l_u, REAL *l_xx, REAL *l_uu, REAL *l_xu, REAL *fl, REAL *fl_x, REAL *fl_xx) {
  const env_t *e = (const env_t *)env;
  int n = e->n, m = e->m;
  for (int b = 0; b < B; b++) {
    env_linearize(e, x + b * n, u + b * m, f_x + b * n * n, f_u + b * n * m, l + b, l_x + b * n, l_u + b * m,
                  l_xx + b * n * n, l_uu + b * m * m, l_xu + b * n * m);
    env_final_quad(e, x + b * n, fl + b, fl_x + b * n, fl_xx + b * n * n);
  }
}

/* ------------------------------------------------------------------ small dense helpers */
/* lower Cholesky in place on a d x d matrix with leading dimension ld; returns 0 ok, 1 not PD
 * (Eigen LLT / LAPACK potrf rule: a non-positive or NaN pivot fails). */
static int chol(REAL *a, int d, int ld) {
  for (int j = 0; j < d; j++) {
    REAL s = a[j * ld + j];
    for (int k = 0; k < j; k++) s -= a[j * ld + k] * a[j * ld + k];
    if (!(s > 0)) return 1;
    REAL dj = (REAL)sqrt((double)s);
    a[j * ld + j] = dj;
    for (int i = j + 1; i < d; i++) {
      REAL t = a[i * ld + j];
      for (int k = 0; k < j; k++) t -= a[i * ld + k] * a[j * ld + k];
      a[i * ld + j] = t / dj;
    }
  }
  return 0;
}
/* solve L L^T y = b in place for one right-hand side with stride */
static void chol_solve(const REAL *L, int d, int ld, REAL *b, int stride) {
  for (int i = 0; i < d; i++) {
    REAL s = b[i * stride];
    for (int k = 0; k < i; k++) s -= L[i * ld + k] * b[k * stride];
    b[i * stride] = s / L[i * ld + i];
  }
  for (int i = d - 1; i >= 0; i--) {
    REAL s = b[i * stride];
    for (int k = i + 1; k < d; k++) s -= L[k * ld + i] * b[k * stride];
    b[i * stride] = s / L[i * ld + i];
  }
}

/* ------------------------------------------------------------------ box-QP
 * utils/optimization.py:6-101 (+ _get_qp_indices :121-127).  x is in/out.  Hfree receives the
 * Cholesky factor of H[free,free] (nfree x nfree, leading dimension m).  Returns 0, or 2 when
 * a factorisation failed (the reference logs and breaks, optimization.py:47-51).
 */
static REAL qp_value(int m, const REAL *H, const REAL *q, const REAL *x) {
  REAL quad = 0, lin = 0; /* f(x) = 1/2 x^T (H x) + q^T x, optimization.py:8-11 */
  for (int i = 0; i < m; i++) {
    REAL hx = 0;
    for (int j = 0; j < m; j++) hx += H[i * m + j] * x[j];
    quad += x[i] * hx;
    lin += q[i] * x[i];
  }
  return (REAL)0.5 * quad + lin;
}

static int boxqp(int m, const REAL *H, const REAL *q, const REAL *lo, const REAL *hi, REAL *x, REAL *Hfree, int *isfree,
                 int *nfree_out, int *iters_out) {
  const REAL rtol = (REAL)1e-8, armijo = (REAL)0.1, eps = (REAL)1e-6;
  const double step_dec = 0.6, min_step = 1e-22;
  int clamped[MAXD], old_clamped[MAXD], idx[MAXD];
  REAL g[MAXD], search[MAXD], xc[MAXD], gc[MAXD];
  int i, j, nfree = 0, status = 0, it;
  for (i = 0; i < m; i++) { clamped[i] = 0; isfree[i] = 1; }
  REAL value = qp_value(m, H, q, x), old_value = 0;
  for (it = 0; it < 100; it++) {
    if (it > 0 && (old_value - value) < rtol * (REAL)fabs((double)old_value)) break; /* :27 */
    old_value = value;
    for (i = 0; i < m; i++) old_clamped[i] = clamped[i];
    for (i = 0; i < m; i++) { /* g = q + H x, :34 */
      REAL s = 0;
      for (j = 0; j < m; j++) s += H[i * m + j] * x[j];
      g[i] = q[i] + s;
    }
    int changed = 0, allc = 1;
    for (i = 0; i < m; i++) { /* :121-127 */
      clamped[i] = ((REAL)fabs((double)(x[i] - lo[i])) < eps && g[i] > 0) || ((REAL)fabs((double)(hi[i] - x[i])) < eps && g[i] < 0);
      isfree[i] = !clamped[i];
      if (clamped[i] != old_clamped[i]) changed = 1;
      if (!clamped[i]) allc = 0;
    }
    if (it == 0 || changed) { /* :37-51 */
      nfree = 0;
      for (i = 0; i < m; i++) if (isfree[i]) idx[nfree++] = i;
      for (i = 0; i < nfree; i++) for (j = 0; j < nfree; j++) Hfree[i * m + j] = H[idx[i] * m + idx[j]];
      if (chol(Hfree, nfree, m)) { status = 2; break; }
    }
    if (allc) break; /* :53 */
    REAL gn = 0;
    for (i = 0; i < m; i++) if (isfree[i]) gn += g[i] * g[i];
    if ((REAL)sqrt((double)gn) < eps) break; /* :58-62 */
    for (i = 0; i < m; i++) { /* grad_clamped = q + H (x * clamped), :65 */
      REAL s = 0;
      for (j = 0; j < m; j++) s += H[i * m + j] * (clamped[j] ? x[j] : (REAL)0);
      gc[i] = q[i] + s;
    }
    REAL rhs[MAXD];
    for (i = 0; i < nfree; i++) rhs[i] = gc[idx[i]];
    chol_solve(Hfree, nfree, m, rhs, 1);
    for (i = 0; i < m; i++) search[i] = 0;
    for (i = 0; i < nfree; i++) search[idx[i]] = -rhs[i] - x[idx[i]]; /* :70 */
    REAL sdotg = 0;
    for (i = 0; i < m; i++) sdotg += search[i] * g[i];
    if (sdotg >= 0) break; /* :75-79 */
    double step = 1.0;
    REAL vc;
    for (;;) { /* :82-95 */
      REAL st = (REAL)step;
      for (i = 0; i < m; i++) {
        REAL v = x[i] + st * search[i];
        xc[i] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
      }
      vc = qp_value(m, H, q, xc);
      if (!((vc - old_value) / (st * sdotg) < armijo)) break;
      step *= step_dec;
      if (step < min_step) { /* the reference recomputes xc, vc once more before breaking */
        st = (REAL)step;
        for (i = 0; i < m; i++) {
          REAL v = x[i] + st * search[i];
          xc[i] = v < lo[i] ? lo[i] : (v > hi[i] ? hi[i] : v);
        }
        vc = qp_value(m, H, q, xc);
        break;
      }
    }
    for (i = 0; i < m; i++) x[i] = xc[i];
    value = vc;
  }
  *nfree_out = nfree;
  if (iters_out) *iters_out = it;
  return status;
}

void oracle_boxqp(int B, int m, const REAL *H, const REAL *q, const REAL *lo, const REAL *hi, REAL *x, REAL *Hfree,
                  int32_t *isfree, int32_t *nfree, int32_t *status) {
  for (int b = 0; b < B; b++) {
    int fr[MAXD], nf = 0;
    status[b] = boxqp(m, H + b * m * m, q + b * m, lo + b * m, hi + b * m, x + b * m, Hfree + b * m * m, fr, &nf, NULL);
    for (int i = 0; i < m; i++) isfree[b * m + i] = fr[i];
    nfree[b] = nf;
  }
}

/* ------------------------------------------------------------------ iLQR stages */
typedef struct { /* linearisation along a trajectory (derivatives(), ilqr.py:84-92) */
  REAL *f_x, *f_u, *l, *l_x, *l_u, *l_xx, *l_uu, *l_xu, fl, *fl_x, *fl_xx;
} lin_t;

static void lin_alloc(lin_t *L, int T, int n, int m) {
  L->f_x = (REAL *)malloc(sizeof(REAL) * T * n * n); L->f_u = (REAL *)malloc(sizeof(REAL) * T * n * m);
  L->l = (REAL *)malloc(sizeof(REAL) * T); L->l_x = (REAL *)malloc(sizeof(REAL) * T * n); L->l_u = (REAL *)malloc(sizeof(REAL) * T * m);
  L->l_xx = (REAL *)malloc(sizeof(REAL) * T * n * n); L->l_uu = (REAL *)malloc(sizeof(REAL) * T * m * m);
  L->l_xu = (REAL *)malloc(sizeof(REAL) * T * n * m); L->fl_x = (REAL *)malloc(sizeof(REAL) * n); L->fl_xx = (REAL *)malloc(sizeof(REAL) * n * n);
}
static void lin_free(lin_t *L) {
  free(L->f_x); free(L->f_u); free(L->l); free(L->l_x); free(L->l_u); free(L->l_xx); free(L->l_uu); free(L->l_xu); free(L->fl_x); free(L->fl_xx);
}
static void derivatives(const env_t *e, int T, const REAL *xh, const REAL *uh, lin_t *L) {
  int n = e->n, m = e->m;
  for (int t = 0; t < T; t++)
    env_linearize(e, xh + t * n, uh + t * m, L->f_x + t * n * n, L->f_u + t * n * m, L->l + t, L->l_x + t * n, L->l_u + t * m,
                  L->l_xx + t * n * n, L->l_uu + t * m * m, L->l_xu + t * n * m);
  env_final_quad(e, xh + T * n, &L->fl, L->fl_x, L->fl_xx);
}

/* C[r,c] = A^T[r,k] B[k,c]  where A is stored [k,r] */
static void matTmul(const REAL *A, const REAL *Bm, REAL *C, int r, int k, int c) {
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      REAL s = 0;
      for (int p = 0; p < k; p++) s += A[p * r + i] * Bm[p * c + j];
      C[i * c + j] = s;
    }
}
static void matmul(const REAL *A, const REAL *Bm, REAL *C, int r, int k, int c) {
  for (int i = 0; i < r; i++)
    for (int j = 0; j < c; j++) {
      REAL s = 0;
      for (int p = 0; p < k; p++) s += A[i * k + p] * Bm[p * c + j];
      C[i * c + j] = s;
    }
}

/* iLQR.backward, ilqr.py:94-172, with the controllers of :357-387.  Returns 0, 1 if the
 * unconstrained Cholesky failed (the caller retries, ilqr.py:305-309), 2 if a box-QP
 * factorisation failed. */
static int ilqr_backward(const env_t *e, int T, const REAL *uh, const lin_t *L, REAL mu, REAL *K, REAL *k, REAL *J_out,
                         REAL *dV1_out, REAL *dV2_out, int *branch_counts) {
  int n = e->n, m = e->m, i, j, t, status = 0;
  REAL V_x[MAXD], V_xx[MAXD * MAXD], Vreg[MAXD * MAXD];
  REAL Q_x[MAXD], Q_u[MAXD], Q_xx[MAXD * MAXD], Q_uu[MAXD * MAXD], Q_ux[MAXD * MAXD], Q_uu_reg[MAXD * MAXD], Q_ux_reg[MAXD * MAXD];
  REAL fxTV[MAXD * MAXD], fuTV[MAXD * MAXD], fuTVr[MAXD * MAXD], tmp[MAXD * MAXD], KtQuu[MAXD * MAXD];
  memcpy(V_x, L->fl_x, sizeof(REAL) * n); memcpy(V_xx, L->fl_xx, sizeof(REAL) * n * n);
  REAL J = L->fl, dV1 = 0, dV2 = 0;
  for (t = T - 1; t >= 0; t--) {
    const REAL *f_x = L->f_x + t * n * n, *f_u = L->f_u + t * n * m, *l_x = L->l_x + t * n, *l_u = L->l_u + t * m;
    const REAL *l_xx = L->l_xx + t * n * n, *l_uu = L->l_uu + t * m * m, *l_xu = L->l_xu + t * n * m;
    REAL *K_t = K + t * m * n, *k_t = k + t * m;
    matTmul(f_x, V_x, Q_x, n, n, 1); for (i = 0; i < n; i++) Q_x[i] = l_x[i] + Q_x[i];     /* :122 */
    matTmul(f_u, V_x, Q_u, m, n, 1); for (i = 0; i < m; i++) Q_u[i] = l_u[i] + Q_u[i];     /* :123 */
    matTmul(f_x, V_xx, fxTV, n, n, n);                                                      /* :125 */
    matTmul(f_u, V_xx, fuTV, m, n, n);                                                      /* :126 */
    for (i = 0; i < n * n; i++) Vreg[i] = V_xx[i];
    for (i = 0; i < n; i++) Vreg[i * n + i] = V_xx[i * n + i] + mu * (REAL)1;
    matTmul(f_u, Vreg, fuTVr, m, n, n);                                                     /* :127 */
    matmul(fxTV, f_x, Q_xx, n, n, n); for (i = 0; i < n * n; i++) Q_xx[i] = l_xx[i] + Q_xx[i]; /* :129 */
    matmul(fuTV, f_u, Q_uu, m, n, m); for (i = 0; i < m * m; i++) Q_uu[i] = l_uu[i] + Q_uu[i]; /* :130 */
    matmul(fuTV, f_x, Q_ux, m, n, n);
    for (i = 0; i < m; i++) for (j = 0; j < n; j++) Q_ux[i * n + j] = l_xu[j * m + i] + Q_ux[i * n + j]; /* :131 */
    matmul(fuTVr, f_u, Q_uu_reg, m, n, m); for (i = 0; i < m * m; i++) Q_uu_reg[i] = l_uu[i] + Q_uu_reg[i]; /* :133 */
    matmul(fuTVr, f_x, Q_ux_reg, m, n, n);
    for (i = 0; i < m; i++) for (j = 0; j < n; j++) Q_ux_reg[i * n + j] = l_xu[j * m + i] + Q_ux_reg[i * n + j]; /* :134 */

    if (e->bounded) { /* :136 */
      int nz = 0;
      for (i = 0; i < n * n; i++) if (V_xx[i] != 0) nz++;
      if (nz > 0) { /* :137-138 -> _get_constrained_controller :364-387 */
        REAL lo[MAXD], hi[MAXD], Hf[MAXD * MAXD];
        int fr[MAXD], nf = 0, idx[MAXD];
        for (i = 0; i < m; i++) { lo[i] = e->low[i] - uh[t * m + i]; hi[i] = e->high[i] - uh[t * m + i]; k_t[i] = (lo[i] + hi[i]) / (REAL)2; }
        int st = boxqp(m, Q_uu_reg, Q_u, lo, hi, k_t, Hf, fr, &nf, NULL);
        if (st) status = 2;
        for (i = 0; i < m * n; i++) K_t[i] = 0;
        int c = 0;
        for (i = 0; i < m; i++) if (fr[i]) idx[c++] = i;
        if (c > 0 && c == nf && !st) { /* K[free] = -cholesky_solve(Hfree, Q_ux_reg[free]) :375-383 */
          for (j = 0; j < n; j++) {
            REAL col[MAXD];
            for (i = 0; i < c; i++) col[i] = Q_ux_reg[idx[i] * n + j];
            chol_solve(Hf, c, m, col, 1);
            for (i = 0; i < c; i++) K_t[idx[i] * n + j] = -col[i];
          }
        }
        if (branch_counts) branch_counts[0]++;
      } else { /* :139-141 bang-bang */
        for (i = 0; i < m * n; i++) K_t[i] = 0;
        for (i = 0; i < m; i++) k_t[i] = (Q_u[i] >= 0) ? e->low[i] - uh[t * m + i] : e->high[i] - uh[t * m + i];
        if (branch_counts) branch_counts[1]++;
      }
    } else { /* :143 -> _get_unconstrained_controller :357-362 */
      REAL R[MAXD * MAXD];
      for (i = 0; i < m * m; i++) R[i] = Q_uu_reg[i];
      if (chol(R, m, m)) return 1;
      for (i = 0; i < m; i++) k_t[i] = Q_u[i];
      chol_solve(R, m, m, k_t, 1);
      for (i = 0; i < m; i++) k_t[i] = -k_t[i];
      for (j = 0; j < n; j++) {
        REAL col[MAXD];
        for (i = 0; i < m; i++) col[i] = Q_ux_reg[i * n + j];
        chol_solve(R, m, m, col, 1);
        for (i = 0; i < m; i++) K_t[i * n + j] = -col[i];
      }
      if (branch_counts) branch_counts[2]++;
    }
    /* value update with the UNregularised Q, :145-162 */
    matTmul(K_t, Q_uu, KtQuu, n, m, m); /* K^T Q_uu  [n,m] */
    REAL a1[MAXD], a2[MAXD], a3[MAXD];
    matTmul(Q_ux, k_t, a1, n, m, 1); matTmul(K_t, Q_u, a2, n, m, 1); matmul(KtQuu, k_t, a3, n, m, 1);
    for (i = 0; i < n; i++) V_x[i] = Q_x[i] + a1[i] + a2[i] + a3[i];
    REAL b1[MAXD * MAXD], b2[MAXD * MAXD], b3[MAXD * MAXD];
    matTmul(Q_ux, K_t, b1, n, m, n); matTmul(K_t, Q_ux, b2, n, m, n); matmul(KtQuu, K_t, b3, n, m, n);
    for (i = 0; i < n * n; i++) tmp[i] = Q_xx[i] + b1[i] + b2[i] + b3[i];
    for (i = 0; i < n; i++) for (j = 0; j < n; j++) V_xx[i * n + j] = (REAL)0.5 * (tmp[i * n + j] + tmp[j * n + i]); /* :162 */
    J += L->l[t];                                                                                   /* :164 */
    REAL d1 = 0; for (i = 0; i < m; i++) d1 += k_t[i] * Q_u[i];
    dV1 += d1;                                                                                      /* :166 */
    REAL kq[MAXD], d2 = 0;
    for (j = 0; j < m; j++) { REAL s = 0; for (i = 0; i < m; i++) s += k_t[i] * Q_uu[i * m + j]; kq[j] = s; }
    for (j = 0; j < m; j++) d2 += kq[j] * k_t[j];
    dV2 += (REAL)0.5 * d2;                                                                          /* :167 */
  }
  *J_out = J; *dV1_out = dV1; *dV2_out = dV2;
  return status;
}

/* iLQR.forward, ilqr.py:174-212 */
static void ilqr_forward(const env_t *e, int T, const REAL *xh, const REAL *uh, const REAL *K, const REAL *k, REAL alpha, REAL *xs,
                         REAL *us, REAL *cs, REAL *J_out, REAL *res_out) {
  int n = e->n, m = e->m, i, j, t;
  REAL J = 0, residual = 0;
  memcpy(xs, xh, sizeof(REAL) * n);
  for (t = 0; t < T; t++) {
    const REAL *x = xs + t * n;
    REAL *u = us + t * m;
    for (i = 0; i < m; i++) {
      REAL s = 0;
      for (j = 0; j < n; j++) s += K[t * m * n + i * n + j] * (x[j] - xh[t * n + j]);
      REAL du = alpha * k[t * m + i] + s;                 /* :194 */
      REAL a = uh[t * m + i] + du;
      u[i] = a < e->low[i] ? e->low[i] : (a > e->high[i] ? e->high[i] : a); /* :197 */
      REAL ad = (REAL)fabs((double)du);
      if (ad > residual) residual = ad;                   /* :206 (pre-clip) */
    }
    cs[t] = env_cost(e, x, u);
    env_step(e, x, u, xs + (t + 1) * n);
    J += cs[t];
  }
  cs[T] = env_final_cost(e, xs + T * n);
  J += cs[T];
  *J_out = J; *res_out = residual;
}

/* iLQR.start with the random actions pinned to u_init (ilqr.py:53-82) */
static void ilqr_start(const env_t *e, int T, const REAL *x0, const REAL *u_init, REAL *xs, REAL *us, REAL *cs) {
  int n = e->n, m = e->m;
  memcpy(xs, x0, sizeof(REAL) * n);
  memcpy(us, u_init, sizeof(REAL) * T * m);
  for (int t = 0; t < T; t++) {
    cs[t] = env_cost(e, xs + t * n, us + t * m);
    env_step(e, xs + t * n, us + t * m, xs + (t + 1) * n);
  }
  cs[T] = env_final_cost(e, xs + T * n);
}

static void line_search_alphas(double alpha_min, double *a) {
  /* np.geomspace(1.0, alpha_min, 11), ilqr.py:322; the default table is numpy's own output */
  static const double dflt[11] = {1.0, 0.5011872336272722, 0.251188643150958, 0.12589254117941676, 0.06309573444801933,
                                  0.03162277660168379, 0.01584893192461114, 0.007943282347242814, 0.003981071705534973,
                                  0.0019952623149688807, 0.001};
  if (alpha_min == 1e-3) { memcpy(a, dflt, sizeof(dflt)); return; }
  double l = log10(alpha_min);
  for (int i = 0; i < 11; i++) a[i] = pow(10.0, l * i / 10.0);
  a[0] = 1.0; a[10] = alpha_min;
}

/* iLQR.solve, ilqr.py:214-283 with _backward :285-315 and _forward :317-355.
 * stats = {iteration index (reference return value), backward passes, rollouts, status}. */
static void ilqr_solve(const env_t *e, const opts_t *o, int T, const REAL *x0, const REAL *u_init, REAL *states, REAL *actions,
                       REAL *costs, int32_t *stats) {
  int n = e->n, m = e->m, i, t;
  size_t sx = (size_t)(T + 1) * n, su = (size_t)T * m;
  REAL *xh = states, *uh = actions, *ch = costs;
  REAL *x = (REAL *)malloc(sizeof(REAL) * sx), *u = (REAL *)malloc(sizeof(REAL) * su), *c = (REAL *)malloc(sizeof(REAL) * (T + 1));
  REAL *K = (REAL *)malloc(sizeof(REAL) * su * n), *k = (REAL *)malloc(sizeof(REAL) * su);
  lin_t L; lin_alloc(&L, T, n, m);
  double alphas[11]; line_search_alphas(o->alpha_min, alphas);
  double mu = 0.0, delta = 1.0; /* python floats, ilqr.py:215-216 */
  int iteration = 0, n_bwd = 0, n_fwd = 0, status = ST_MAXITER;
  const REAL atol = (REAL)o->atol;
  ilqr_start(e, T, x0, u_init, xh, uh, ch);
  for (iteration = 0; iteration < o->max_iterations; iteration++) {
    derivatives(e, T, xh, uh, &L);
    int converged = 0, guard = 0;
    for (;;) {
      REAL J_hat, dV1, dV2;
      /* _backward: retry with a larger *local* mu while the Cholesky fails (:292-309) */
      double mu_l = mu, delta_l = delta;
      int bst, tries = 0;
      for (;;) {
        bst = ilqr_backward(e, T, uh, &L, (REAL)mu_l, K, k, &J_hat, &dV1, &dV2, NULL);
        n_bwd++;
        if (bst != 1 || ++tries > 200) break;
        delta_l = fmax(o->delta_0, delta_l * o->delta_0);
        mu_l = fmax(o->mu_min, mu_l * delta_l);
      }
      if (bst) { status = ST_NONPD; converged = 1; break; }
      REAL g = 0; /* :243 */
      for (t = 0; t < T; t++) {
        REAL mx = 0;
        for (i = 0; i < m; i++) {
          REAL v = (REAL)fabs((double)k[t * m + i]) / ((REAL)fabs((double)uh[t * m + i]) + (REAL)1.0);
          if (i == 0 || v > mx) mx = v;
        }
        g += mx;
      }
      g = g / (REAL)T;
      if (!(g == g)) { status = ST_NAN; converged = 1; break; }
      if (g < atol) { converged = 1; status = ST_OK; break; } /* :245-248 */
      /* _forward :317-355 */
      int accept = 0;
      REAL residual = 0;
      for (int ai = 0; ai < 11; ai++) {
        REAL alpha = (REAL)alphas[ai], J;
        ilqr_forward(e, T, xh, uh, K, k, alpha, x, u, c, &J, &residual);
        n_fwd++;
        REAL delta_J = -alpha * (dV1 + alpha * dV2); /* :339 */
        REAL dcost = J_hat - J, z;
        if (delta_J > 0) z = dcost / delta_J; else z = sgn(dcost);
        if (z >= (REAL)o->c1) { accept = 1; break; }   /* :351 */
      }
      if (residual < atol) { /* :253-257: candidate taken even if rejected */
        converged = 1; status = ST_OK;
        memcpy(xh, x, sizeof(REAL) * sx); memcpy(uh, u, sizeof(REAL) * su); memcpy(ch, c, sizeof(REAL) * (T + 1));
        break;
      }
      if (accept) { /* :259-266 */
        delta = fmin(1.0 / o->delta_0, delta / o->delta_0);
        mu = mu * delta * (double)(mu * delta > o->mu_min);
        memcpy(xh, x, sizeof(REAL) * sx); memcpy(uh, u, sizeof(REAL) * su); memcpy(ch, c, sizeof(REAL) * (T + 1));
        break;
      } else { /* :267-270 */
        delta = fmax(o->delta_0, delta * o->delta_0);
        mu = fmax(o->mu_min, mu * delta);
      }
      if (++guard > 200) { status = ST_REGLOOP; converged = 1; break; }
    }
    if (converged) break;
  }
  if (iteration >= o->max_iterations) iteration = o->max_iterations - 1; /* python's loop variable after exhaustion */
  stats[0] = iteration; stats[1] = n_bwd; stats[2] = n_fwd; stats[3] = status;
  lin_free(&L); free(x); free(u); free(c); free(K); free(k);
}

/* ------------------------------------------------------------------ exported stage / solve entry points */
void oracle_ilqr_start(const void *env, int B, int T, const REAL *x0, const REAL *u_init, REAL *states, REAL *actions, REAL *costs) {
  const env_t *e = (const env_t *)env;
  for (int b = 0; b < B; b++)
    ilqr_start(e, T, x0 + b * e->n, u_init + (size_t)b * T * e->m, states + (size_t)b * (T + 1) * e->n,
               actions + (size_t)b * T * e->m, costs + (size_t)b * (T + 1));
}

/* returns per-problem status (0 ok / 1 cholesky failure / 2 box-QP failure); branch[b*3..] counts
 * box-QP / bang-bang / cholesky timesteps */
void oracle_ilqr_backward(const void *env, int B, int T, const REAL *states, const REAL *actions, double mu, REAL *K, REAL *k,
                          REAL *J, REAL *dV1, REAL *dV2, int32_t *status, int32_t *branch) {
  const env_t *e = (const env_t *)env;
  int n = e->n, m = e->m;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++) {
    lin_t L; lin_alloc(&L, T, n, m);
    derivatives(e, T, states + (size_t)b * (T + 1) * n, actions + (size_t)b * T * m, &L);
    int bc[3] = {0, 0, 0};
    status[b] = ilqr_backward(e, T, actions + (size_t)b * T * m, &L, (REAL)mu, K + (size_t)b * T * m * n, k + (size_t)b * T * m,
                              J + b, dV1 + b, dV2 + b, bc);
    if (branch) { branch[b * 3] = bc[0]; branch[b * 3 + 1] = bc[1]; branch[b * 3 + 2] = bc[2]; }
    lin_free(&L);
  }
}

void oracle_ilqr_forward(const void *env, int B, int T, const REAL *states, const REAL *actions, const REAL *K, const REAL *k,
                         double alpha, REAL *xs, REAL *us, REAL *cs, REAL *J, REAL *residual) {
  const env_t *e = (const env_t *)env;
  int n = e->n, m = e->m;
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++)
    ilqr_forward(e, T, states + (size_t)b * (T + 1) * n, actions + (size_t)b * T * m, K + (size_t)b * T * m * n, k + (size_t)b * T * m,
                 (REAL)alpha, xs + (size_t)b * (T + 1) * n, us + (size_t)b * T * m, cs + (size_t)b * (T + 1), J + b, residual + b);
}

void oracle_ilqr_solve(const void *env, int B, int T, const REAL *x0, const REAL *u_init, double atol, int max_iterations,
                       double mu_min, double delta_0, double c1, double alpha_min, REAL *states, REAL *actions, REAL *costs,
                       int32_t *stats, int nthreads) {
  const env_t *e = (const env_t *)env;
  opts_t o = {atol, max_iterations, mu_min, delta_0, c1, alpha_min};
  int n = e->n, m = e->m;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(dynamic)
  for (int b = 0; b < B; b++)
    ilqr_solve(e, &o, T, x0 + (size_t)b * n, u_init + (size_t)b * T * m, states + (size_t)b * (T + 1) * n,
               actions + (size_t)b * T * m, costs + (size_t)b * (T + 1), stats + b * 4);
}

int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* ------------------------------------------------------------------ LQR (lqr.py:59-166) */
/* general inverse by Gauss-Jordan with partial pivoting (tf.linalg.inv is LU-based, lqr.py:84) */
static int inverse(const REAL *A, REAL *inv, int d) {
  REAL w[2 * 64 * 64];
  int i, j, k;
  for (i = 0; i < d; i++) for (j = 0; j < d; j++) { w[i * 2 * d + j] = A[i * d + j]; w[i * 2 * d + d + j] = (i == j) ? (REAL)1 : (REAL)0; }
  for (k = 0; k < d; k++) {
    int p = k; REAL best = (REAL)fabs((double)w[k * 2 * d + k]);
    for (i = k + 1; i < d; i++) { REAL v = (REAL)fabs((double)w[i * 2 * d + k]); if (v > best) { best = v; p = i; } }
    if (best == 0) return 1;
    if (p != k) for (j = 0; j < 2 * d; j++) { REAL tv = w[k * 2 * d + j]; w[k * 2 * d + j] = w[p * 2 * d + j]; w[p * 2 * d + j] = tv; }
    REAL piv = w[k * 2 * d + k];
    for (j = 0; j < 2 * d; j++) w[k * 2 * d + j] /= piv;
    for (i = 0; i < d; i++) if (i != k) {
      REAL fct = w[i * 2 * d + k];
      if (fct != 0) for (j = 0; j < 2 * d; j++) w[i * 2 * d + j] -= fct * w[k * 2 * d + j];
    }
  }
  for (i = 0; i < d; i++) for (j = 0; j < d; j++) inv[i * d + j] = w[i * 2 * d + d + j];
  return 0;
}

/* One problem.  N = n+m <= 64.  terminal_zero=1 reproduces the stale README table (V_T = v_T = 0
 * and no final cost, SURVEY finding 4); 0 is the v0.7.0 source (lqr.py:67-68,154-155). */
static int lqr_solve_one(int n, int m, int T, const REAL *F, const REAL *f, const REAL *C, const REAL *c, const REAL *x0,
                         int terminal_zero, REAL *states, REAL *actions, REAL *costs, REAL *Ko, REAL *ko, REAL *Vo, REAL *vo,
                         REAL *consto) {
  int N = n + m, i, j, t, p;
  REAL V[64 * 64], v[64], Q[64 * 64], q[64], FtV[64 * 64], Vf[64], inv[64 * 64], Quu[64 * 64];
  REAL *K = (REAL *)malloc(sizeof(REAL) * T * m * n), *k = (REAL *)malloc(sizeof(REAL) * T * m);
  REAL cst = 0;
  int rc = 0;
  for (i = 0; i < n; i++) { v[i] = terminal_zero ? (REAL)0 : c[i]; for (j = 0; j < n; j++) V[i * n + j] = terminal_zero ? (REAL)0 : C[i * N + j]; }
  for (t = T - 1; t >= 0; t--) {
    /* F^T V  [N,n] ; Q = C + F^T V F ; q = c + F^T V f + F^T v   (:74-78) */
    for (i = 0; i < N; i++) for (j = 0; j < n; j++) { REAL s = 0; for (p = 0; p < n; p++) s += F[p * N + i] * V[p * n + j]; FtV[i * n + j] = s; }
    for (i = 0; i < N; i++) for (j = 0; j < N; j++) { REAL s = 0; for (p = 0; p < n; p++) s += FtV[i * n + p] * F[p * N + j]; Q[i * N + j] = C[i * N + j] + s; }
    for (i = 0; i < N; i++) {
      REAL s1 = 0, s2 = 0;
      for (p = 0; p < n; p++) { s1 += FtV[i * n + p] * f[p]; s2 += F[p * N + i] * v[p]; }
      q[i] = c[i] + s1 + s2;
    }
    for (i = 0; i < m; i++) for (j = 0; j < m; j++) Quu[i * m + j] = Q[(n + i) * N + n + j];
    if (inverse(Quu, inv, m)) { rc = 1; break; }
    REAL *K_t = K + t * m * n, *k_t = k + t * m;
    for (i = 0; i < m; i++) { /* K = -inv Q_ux, k = -inv q_u  (:84-87) */
      for (j = 0; j < n; j++) { REAL s = 0; for (p = 0; p < m; p++) s += inv[i * m + p] * Q[(n + p) * N + j]; K_t[i * n + j] = -s; }
      REAL s = 0; for (p = 0; p < m; p++) s += inv[i * m + p] * q[n + p]; k_t[i] = -s;
    }
    /* const terms use W == Q, w == q built from the OLD V, v  (:107-121) */
    for (i = 0; i < n; i++) { REAL s = 0; for (p = 0; p < n; p++) s += V[i * n + p] * f[p]; Vf[i] = s; }
    REAL c1 = 0, c2 = 0, c3a = 0, c3b = 0;
    for (i = 0; i < m; i++) { REAL s = 0; for (p = 0; p < m; p++) s += Quu[i * m + p] * k_t[p]; c1 += k_t[i] * s; c2 += k_t[i] * q[n + i]; }
    for (i = 0; i < n; i++) { c3a += f[i] * Vf[i]; c3b += f[i] * v[i]; }
    cst += ((REAL)0.5 * c1 + c2 + ((REAL)0.5 * c3a + c3b));
    /* V, v update (:97-105) */
    REAL KtQuu[64 * 64], Vn[64 * 64], vn[64];
    for (i = 0; i < n; i++) for (j = 0; j < m; j++) { REAL s = 0; for (p = 0; p < m; p++) s += K_t[p * n + i] * Quu[p * m + j]; KtQuu[i * m + j] = s; }
    for (i = 0; i < n; i++) {
      for (j = 0; j < n; j++) {
        REAL s1 = 0, s2 = 0, s3 = 0;
        for (p = 0; p < m; p++) { s1 += Q[i * N + n + p] * K_t[p * n + j]; s2 += K_t[p * n + i] * Q[(n + p) * N + j]; s3 += KtQuu[i * m + p] * K_t[p * n + j]; }
        Vn[i * n + j] = Q[i * N + j] + s1 + s2 + s3;
      }
      REAL s1 = 0, s2 = 0, s3 = 0;
      for (p = 0; p < m; p++) { s1 += Q[i * N + n + p] * k_t[p]; s2 += K_t[p * n + i] * q[n + p]; s3 += KtQuu[i * m + p] * k_t[p]; }
      vn[i] = q[i] + s1 + s2 + s3;
    }
    memcpy(V, Vn, sizeof(REAL) * n * n); memcpy(v, vn, sizeof(REAL) * n);
    if (Vo) { memcpy(Vo + t * n * n, V, sizeof(REAL) * n * n); memcpy(vo + t * n, v, sizeof(REAL) * n); consto[t] = cst; }
  }
  if (!rc) { /* forward, :131-161 */
    REAL z[64];
    memcpy(states, x0, sizeof(REAL) * n);
    for (t = 0; t < T; t++) {
      const REAL *x = states + t * n; REAL *u = actions + t * m, *xn = states + (t + 1) * n;
      for (i = 0; i < m; i++) { REAL s = 0; for (p = 0; p < n; p++) s += K[t * m * n + i * n + p] * x[p]; u[i] = s + k[t * m + i]; }
      for (i = 0; i < n; i++) z[i] = x[i];
      for (i = 0; i < m; i++) z[n + i] = u[i];
      for (i = 0; i < n; i++) { REAL s = 0; for (p = 0; p < N; p++) s += F[i * N + p] * z[p]; xn[i] = s + f[i]; }
      REAL quad = 0, lin = 0;
      for (j = 0; j < N; j++) { REAL s = 0; for (p = 0; p < N; p++) s += z[p] * C[p * N + j]; quad += s * z[j]; lin += z[j] * c[j]; }
      costs[t] = (REAL)0.5 * quad + lin;
    }
    const REAL *x = states + T * n;
    REAL quad = 0, lin = 0;
    for (j = 0; j < n; j++) { REAL s = 0; for (p = 0; p < n; p++) s += x[p] * C[p * N + j]; quad += s * x[j]; lin += x[j] * c[j]; }
    costs[T] = terminal_zero ? (REAL)0 : (REAL)0.5 * quad + lin;
    if (Ko) { memcpy(Ko, K, sizeof(REAL) * T * m * n); memcpy(ko, k, sizeof(REAL) * T * m); }
  }
  free(K); free(k);
  return rc;
}

/* strides (in problems) of 0 share a matrix across the batch */
void oracle_lqr_solve(int B, int n, int m, int T, const REAL *F, int sF, const REAL *f, int sf, const REAL *C, int sC, const REAL *c,
                      int sc, const REAL *x0, int terminal_zero, REAL *states, REAL *actions, REAL *costs, REAL *K, REAL *k, REAL *V,
                      REAL *v, REAL *cst, int32_t *status, int nthreads) {
  int N = n + m;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (int b = 0; b < B; b++)
    status[b] = lqr_solve_one(n, m, T, F + (size_t)b * sF * n * N, f + (size_t)b * sf * n, C + (size_t)b * sC * N * N, c + (size_t)b * sc * N,
                              x0 + (size_t)b * n, terminal_zero, states + (size_t)b * (T + 1) * n, actions + (size_t)b * T * m,
                              costs + (size_t)b * (T + 1), K ? K + (size_t)b * T * m * n : NULL, k ? k + (size_t)b * T * m : NULL,
                              V ? V + (size_t)b * T * n * n : NULL, v ? v + (size_t)b * T * n : NULL, cst ? cst + (size_t)b * T : NULL);
}
