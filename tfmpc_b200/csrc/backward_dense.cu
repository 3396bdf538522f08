// Generic DENSE iLQR backward pass over STAGED derivative models (any n, m <= 32).
//
// This is iLQR.backward exactly as the reference defines it (tfmpc/solvers/ilqr.py:94-172): it consumes the
// TransitionApprox / CostApprox / FinalCostApprox buffers (f_x, f_u, l, l_x, l_u, l_xx, l_uu, l_xu per timestep; the
// caller may have produced them with tfmpc_env_linearize or by any other means) and runs all three controllers
// (:357-387 Cholesky, box-QP, bang-bang) with the reference's dispatch rule (:136-143).  The environment-specialised
// kernels (ilqr_small.cu, ilqr_warp.cu) fuse the linearisation instead; this kernel is the general path.
//
// Mapping: ONE WARP PER PROBLEM, persistent (grid-stride over problems), lane i owns row i of every matrix.
//  * each timestep's derivative block (up to 5 n^2 + 2n floats = 20.7 KB at n = m = 32) is staged into shared memory by
//    TMA bulk copies (cp.async.bulk + mbarrier, double buffered: the block of step t-1 lands while step t computes);
//    blocks whose rows are not 16-byte multiples (n or m not a multiple of 4) fall back to cooperative loads;
//  * products A^T B keep the result row in registers (lane = row, operand rows broadcast from shared memory with 128-bit
//    loads), so Q_xx, Q_uu, Q_ux and their regularised twins never round-trip through memory between the two GEMMs;
//  * the Cholesky factorisation of Q_uu_reg (and of the free block inside the box-QP) is right-looking with the row
//    in registers and the pivot column broadcast by warp shuffles; triangular solves run one right-hand-side column per
//    lane; the box-QP's control flow is per problem, hence warp-uniform: no divergence.
// Matrices are <= 32x32, so tensor cores are deliberately not used (BASELINE.json north_star).
#include <algorithm>

#include "small_core.cuh"

namespace {

constexpr unsigned FULL = 0xffffffffu;
constexpr int NP = 32;       // padded dimension
constexpr int LD = NP + 4;   // leading dimension of the work matrices: rows stay 16-byte aligned, column reads conflict-free

struct DenseArgs {
  int64_t B;
  int T, n, m, bounded;
  real mu;
  real low[NP], high[NP];
  const real *actions, *f_x, *f_u, *l, *l_x, *l_u, *l_xx, *l_uu, *l_xu, *fl, *fl_x, *fl_xx;
  real *K, *k, *J, *dV1, *dV2;
  int32_t *status;
};

// ---- mbarrier / bulk-copy wrappers (PTX ISA 8.x, sm_90+; SASS: UBLKCP, SYNCS) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
               "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// ---- small warp helpers ----------------------------------------------------------------------------------------------
__device__ __forceinline__ real warp_sum(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ real warp_max(real v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = r_max(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// acc[j] += a * row[j] for j < cols; `row` is a shared-memory row, read as broadcast (same address in every lane)
template <bool VEC>
__device__ __forceinline__ void axpy_row(real (&acc)[NP], real a, const real *row, int cols) {
  if (VEC) {
#pragma unroll
    for (int c = 0; c < NP / 4; c++)
      if (4 * c < cols) {
        const R4 r = *reinterpret_cast<const R4 *>(row + 4 * c);
#pragma unroll
        for (int q = 0; q < 4; q++) acc[4 * c + q] += a * r.v[q];
      }
  } else {
#pragma unroll
    for (int j = 0; j < NP; j++)
      if (j < cols) acc[j] += a * row[j];
  }
}
__device__ __forceinline__ void store_row(real *dst, const real (&v)[NP]) {  // dst: LD-layout row (16-byte aligned)
#pragma unroll
  for (int c = 0; c < NP / 4; c++) {
    R4 r;
#pragma unroll
    for (int q = 0; q < 4; q++) r.v[q] = v[4 * c + q];
    *reinterpret_cast<R4 *>(dst + 4 * c) = r;
  }
}
__device__ __forceinline__ void load_row(const real *src, real (&v)[NP]) {
#pragma unroll
  for (int c = 0; c < NP / 4; c++) {
    const R4 r = *reinterpret_cast<const R4 *>(src + 4 * c);
#pragma unroll
    for (int q = 0; q < 4; q++) v[4 * c + q] = r.v[q];
  }
}

// Right-looking Cholesky with row `lane` of the (masked) matrix in registers; the factor's rows are left in `row` and
// also written to Ls (LD layout).  fr = free mask (bit i set = row/column i participates; others become identity).
// Same operation order per element as the left-looking loop of the oracle.  Returns true on a non-positive pivot.
__device__ __forceinline__ bool chol_rows(real (&row)[NP], int d, unsigned fr, real *Ls, int lane) {
  bool fail = false;
  const bool me = (fr >> lane) & 1u;
#pragma unroll
  for (int j = 0; j < NP; j++) {
    if (j < d) {
      const bool fj = (fr >> j) & 1u;
      row[j] = (me && fj) ? row[j] : ((lane == j) ? (real)1 : (real)0);
    } else row[j] = 0;
  }
#pragma unroll
  for (int j = 0; j < NP; j++) {
    if (j < d) {
      real piv = __shfl_sync(FULL, row[j], j);  // A[j][j] after the previous updates
      if (!(piv > 0)) { fail = true; piv = 1; }
      const real dj = r_sqrt(piv);
      real lij = (lane == j) ? dj : ((lane > j) ? row[j] / dj : (real)0);
      row[j] = lij;
      // trailing update: A[i][c] -= L[i][j] L[c][j] for j < c <= i
#pragma unroll
      for (int c = j + 1; c < NP; c++) {
        if (c < d) {
          const real lcj = __shfl_sync(FULL, lij, c);
          if (lane >= c) row[c] -= lij * lcj;
        }
      }
    }
  }
  // keep the lower triangle only
#pragma unroll
  for (int j = 0; j < NP; j++)
    if (j > lane) row[j] = 0;
  store_row(Ls + lane * LD, row);
  __syncwarp();
  return fail;
}

// Solve (L L^T) Y = R for up to 32 right-hand-side columns, one column per lane: R and Y are LD-layout matrices
// (R[i*LD + col]); lanes with col >= ncols idle.  Order of accumulation = the oracle's chol_solve.
__device__ __forceinline__ void chol_solve_cols(const real *Ls, int d, const real *R, real *Y, int ncols, int lane, real scale) {
  real y[NP];
#pragma unroll
  for (int i = 0; i < NP; i++) {
    if (i < d) {
      real s = R[i * LD + lane];
#pragma unroll
      for (int k = 0; k < i; k++) s -= Ls[i * LD + k] * y[k];
      y[i] = s / Ls[i * LD + i];
    } else y[i] = 0;
  }
#pragma unroll
  for (int i = NP - 1; i >= 0; i--) {
    if (i < d) {
      real s = y[i];
#pragma unroll
      for (int k = i + 1; k < NP; k++)
        if (k < d) s -= Ls[k * LD + i] * y[k];
      y[i] = s / Ls[i * LD + i];
    }
  }
  if (lane < ncols) {
#pragma unroll
    for (int i = 0; i < NP; i++)
      if (i < d) Y[i * LD + lane] = scale * y[i];
  }
  __syncwarp();
}

// single right-hand side held one element per lane; result one element per lane
__device__ __forceinline__ real chol_solve_vec(const real *Ls, int d, real b, real *scratch, int lane) {
  scratch[lane * LD] = b;  // column 0 of an LD-layout scratch matrix
  __syncwarp();
  chol_solve_cols(Ls, d, scratch, scratch, 1, lane, (real)1);
  const real r = scratch[lane * LD];
  __syncwarp();
  return r;
}

// f(x) = 1/2 x^T H x + q^T x with row `lane` of H in registers and x, q one element per lane (optimization.py:8-11)
__device__ __forceinline__ real qp_value_w(const real (&H)[NP], real q, real x, int m, int lane) {
  real hx = 0;
#pragma unroll
  for (int j = 0; j < NP; j++)
    if (j < m) hx += H[j] * __shfl_sync(FULL, x, j);
  const bool act = lane < m;
  const real quad = warp_sum(act ? x * hx : (real)0), lin = warp_sum(act ? q * x : (real)0);
  return (real)0.5 * quad + lin;
}

// Projected-Newton box-QP, one problem per warp, lane = component (optimization.py:6-101, :121-127).  Hrow = row `lane`
// of H.  On exit: x (per lane), the free mask, and the masked Cholesky factor of H[free,free] in Ls.  Returns 0 / 2.
__device__ __forceinline__ int boxqp_dense(const real (&Hrow)[NP], real q, real lo, real hi, real &x, unsigned &free_mask, real *Ls, real *scratch,
                                           int m, int lane) {
  const real rtol = (real)1e-8, armijo = (real)0.1, eps = (real)1e-6;
  const bool act = lane < m;
  const unsigned all = (m >= 32) ? FULL : ((1u << m) - 1u);
  unsigned clamped = 0;
  free_mask = all;
  real value = qp_value_w(Hrow, q, x, m, lane), old_value = 0;
  int status = 0;
  for (int it = 0; it < 100; it++) {
    if (it > 0 && (old_value - value) < rtol * r_abs(old_value)) break;  // :27
    old_value = value;
    real hx = 0;
#pragma unroll
    for (int j = 0; j < NP; j++)
      if (j < m) hx += Hrow[j] * __shfl_sync(FULL, x, j);
    const real g = q + hx;  // :34
    const bool c = act && ((r_abs(x - lo) < eps && g > 0) || (r_abs(hi - x) < eps && g < 0));  // :121-127
    const unsigned cm = __ballot_sync(FULL, c);
    const bool changed = cm != clamped;
    clamped = cm;
    free_mask = all & ~cm;
    if (it == 0 || changed) {  // :37-51
      real row[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) row[j] = Hrow[j];
      if (chol_rows(row, m, free_mask, Ls, lane)) { status = 2; break; }
    }
    if (cm == all) break;  // :53
    const bool fr = act && !c;
    const real gn = warp_sum(fr ? g * g : (real)0);
    if (r_sqrt(gn) < eps) break;  // :58-62
    real hxc = 0;  // grad_clamped = q + H (x * clamped), :65
    {
      const real xc0 = c ? x : (real)0;
#pragma unroll
      for (int j = 0; j < NP; j++)
        if (j < m) hxc += Hrow[j] * __shfl_sync(FULL, xc0, j);
    }
    const real rhs = fr ? q + hxc : (real)0;
    const real sol = chol_solve_vec(Ls, m, rhs, scratch, lane);
    const real search = fr ? -sol - x : (real)0;  // :70
    const real sdotg = warp_sum(act ? search * g : (real)0);
    if (sdotg >= 0) break;  // :75-79
    double step = 1.0;
    real xc, vc;
    for (;;) {  // :82-95
      const real st = (real)step;
      xc = act ? r_clip(x + st * search, lo, hi) : (real)0;
      if (!__any_sync(FULL, act && xc != x)) { xc = x; vc = old_value; break; }  // degenerate backtracking (see small_core.cuh)
      vc = qp_value_w(Hrow, q, xc, m, lane);
      if (!((vc - old_value) / (st * sdotg) < armijo)) break;
      step *= 0.6;
      if (step < 1e-22) {
        const real s2 = (real)step;
        xc = act ? r_clip(x + s2 * search, lo, hi) : (real)0;
        vc = qp_value_w(Hrow, q, xc, m, lane);
        break;
      }
    }
    x = xc;
    value = vc;
  }
  return status;
}

template <bool TMA>
__global__ void __launch_bounds__(32) k_backward_dense(DenseArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x, n = a.n, m = a.m, T = a.T;
  // staged block layout (compact, as in global memory): f_x n*n | f_u n*m | l_xx n*n | l_uu m*m | l_xu n*m | l_x n | l_u m
  const int o_fx = 0, o_fu = o_fx + n * n, o_lxx = o_fu + n * m, o_luu = o_lxx + n * n, o_lxu = o_luu + m * m, o_lx = o_lxu + n * m,
            o_lu = o_lx + n;
  const int blk = ((o_lu + m + 3) / 4) * 4;
  real *stage = reinterpret_cast<real *>(smem_raw);                 // [2][blk]
  real *Vxx = stage + 2 * blk;                                     // work matrices, LD layout
  real *Qxx = Vxx + NP * LD, *Quu = Qxx + NP * LD, *Qux = Quu + NP * LD, *QuuR = Qux + NP * LD, *QuxR = QuuR + NP * LD;
  real *Km = QuxR + NP * LD, *Ls = Km + NP * LD, *Tmp = Ls + NP * LD;
  uint64_t *bars = reinterpret_cast<uint64_t *>(Tmp + NP * LD);    // [2]
  const bool vec = (n % 4 == 0) && (m % 4 == 0);

  if (TMA && lane == 0) { mbar_init(&bars[0], 1); mbar_init(&bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  uint32_t phase[2] = {0, 0};

  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    auto issue = [&](int t, int buf) {  // stage the derivative block of timestep t
      real *dst = stage + buf * blk;
      const int64_t bt = b * T + t;
      const real *srcs[7] = {a.f_x + bt * n * n, a.f_u + bt * n * m, a.l_xx + bt * n * n, a.l_uu + bt * m * m, a.l_xu + bt * n * m,
                             a.l_x + bt * n, a.l_u + bt * m};
      const int offs[7] = {o_fx, o_fu, o_lxx, o_luu, o_lxu, o_lx, o_lu};
      const int cnts[7] = {n * n, n * m, n * n, m * m, n * m, n, m};
      if (TMA) {
        if (lane == 0) {
          uint32_t bytes = 0;
#pragma unroll
          for (int s = 0; s < 7; s++) bytes += (uint32_t)(cnts[s] * sizeof(real));
          mbar_expect_tx(&bars[buf], bytes);
#pragma unroll
          for (int s = 0; s < 7; s++) bulk_g2s(dst + offs[s], srcs[s], (uint32_t)(cnts[s] * sizeof(real)), &bars[buf]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < 7; s++)
          for (int i = lane; i < cnts[s]; i += 32) dst[offs[s] + i] = srcs[s][i];
      }
    };
    auto wait_stage = [&](int buf) {
      if (TMA) { mbar_wait(&bars[buf], phase[buf]); phase[buf] ^= 1; }
      else __syncwarp();
    };

    // ---- terminal value function (ilqr.py:101-106)
    for (int i = lane; i < NP * LD; i += 32) Vxx[i] = 0;
    __syncwarp();
    for (int i = lane; i < n * n; i += 32) Vxx[(i / n) * LD + (i % n)] = a.fl_xx[b * n * n + i];
    real Vx = lane < n ? a.fl_x[b * n + lane] : (real)0;
    real J = a.fl[b], dV1 = 0, dV2 = 0;
    int status = 0;
    __syncwarp();
    issue(T - 1, 0);

    for (int t = T - 1; t >= 0; t--) {
      const int buf = (T - 1 - t) & 1;
      wait_stage(buf);
      if (t > 0) issue(t - 1, buf ^ 1);
      const real *S = stage + buf * blk;
      const real *fx = S + o_fx, *fu = S + o_fu, *lxx = S + o_lxx, *luu = S + o_luu, *lxu = S + o_lxu;
      const real u = lane < m ? a.actions[(b * T + t) * m + lane] : (real)0;

      // Q_x = l_x + f_x^T V_x ; Q_u = l_u + f_u^T V_x   (:122-123)
      real Qx = 0, Qu = 0;
      for (int p = 0; p < n; p++) {
        const real vp = __shfl_sync(FULL, Vx, p);
        if (lane < n) Qx += fx[p * n + lane] * vp;
        if (lane < m) Qu += fu[p * m + lane] * vp;
      }
      Qx = lane < n ? S[o_lx + lane] + Qx : (real)0;
      Qu = lane < m ? S[o_lu + lane] + Qu : (real)0;

      // count_nonzero(V_xx) > 0  (:137), before V_xx is overwritten
      bool nzl = false;
      if (lane < n)
        for (int j = 0; j < n; j++) nzl = nzl || (Vxx[lane * LD + j] != 0);
      const bool any_nz = __any_sync(FULL, nzl);

      // row `lane` of f_x^T V_xx (:125), then Q_xx = l_xx + (f_x^T V_xx) f_x (:129)
      real r1[NP], acc[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) r1[j] = 0;
      for (int p = 0; p < n; p++) {
        const real av = lane < n ? fx[p * n + lane] : (real)0;
        axpy_row<true>(r1, av, Vxx + p * LD, n);
      }
#pragma unroll
      for (int j = 0; j < NP; j++) acc[j] = 0;
#pragma unroll
      for (int p = 0; p < NP; p++)
        if (p < n) { if (vec) axpy_row<true>(acc, r1[p], fx + p * n, n); else axpy_row<false>(acc, r1[p], fx + p * n, n); }
#pragma unroll
      for (int j = 0; j < NP; j++) acc[j] = (lane < n && j < n) ? lxx[lane * n + j] + acc[j] : (real)0;
      store_row(Qxx + lane * LD, acc);

      // rows of f_u^T V_xx and f_u^T (V_xx + mu I) (:126-127), then Q_uu, Q_ux and their regularised twins (:130-134)
      real r2[NP], r2r[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) { r2[j] = 0; r2r[j] = 0; }
#pragma unroll
      for (int p = 0; p < NP; p++) {
        if (p < n) {
          const real av = lane < m ? fu[p * m + lane] : (real)0;
          real vrow[NP];
          load_row(Vxx + p * LD, vrow);
#pragma unroll
          for (int j = 0; j < NP; j++) {
            r2[j] += av * vrow[j];
            r2r[j] += av * (j == p ? vrow[j] + a.mu * (real)1 : vrow[j]);
          }
        }
      }
      real quu[NP], qux[NP];
#pragma unroll
      for (int pass = 0; pass < 2; pass++) {
#pragma unroll
        for (int j = 0; j < NP; j++) { quu[j] = 0; qux[j] = 0; }
#pragma unroll
        for (int p = 0; p < NP; p++) {
          if (p < n) {
            const real rp = pass == 0 ? r2[p] : r2r[p];
            if (vec) { axpy_row<true>(quu, rp, fu + p * m, m); axpy_row<true>(qux, rp, fx + p * n, n); }
            else { axpy_row<false>(quu, rp, fu + p * m, m); axpy_row<false>(qux, rp, fx + p * n, n); }
          }
        }
#pragma unroll
        for (int j = 0; j < NP; j++) {
          quu[j] = (lane < m && j < m) ? luu[lane * m + j] + quu[j] : (real)0;
          qux[j] = (lane < m && j < n) ? lxu[j * m + lane] + qux[j] : (real)0;  // l_xu^T
        }
        store_row((pass == 0 ? Quu : QuuR) + lane * LD, quu);
        store_row((pass == 0 ? Qux : QuxR) + lane * LD, qux);
      }
      __syncwarp();
      // here quu / qux hold the REGULARISED rows (pass 1)

      // ---- controller (:136-143)
      real kk = 0;
      if (a.bounded && any_nz) {  // _get_constrained_controller :364-387
        const real lo = lane < m ? a.low[lane] - u : (real)0, hi = lane < m ? a.high[lane] - u : (real)0;
        kk = (lo + hi) / (real)2;
        unsigned fr;
        const int st = boxqp_dense(quu, Qu, lo, hi, kk, fr, Ls, Tmp, m, lane);
        if (st) status = 2;
        // K[free] = -cholesky_solve(Hfree, Q_ux_reg[free]), clamped rows 0: zero the clamped rows of the right-hand side
        for (int i = lane; i < NP * LD; i += 32) Tmp[i] = 0;
        __syncwarp();
        if (lane < m && ((fr >> lane) & 1u) && !st) store_row(Tmp + lane * LD, qux);
        __syncwarp();
        chol_solve_cols(Ls, m, Tmp, Km, n, lane, (real)-1);
        if (st) { for (int i = lane; i < NP * LD; i += 32) Km[i] = 0; __syncwarp(); }
        // clamped rows of the solution are exactly 0 (identity rows, zero right-hand side); normalise -0 to 0
      } else if (a.bounded) {  // bang-bang :139-141
        for (int i = lane; i < NP * LD; i += 32) Km[i] = 0;
        kk = lane < m ? ((Qu >= 0) ? a.low[lane] - u : a.high[lane] - u) : (real)0;
        __syncwarp();
      } else {  // _get_unconstrained_controller :357-362
        real row[NP];
#pragma unroll
        for (int j = 0; j < NP; j++) row[j] = quu[j];
        const unsigned all = (m >= 32) ? FULL : ((1u << m) - 1u);
        if (chol_rows(row, m, all, Ls, lane)) {  // the caller retries with a larger mu (ilqr.py:305-309)
          status = 1;
          if (t > 0) wait_stage(buf ^ 1);  // drain the bulk copy already in flight so the barrier phases stay in step
          break;
        }
        kk = -chol_solve_vec(Ls, m, Qu, Tmp, lane);
        chol_solve_cols(Ls, m, QuxR, Km, n, lane, (real)-1);
      }
      if (lane >= m) kk = 0;

      // ---- value update with the UNregularised Q (:145-162)
      // KtQuu row (lane < n): sum_p K[p][lane] Q_uu[p][:]
      real kq[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) kq[j] = 0;
      for (int p = 0; p < m; p++) {
        const real av = lane < n ? Km[p * LD + lane] : (real)0;
        axpy_row<true>(kq, av, Quu + p * LD, m);
      }
      // V_x = Q_x + Q_ux^T k + K^T Q_u + (K^T Q_uu) k
      real a1 = 0, a2 = 0, a3 = 0;
      for (int p = 0; p < m; p++) {
        const real kp = __shfl_sync(FULL, kk, p), qup = __shfl_sync(FULL, Qu, p);
        if (lane < n) { a1 += Qux[p * LD + lane] * kp; a2 += Km[p * LD + lane] * qup; }
      }
#pragma unroll
      for (int p = 0; p < NP; p++)
        if (p < m) a3 += kq[p] * __shfl_sync(FULL, kk, p);
      Vx = lane < n ? Qx + a1 + a2 + a3 : (real)0;
      // V_xx row = Q_xx + Q_ux^T K + K^T Q_ux + (K^T Q_uu) K
      real b1[NP], b2[NP], b3[NP];
#pragma unroll
      for (int j = 0; j < NP; j++) { b1[j] = 0; b2[j] = 0; b3[j] = 0; }
      for (int p = 0; p < m; p++) {
        const real q1 = lane < n ? Qux[p * LD + lane] : (real)0, k1 = lane < n ? Km[p * LD + lane] : (real)0;
        axpy_row<true>(b1, q1, Km + p * LD, n);
        axpy_row<true>(b2, k1, Qux + p * LD, n);
      }
#pragma unroll
      for (int p = 0; p < NP; p++)
        if (p < m) axpy_row<true>(b3, kq[p], Km + p * LD, n);
      real vn[NP];
      load_row(Qxx + lane * LD, vn);
#pragma unroll
      for (int j = 0; j < NP; j++) vn[j] = (lane < n && j < n) ? vn[j] + b1[j] + b2[j] + b3[j] : (real)0;
      store_row(Tmp + lane * LD, vn);
      __syncwarp();
#pragma unroll
      for (int j = 0; j < NP; j++) vn[j] = (lane < n && j < n) ? (real)0.5 * (vn[j] + Tmp[j * LD + lane]) : (real)0;  // :162
      store_row(Vxx + lane * LD, vn);

      // ---- J, dV1, dV2 (:164-167) and outputs
      J += a.l[b * T + t];
      dV1 += warp_sum(lane < m ? kk * Qu : (real)0);
      {
        real s = 0;  // (k^T Q_uu)_lane = sum_i k_i Q_uu[i][lane]
        for (int i = 0; i < m; i++) s += __shfl_sync(FULL, kk, i) * Quu[i * LD + lane];
        dV2 += (real)0.5 * warp_sum(lane < m ? s * kk : (real)0);
      }
      if (lane < m) a.k[(b * T + t) * m + lane] = kk;
      for (int i = lane; i < m * n; i += 32) a.K[(b * T + t) * m * n + i] = Km[(i / n) * LD + (i % n)];
      __syncwarp();
    }
    if (lane == 0) {
      a.J[b] = J; a.dV1[b] = dV1; a.dV2[b] = dV2;
      if (a.status) a.status[b] = status;
    }
    __syncwarp();
  }
}

}  // namespace

int dense_backward_launch(int64_t B, int T, int n, int m, int bounded, const double *low, const double *high, const real *actions,
                          const real *f_x, const real *f_u, const real *l, const real *l_x, const real *l_u, const real *l_xx,
                          const real *l_uu, const real *l_xu, const real *fl, const real *fl_x, const real *fl_xx, double mu, real *K,
                          real *k, real *J, real *dV1, real *dV2, int32_t *status, cudaStream_t s) {
  DenseArgs a;
  a.B = B; a.T = T; a.n = n; a.m = m; a.bounded = bounded; a.mu = (real)mu;
  for (int i = 0; i < NP; i++) { a.low[i] = i < m ? (real)low[i] : (real)0; a.high[i] = i < m ? (real)high[i] : (real)0; }
  a.actions = actions; a.f_x = f_x; a.f_u = f_u; a.l = l; a.l_x = l_x; a.l_u = l_u; a.l_xx = l_xx; a.l_uu = l_uu; a.l_xu = l_xu;
  a.fl = fl; a.fl_x = fl_x; a.fl_xx = fl_xx; a.K = K; a.k = k; a.J = J; a.dV1 = dV1; a.dV2 = dV2; a.status = status;
  const int blk = ((2 * n * n + 2 * n * m + m * m + n + m + 3) / 4) * 4;
  const size_t smem = sizeof(real) * ((size_t)2 * blk + (size_t)9 * NP * LD) + 2 * sizeof(uint64_t) + 16;
  // bulk copies need 16-byte aligned, 16-byte-multiple segments: every per-(b,t) segment starts at a multiple of its own
  // size, so it suffices that n and m are multiples of 4 (fp32) / 2 (fp64) and the base pointers are 16-byte aligned
  const int q = 16 / (int)sizeof(real);
  bool tma = (n % q == 0) && (m % q == 0);
  const void *ptrs[7] = {f_x, f_u, l_xx, l_uu, l_xu, l_x, l_u};
  for (int i = 0; i < 7; i++) tma = tma && (((uintptr_t)ptrs[i]) % 16 == 0);
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)sms * 4);
  if (tma) {
    CUDA_TRY(cudaFuncSetAttribute(k_backward_dense<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_backward_dense<true><<<grid, 32, smem, s>>>(a);
  } else {
    CUDA_TRY(cudaFuncSetAttribute(k_backward_dense<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_backward_dense<false><<<grid, 32, smem, s>>>(a);
  }
  LAUNCH_CHECK();
  return TFMPC_OK;
}
