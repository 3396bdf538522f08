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
#include <cstring>

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

// ---- shared-memory carve-up of one warp ---------------------------------------------------------------------------------
struct Smem {
  real *stage;  // [2][blk] staged derivative blocks (compact, as in global memory)
  real *Vxx, *Qxx, *Quu, *Qux, *QuuR, *QuxR, *Km, *Ls, *Tmp;  // work matrices, LD layout
  uint64_t *bars;
  int blk, o_fx, o_fu, o_lxx, o_luu, o_lxu, o_lx, o_lu;
};
__device__ __forceinline__ Smem carve_smem(unsigned char *raw, int n, int m) {
  Smem s;
  s.o_fx = 0; s.o_fu = s.o_fx + n * n; s.o_lxx = s.o_fu + n * m; s.o_luu = s.o_lxx + n * n; s.o_lxu = s.o_luu + m * m;
  s.o_lx = s.o_lxu + n * m; s.o_lu = s.o_lx + n;
  s.blk = ((s.o_lu + m + 3) / 4) * 4;
  s.stage = reinterpret_cast<real *>(raw);
  s.Vxx = s.stage + 2 * s.blk;
  s.Qxx = s.Vxx + NP * LD; s.Quu = s.Qxx + NP * LD; s.Qux = s.Quu + NP * LD; s.QuuR = s.Qux + NP * LD; s.QuxR = s.QuuR + NP * LD;
  s.Km = s.QuxR + NP * LD; s.Ls = s.Km + NP * LD; s.Tmp = s.Ls + NP * LD;
  s.bars = reinterpret_cast<uint64_t *>(s.Tmp + NP * LD);
  return s;
}
inline size_t dense_smem_bytes(int n, int m) {
  const int blk = ((2 * n * n + 2 * n * m + m * m + n + m + 3) / 4) * 4;
  return sizeof(real) * ((size_t)2 * blk + (size_t)9 * NP * LD) + 2 * sizeof(uint64_t) + 16;
}

// One timestep of iLQR.backward (ilqr.py:119-167) for one problem, the derivative block being in shared memory at S.
// Carries V_x (one element per lane), V_xx (shared memory), J, dV1, dV2; leaves K_t in sm.Km (LD layout) and returns k_t
// (one element per lane) in kk.  Returns 0, 1 (Cholesky of Q_uu_reg failed) or 2 (box-QP factorisation failed).
__device__ __noinline__ int dense_step(const Smem &sm, const real *S, int n, int m, int bounded, real mu, const real *low, const real *high,
                                       real u, real l_t, real &Vx, real &J, real &dV1, real &dV2, real &kk, int lane) {
  const bool vec = (n % 4 == 0) && (m % 4 == 0);
  const real *fx = S + sm.o_fx, *fu = S + sm.o_fu, *lxx = S + sm.o_lxx, *luu = S + sm.o_luu, *lxu = S + sm.o_lxu;
  real *Vxx = sm.Vxx, *Qxx = sm.Qxx, *Quu = sm.Quu, *Qux = sm.Qux, *QuuR = sm.QuuR, *QuxR = sm.QuxR, *Km = sm.Km, *Ls = sm.Ls, *Tmp = sm.Tmp;
  int status = 0;

  // Q_x = l_x + f_x^T V_x ; Q_u = l_u + f_u^T V_x   (:122-123)
  real Qx = 0, Qu = 0;
  for (int p = 0; p < n; p++) {
    const real vp = __shfl_sync(FULL, Vx, p);
    if (lane < n) Qx += fx[p * n + lane] * vp;
    if (lane < m) Qu += fu[p * m + lane] * vp;
  }
  Qx = lane < n ? S[sm.o_lx + lane] + Qx : (real)0;
  Qu = lane < m ? S[sm.o_lu + lane] + Qu : (real)0;

  // count_nonzero(V_xx) > 0  (:137), before V_xx is overwritten
  bool nzl = false;
  if (lane < n)
    for (int j = 0; j < n; j++) nzl = nzl || (Vxx[lane * LD + j] != 0);
  const bool any_nz = __any_sync(FULL, nzl);

  real acc[NP], acc2[NP];
  // row `lane` of f_x^T V_xx (:125) -> Tmp, then Q_xx = l_xx + (f_x^T V_xx) f_x (:129)
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = 0;
  for (int p = 0; p < n; p++) axpy_row<true>(acc, lane < n ? fx[p * n + lane] : (real)0, Vxx + p * LD, n);
  store_row(Tmp + lane * LD, acc);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = 0;
  for (int p = 0; p < n; p++) {
    const real rp = Tmp[lane * LD + p];
    if (vec) axpy_row<true>(acc, rp, fx + p * n, n); else axpy_row<false>(acc, rp, fx + p * n, n);
  }
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = (lane < n && j < n) ? lxx[lane * n + j] + acc[j] : (real)0;
  store_row(Qxx + lane * LD, acc);
  __syncwarp();

  // rows of f_u^T V_xx -> Tmp and f_u^T (V_xx + mu I) -> Ls (scratch here) (:126-127)
#pragma unroll
  for (int j = 0; j < NP; j++) { acc[j] = 0; acc2[j] = 0; }
  for (int p = 0; p < n; p++) {
    const real av = lane < m ? fu[p * m + lane] : (real)0;
    real vrow[NP];
    load_row(Vxx + p * LD, vrow);
#pragma unroll
    for (int j = 0; j < NP; j++) {
      acc[j] += av * vrow[j];
      acc2[j] += av * (j == p ? vrow[j] + mu * (real)1 : vrow[j]);
    }
  }
  store_row(Tmp + lane * LD, acc);
  store_row(Ls + lane * LD, acc2);
  __syncwarp();
  // Q_uu, Q_ux (:130-131) and the regularised twins (:133-134)
  real quu[NP], qux[NP];
#pragma unroll
  for (int pass = 0; pass < 2; pass++) {
    const real *R = pass == 0 ? Tmp : Ls;
#pragma unroll
    for (int j = 0; j < NP; j++) { quu[j] = 0; qux[j] = 0; }
    for (int p = 0; p < n; p++) {
      const real rp = R[lane * LD + p];
      if (vec) { axpy_row<true>(quu, rp, fu + p * m, m); axpy_row<true>(qux, rp, fx + p * n, n); }
      else { axpy_row<false>(quu, rp, fu + p * m, m); axpy_row<false>(qux, rp, fx + p * n, n); }
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
  // quu / qux now hold the REGULARISED rows

  // ---- controller (:136-143)
  kk = 0;
  if (bounded && any_nz) {  // _get_constrained_controller :364-387
    const real lo = lane < m ? low[lane] - u : (real)0, hi = lane < m ? high[lane] - u : (real)0;
    kk = (lo + hi) / (real)2;
    unsigned fr;
    const int st = boxqp_dense(quu, Qu, lo, hi, kk, fr, Ls, Tmp, m, lane);
    if (st) status = 2;
    // K[free] = -cholesky_solve(Hfree, Q_ux_reg[free]); clamped rows 0 (zero right-hand side under an identity row)
    for (int i = lane; i < NP * LD; i += 32) Tmp[i] = 0;
    __syncwarp();
    if (lane < m && ((fr >> lane) & 1u) && !st) store_row(Tmp + lane * LD, qux);
    __syncwarp();
    chol_solve_cols(Ls, m, Tmp, Km, n, lane, (real)-1);
    if (st) { for (int i = lane; i < NP * LD; i += 32) Km[i] = 0; __syncwarp(); }
  } else if (bounded) {  // bang-bang :139-141
    for (int i = lane; i < NP * LD; i += 32) Km[i] = 0;
    kk = lane < m ? ((Qu >= 0) ? low[lane] - u : high[lane] - u) : (real)0;
    __syncwarp();
  } else {  // _get_unconstrained_controller :357-362
#pragma unroll
    for (int j = 0; j < NP; j++) acc[j] = quu[j];
    const unsigned all = (m >= 32) ? FULL : ((1u << m) - 1u);
    if (chol_rows(acc, m, all, Ls, lane)) return 1;  // the caller retries with a larger mu (ilqr.py:305-309)
    kk = -chol_solve_vec(Ls, m, Qu, Tmp, lane);
    chol_solve_cols(Ls, m, QuxR, Km, n, lane, (real)-1);
  }
  if (lane >= m) kk = 0;

  // ---- value update with the UNregularised Q (:145-162)
  // row `lane` of K^T Q_uu -> Tmp
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = 0;
  for (int p = 0; p < m; p++) axpy_row<true>(acc, lane < n ? Km[p * LD + lane] : (real)0, Quu + p * LD, m);
  store_row(Tmp + lane * LD, acc);
  __syncwarp();
  // V_x = Q_x + Q_ux^T k + K^T Q_u + (K^T Q_uu) k
  real a1 = 0, a2 = 0, a3 = 0;
  for (int p = 0; p < m; p++) {
    const real kp = __shfl_sync(FULL, kk, p), qup = __shfl_sync(FULL, Qu, p);
    if (lane < n) { a1 += Qux[p * LD + lane] * kp; a2 += Km[p * LD + lane] * qup; a3 += Tmp[lane * LD + p] * kp; }
  }
  Vx = lane < n ? Qx + a1 + a2 + a3 : (real)0;
  // V_xx row = Q_xx + Q_ux^T K + K^T Q_ux + (K^T Q_uu) K
  real b1[NP];
  load_row(Qxx + lane * LD, acc);
#pragma unroll
  for (int j = 0; j < NP; j++) b1[j] = 0;
  for (int p = 0; p < m; p++) axpy_row<true>(b1, lane < n ? Qux[p * LD + lane] : (real)0, Km + p * LD, n);
#pragma unroll
  for (int j = 0; j < NP; j++) { acc[j] += b1[j]; b1[j] = 0; }
  for (int p = 0; p < m; p++) axpy_row<true>(b1, lane < n ? Km[p * LD + lane] : (real)0, Qux + p * LD, n);
#pragma unroll
  for (int j = 0; j < NP; j++) { acc[j] += b1[j]; b1[j] = 0; }
  for (int p = 0; p < m; p++) axpy_row<true>(b1, Tmp[lane * LD + p], Km + p * LD, n);
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = (lane < n && j < n) ? acc[j] + b1[j] : (real)0;
  __syncwarp();
  store_row(Tmp + lane * LD, acc);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < NP; j++) acc[j] = (lane < n && j < n) ? (real)0.5 * (acc[j] + Tmp[j * LD + lane]) : (real)0;  // :162
  store_row(Vxx + lane * LD, acc);

  // ---- J, dV1, dV2 (:164-167)
  J += l_t;
  dV1 += warp_sum(lane < m ? kk * Qu : (real)0);
  {
    real s = 0;  // (k^T Q_uu)_lane = sum_i k_i Q_uu[i][lane]
    for (int i = 0; i < m; i++) s += __shfl_sync(FULL, kk, i) * Quu[i * LD + lane];
    dV2 += (real)0.5 * warp_sum(lane < m ? s * kk : (real)0);
  }
  __syncwarp();
  return status;
}

__device__ __forceinline__ void init_value(const Smem &sm, int n, const real *fl_xx_b, int lane) {
  for (int i = lane; i < NP * LD; i += 32) sm.Vxx[i] = 0;
  __syncwarp();
  for (int i = lane; i < n * n; i += 32) sm.Vxx[(i / n) * LD + (i % n)] = fl_xx_b[i];
  __syncwarp();
}

// Stage API: backward pass over caller-supplied models, blocks staged by TMA bulk copies (or cooperative loads).
template <bool TMA>
__global__ void __launch_bounds__(32) k_backward_dense(DenseArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x, n = a.n, m = a.m, T = a.T;
  const Smem sm = carve_smem(smem_raw, n, m);
  if (TMA && lane == 0) { mbar_init(&sm.bars[0], 1); mbar_init(&sm.bars[1], 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncwarp();
  uint32_t phase[2] = {0, 0};

  for (int64_t b = blockIdx.x; b < a.B; b += gridDim.x) {
    auto issue = [&](int t, int buf) {  // stage the derivative block of timestep t
      real *dst = sm.stage + buf * sm.blk;
      const int64_t bt = b * T + t;
      const real *srcs[7] = {a.f_x + bt * n * n, a.f_u + bt * n * m, a.l_xx + bt * n * n, a.l_uu + bt * m * m, a.l_xu + bt * n * m,
                             a.l_x + bt * n, a.l_u + bt * m};
      const int offs[7] = {sm.o_fx, sm.o_fu, sm.o_lxx, sm.o_luu, sm.o_lxu, sm.o_lx, sm.o_lu};
      const int cnts[7] = {n * n, n * m, n * n, m * m, n * m, n, m};
      if (TMA) {
        if (lane == 0) {
          uint32_t bytes = 0;
#pragma unroll
          for (int s = 0; s < 7; s++) bytes += (uint32_t)(cnts[s] * sizeof(real));
          mbar_expect_tx(&sm.bars[buf], bytes);
#pragma unroll
          for (int s = 0; s < 7; s++) bulk_g2s(dst + offs[s], srcs[s], (uint32_t)(cnts[s] * sizeof(real)), &sm.bars[buf]);
        }
      } else {
#pragma unroll
        for (int s = 0; s < 7; s++)
          for (int i = lane; i < cnts[s]; i += 32) dst[offs[s] + i] = srcs[s][i];
      }
    };
    auto wait_stage = [&](int buf) {
      if (TMA) { mbar_wait(&sm.bars[buf], phase[buf]); phase[buf] ^= 1; }
      else __syncwarp();
    };

    init_value(sm, n, a.fl_xx + b * n * n, lane);  // terminal value function (ilqr.py:101-106)
    real Vx = lane < n ? a.fl_x[b * n + lane] : (real)0;
    real J = a.fl[b], dV1 = 0, dV2 = 0;
    int status = 0;
    issue(T - 1, 0);
    for (int t = T - 1; t >= 0; t--) {
      const int buf = (T - 1 - t) & 1;
      wait_stage(buf);
      if (t > 0) issue(t - 1, buf ^ 1);
      const real u = lane < m ? a.actions[(b * T + t) * m + lane] : (real)0;
      real kk;
      const int st = dense_step(sm, sm.stage + buf * sm.blk, n, m, a.bounded, a.mu, a.low, a.high, u, a.l[b * T + t], Vx, J, dV1, dV2, kk, lane);
      if (st == 1) {
        status = 1;
        if (t > 0) wait_stage(buf ^ 1);  // drain the bulk copy already in flight so the barrier phases stay in step
        break;
      }
      if (st) status = st;
      if (lane < m) a.k[(b * T + t) * m + lane] = kk;
      for (int i = lane; i < m * n; i += 32) a.K[(b * T + t) * m * n + i] = sm.Km[(i / n) * LD + (i % n)];
      __syncwarp();
    }
    if (lane == 0) {
      a.J[b] = J; a.dV1[b] = dV1; a.dV2[b] = dV2;
      if (a.status) a.status[b] = status;
    }
    __syncwarp();
  }
}

// ---- full solve for NavigationLQR of any dimension (the environments without a specialised kernel) --------------------------
// One warp per problem, persistent with an atomic work queue (the structure of kw_solve in ilqr_warp.cu): start rollout,
// then the reference's outer loop (ilqr.py:214-283) with the dense backward sweep above -- each timestep's derivative
// block is written straight into the shared-memory stage buffer by the lanes (f_x = f_u = I, l_xx = 2I, l_uu = 2 beta I,
// reference lqr/navigation/__init__.py:30-47), so nothing is staged through HBM -- and a first-accept line search whose
// rollouts apply the dense feedback K_t (x - x_hat) with warp shuffles.
struct DenseSolveArgs {
  int64_t B;
  int T, n, bounded;
  real beta, goal[NP], low[NP], high[NP];
  IlqrOpts o;
  const real *x0, *u_init;
  real *states, *actions, *costs;
  int32_t *stats;
  unsigned long long *counter;
  real *ws;  // per problem: candidate X (T+1)n | candidate U Tn | k Tn | K T n n
};

__device__ __forceinline__ real navlqr_cost(const DenseSolveArgs &a, real x, real u, bool final, int lane) {
  const bool act = lane < a.n;
  const real c1 = warp_sum(act ? (x - a.goal[lane]) * (x - a.goal[lane]) : (real)0);  // lqr/navigation/__init__.py:34-47
  if (final) return c1;
  const real c2 = warp_sum(act ? u * u : (real)0);
  return c1 + a.beta * c2;
}

__global__ void __launch_bounds__(32) k_solve_dense_navlqr(DenseSolveArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x, n = a.n, m = a.n, T = a.T;
  const Smem sm = carve_smem(smem_raw, n, m);
  const bool act = lane < n;
  const int64_t per = (int64_t)(3 * T + 1) * n + (int64_t)T * n * n;
  // the derivative block is constant except l_x, l_u: fill the constant part once
  real *S = sm.stage;
  for (int i = lane; i < sm.blk; i += 32) S[i] = 0;
  __syncwarp();
  if (act) {
    S[sm.o_fx + lane * n + lane] = 1; S[sm.o_fu + lane * m + lane] = 1;
    S[sm.o_lxx + lane * n + lane] = 2; S[sm.o_luu + lane * m + lane] = (real)2 * a.beta;
  }
  __syncwarp();
  for (;;) {
    unsigned long long first = 0;
    if (lane == 0) first = atomicAdd(a.counter, 1ull);
    first = __shfl_sync(FULL, first, 0);
    if ((int64_t)first >= a.B) break;
    const int64_t b = (int64_t)first;
    real *Xb[2] = {a.states + b * (T + 1) * n, a.ws + b * per};
    real *Ub[2] = {a.actions + b * T * n, a.ws + b * per + (int64_t)(T + 1) * n};
    real *kb = a.ws + b * per + (int64_t)(2 * T + 1) * n, *Kb = kb + (int64_t)T * n;
    // start (ilqr.py:53-82)
    {
      real x = act ? a.x0[b * n + lane] : (real)0;
      if (act) Xb[0][lane] = x;
      for (int t = 0; t < T; t++) {
        const real u = act ? a.u_init[(b * T + t) * n + lane] : (real)0;
        x = x + u;
        if (act) { Ub[0][t * n + lane] = u; Xb[0][(t + 1) * n + lane] = x; }
      }
    }
    __syncwarp();
    double mu = 0.0, delta = 1.0;
    int cur = 0, n_bwd = 0, n_fwd = 0, status = TFMPC_ST_MAXITER, iteration = 0, guard = 0;
    bool done = false;
    while (!done) {
      // ---- _backward (:285-315): retry with a larger local mu while the Cholesky fails
      real J_hat = 0, dV1 = 0, dV2 = 0, gsum = 0;
      double mu_l = mu, delta_l = delta;
      int bst = 0, tries = 0;
      for (;;) {
        const real xT = act ? Xb[cur][T * n + lane] : (real)0;
        for (int i = lane; i < NP * LD; i += 32) sm.Vxx[i] = 0;
        __syncwarp();
        if (act) sm.Vxx[lane * LD + lane] = 2;  // final l_xx = 2I (diffenv.py:85-101)
        __syncwarp();
        real Vx = act ? (real)2 * (xT - a.goal[lane]) : (real)0;
        J_hat = navlqr_cost(a, xT, 0, true, lane);
        dV1 = 0; dV2 = 0; gsum = 0; bst = 0;
        for (int t = T - 1; t >= 0; t--) {
          const real x = act ? Xb[cur][t * n + lane] : (real)0, u = act ? Ub[cur][t * n + lane] : (real)0;
          if (act) { S[sm.o_lx + lane] = (real)2 * (x - a.goal[lane]); S[sm.o_lu + lane] = (real)2 * a.beta * u; }
          __syncwarp();
          const real l_t = navlqr_cost(a, x, u, false, lane);
          real kk;
          const int st = dense_step(sm, S, n, m, a.bounded, (real)mu_l, a.low, a.high, u, l_t, Vx, J_hat, dV1, dV2, kk, lane);
          if (st == 1) { bst = 1; break; }
          if (st) bst = st;
          gsum += warp_max(act ? r_abs(kk) / (r_abs(u) + (real)1.0) : (real)0);  // :243
          if (act) kb[t * n + lane] = kk;
          for (int i = lane; i < n * n; i += 32) Kb[(int64_t)t * n * n + i] = sm.Km[(i / n) * LD + (i % n)];
          __syncwarp();
        }
        n_bwd++;
        if (bst != 1 || ++tries > 200) break;
        delta_l = fmax(a.o.delta_0, delta_l * a.o.delta_0);
        mu_l = fmax(a.o.mu_min, mu_l * delta_l);
      }
      if (bst) { status = TFMPC_ST_NONPD; break; }
      const real g = gsum / (real)T;
      if (!(g == g)) { status = TFMPC_ST_NAN; break; }
      if (g < a.o.atol) { status = TFMPC_ST_CONVERGED; break; }  // :245-248
      // ---- _forward (:317-355): first-accept backtracking
      bool accept = false;
      real residual = 0;
      int rollouts = 0;
      for (int ai = 0; ai < N_ALPHA && !accept; ai++) {
        const real alpha = a.o.alphas[ai];
        real x = act ? Xb[cur][lane] : (real)0, J = 0, res = 0;
        if (act) Xb[cur ^ 1][lane] = x;
        for (int t = 0; t < T; t++) {
          const real dx = act ? x - Xb[cur][t * n + lane] : (real)0;
          real s = 0;
          for (int j = 0; j < n; j++) s += (act ? Kb[(int64_t)t * n * n + lane * n + j] : (real)0) * __shfl_sync(FULL, dx, j);
          const real du = act ? alpha * kb[t * n + lane] + s : (real)0;              // :194
          const real u = act ? r_clip(Ub[cur][t * n + lane] + du, a.low[lane], a.high[lane]) : (real)0;  // :196-197
          res = r_max(res, r_abs(du));                                             // :206
          J += navlqr_cost(a, x, u, false, lane);
          x = x + u;
          if (act) { Ub[cur ^ 1][t * n + lane] = u; Xb[cur ^ 1][(t + 1) * n + lane] = x; }
        }
        J += navlqr_cost(a, x, 0, true, lane);
        residual = warp_max(res);
        rollouts++;
        accept = ls_accepts(a.o, alpha, J_hat, dV1, dV2, J);
        __syncwarp();
      }
      n_fwd += rollouts;
      // ---- solve's bookkeeping (:253-270), as tick_finish() in small_core.cuh
      if (residual < a.o.atol) { status = TFMPC_ST_CONVERGED; cur ^= 1; break; }
      if (accept) {
        delta = fmin(1.0 / a.o.delta_0, delta / a.o.delta_0);
        mu = mu * delta * (double)(mu * delta > a.o.mu_min);
        guard = 0;
        cur ^= 1;
        if (iteration + 1 >= a.o.max_iterations) { status = TFMPC_ST_MAXITER; break; }
        iteration++;
      } else {
        delta = fmax(a.o.delta_0, delta * a.o.delta_0);
        mu = fmax(a.o.mu_min, mu * delta);
        if (++guard > 200) { status = TFMPC_ST_REGLOOP; break; }
      }
    }
    // results
    if (cur == 1 && act) {
      for (int t = 0; t <= T; t++) Xb[0][t * n + lane] = Xb[1][t * n + lane];
      for (int t = 0; t < T; t++) Ub[0][t * n + lane] = Ub[1][t * n + lane];
    }
    __syncwarp();
    for (int t = 0; t <= T; t++) {
      const real x = act ? Xb[0][t * n + lane] : (real)0, u = (act && t < T) ? Ub[0][t * n + lane] : (real)0;
      const real c = navlqr_cost(a, x, u, t == T, lane);
      if (lane == 0) a.costs[b * (T + 1) + t] = c;
    }
    if (lane == 0) { a.stats[b * 4] = iteration; a.stats[b * 4 + 1] = n_bwd; a.stats[b * 4 + 2] = n_fwd; a.stats[b * 4 + 3] = status; }
    __syncwarp();
  }
}

// iLQR.forward (ilqr.py:174-212) / iLQR.start (K == nullptr: plain rollout of `actions`) for NavigationLQR of any n
__global__ void __launch_bounds__(128) k_forward_dense_navlqr(DenseSolveArgs a, const real *__restrict__ xh, const real *__restrict__ uh,
                                                              const real *__restrict__ K, const real *__restrict__ k, real alpha,
                                                              real *__restrict__ xs, real *__restrict__ us, real *__restrict__ cs,
                                                              real *__restrict__ Jo, real *__restrict__ reso) {
  const int lane = threadIdx.x & 31, n = a.n, T = a.T;
  const bool act = lane < n;
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t b = w; b < a.B; b += nw) {
    real x = act ? (K ? xh[b * (T + 1) * n + lane] : a.x0[b * n + lane]) : (real)0, J = 0, res = 0;
    if (act) xs[b * (T + 1) * n + lane] = x;
    for (int t = 0; t < T; t++) {
      real u;
      if (K) {
        const real dx = act ? x - xh[(b * (T + 1) + t) * n + lane] : (real)0;
        real s = 0;
        for (int j = 0; j < n; j++) s += (act ? K[((b * T + t) * n + lane) * n + j] : (real)0) * __shfl_sync(FULL, dx, j);
        const real du = act ? alpha * k[(b * T + t) * n + lane] + s : (real)0;
        u = act ? r_clip(uh[(b * T + t) * n + lane] + du, a.low[lane], a.high[lane]) : (real)0;
        res = r_max(res, r_abs(du));
      } else {
        u = act ? uh[(b * T + t) * n + lane] : (real)0;
      }
      const real c = navlqr_cost(a, x, u, false, lane);
      J += c;
      x = x + u;
      if (act) { us[(b * T + t) * n + lane] = u; xs[(b * (T + 1) + t + 1) * n + lane] = x; }
      if (lane == 0) cs[b * (T + 1) + t] = c;
    }
    const real cf = navlqr_cost(a, x, 0, true, lane);
    J += cf;
    res = warp_max(res);
    if (lane == 0) { cs[b * (T + 1) + T] = cf; if (Jo) Jo[b] = J; if (reso) reso[b] = res; }
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
  const size_t smem = dense_smem_bytes(n, m);
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

int64_t dense_navlqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T) {
  return 256 + ((int64_t)(3 * T + 1) * e->n + (int64_t)T * e->n * e->n) * B * (int64_t)sizeof(real);
}

static void fill_navlqr_args(DenseSolveArgs &a, const tfmpc_env *e, int64_t B, int T);

int dense_navlqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init,
                       const IlqrOpts &o, real *states, real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes,
                       cudaStream_t s) {
  if (ws_bytes < dense_navlqr_workspace_bytes(e, B, T)) return tfmpc_set_error(TFMPC_E_WORKSPACE, "workspace too small");
  DenseSolveArgs a;
  memset(&a, 0, sizeof(a));
  fill_navlqr_args(a, e, B, T);
  a.o = o;
  a.x0 = x0; a.u_init = u_init; a.states = states; a.actions = actions; a.costs = costs; a.stats = stats;
  a.counter = (unsigned long long *)ws;
  a.ws = (real *)((char *)ws + 256);
  CUDA_TRY(cudaMemsetAsync(a.counter, 0, 256, s));
  const size_t smem = dense_smem_bytes(e->n, e->n);
  CUDA_TRY(cudaFuncSetAttribute(k_solve_dense_navlqr, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int sms = 148;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, e->device);
  const unsigned grid = (unsigned)std::min<int64_t>(B, (int64_t)sms * 2);
  k_solve_dense_navlqr<<<grid, 32, smem, s>>>(a);
  LAUNCH_CHECK();
  return TFMPC_OK;
}

static void fill_navlqr_args(DenseSolveArgs &a, const tfmpc_env *e, int64_t B, int T) {
  a.B = B; a.T = T; a.n = e->n; a.bounded = e->bounded; a.beta = (real)e->beta;
  for (int i = 0; i < NP; i++) {
    a.goal[i] = i < e->n ? (real)e->goal[i] : (real)0;
    a.low[i] = i < e->n ? (real)e->low[i] : (real)0;
    a.high[i] = i < e->n ? (real)e->high[i] : (real)0;
  }
}

int dense_navlqr_forward(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *xh, const real *uh, const real *K, const real *k,
                         double alpha, real *xs, real *us, real *cs, real *J, real *residual, cudaStream_t s) {
  DenseSolveArgs a;
  memset(&a, 0, sizeof(a));
  fill_navlqr_args(a, e, B, T);
  a.x0 = x0;
  const unsigned grid = (unsigned)std::min<int64_t>((B + 3) / 4, 148 * 16);
  k_forward_dense_navlqr<<<grid, 128, 0, s>>>(a, xh, uh, K, k, (real)alpha, xs, us, cs, J, residual);
  LAUNCH_CHECK();
  return TFMPC_OK;
}
