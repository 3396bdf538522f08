// Batched finite-horizon LQR: LQR.backward + LQR.forward of the reference
// (tfmpc/solvers/lqr.py:59-161) for B independent problems in one launch.
//
// Two kernels:
//  * thread-per-problem, (n, m) compile-time, everything in registers -- the shapes of the
//    BASELINE configs C1 (3,2) and C2 (2,2).  F, f, C (and c) may be shared across the batch
//    (stride 0): all lanes then read the same address, which the LSU broadcasts.
//  * warp-per-problem with the matrices in shared memory for any n, m <= 32.
#include "common.cuh"

namespace {

constexpr int kThreads = 128;
constexpr unsigned FULL = 0xffffffffu;

// general inverse with partial pivoting (tf.linalg.inv is LU-based, lqr.py:84); static indexing only
template <int M>
__device__ __forceinline__ int inverse_small(const real *A, real *inv) {
  real w[M][2 * M];
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < M; j++) { w[i][j] = A[i * M + j]; w[i][M + j] = (i == j) ? (real)1 : (real)0; }
  int fail = 0;
#pragma unroll
  for (int k = 0; k < M; k++) {
    int p = k;
    real best = r_abs(w[k][k]);
#pragma unroll
    for (int i = k + 1; i < M; i++) { real v = r_abs(w[i][k]); if (v > best) { best = v; p = i; } }
    if (best == 0) { fail = 1; best = 1; }
#pragma unroll
    for (int i = k + 1; i < M; i++)
      if (p == i) {
#pragma unroll
        for (int j = 0; j < 2 * M; j++) { real t = w[k][j]; w[k][j] = w[i][j]; w[i][j] = t; }
      }
    real piv = fail ? (real)1 : w[k][k];
#pragma unroll
    for (int j = 0; j < 2 * M; j++) w[k][j] /= piv;
#pragma unroll
    for (int i = 0; i < M; i++)
      if (i != k) {
        real f = w[i][k];
#pragma unroll
        for (int j = 0; j < 2 * M; j++) w[i][j] -= f * w[k][j];
      }
  }
#pragma unroll
  for (int i = 0; i < M; i++)
#pragma unroll
    for (int j = 0; j < M; j++) inv[i * M + j] = w[i][M + j];
  return fail;
}

// shared -> global bulk store of `bytes` (multiple of 16, both addresses 16-byte aligned), issued by one lane
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"((unsigned)__cvta_generic_to_shared(ssrc)), "r"(bytes)
               : "memory");
}

// ---- output sinks of the thread-per-problem solve
// DirectOut: every thread writes its own rows of the reference-layout outputs [B, T, ...] (a problem's row is contiguous,
//            neighbouring threads are a whole row apart -> scattered 4-byte stores; the fallback for long rows).
// TileOut:   the 32 problems of a warp own one CONTIGUOUS block of every output array (32 consecutive rows), so the warp
//            assembles each block in shared memory and one lane writes it with a TMA bulk store (cp.async.bulk
//            shared -> global): full-line HBM writes, no per-element store instructions, and the gains never leave
//            the SM between the backward sweep and the rollout.
template <int N, int M>
struct DirectOut {
  real *states, *actions, *costs, *Ko, *ko, *Vo, *vo, *csto, *scratch;
  int64_t b, S;
  int T;
  __device__ __forceinline__ void put_gain(int t, const real *K, const real *k) const {
    if (Ko == nullptr) {
#pragma unroll
      for (int i = 0; i < M * N; i++) scratch[((int64_t)t * (M * N + M) + i) * S + b] = K[i];
#pragma unroll
      for (int i = 0; i < M; i++) scratch[((int64_t)t * (M * N + M) + M * N + i) * S + b] = k[i];
    } else {
#pragma unroll
      for (int i = 0; i < M * N; i++) Ko[(b * T + t) * M * N + i] = K[i];
#pragma unroll
      for (int i = 0; i < M; i++) ko[(b * T + t) * M + i] = k[i];
    }
  }
  __device__ __forceinline__ void get_gain(int t, real *K, real *k) const {
    if (Ko == nullptr) {
#pragma unroll
      for (int i = 0; i < M * N; i++) K[i] = scratch[((int64_t)t * (M * N + M) + i) * S + b];
#pragma unroll
      for (int i = 0; i < M; i++) k[i] = scratch[((int64_t)t * (M * N + M) + M * N + i) * S + b];
    } else {
#pragma unroll
      for (int i = 0; i < M * N; i++) K[i] = Ko[(b * T + t) * M * N + i];
#pragma unroll
      for (int i = 0; i < M; i++) k[i] = ko[(b * T + t) * M + i];
    }
  }
  __device__ __forceinline__ void put_value(int t, const real *V, const real *v, real cst) const {
    if (Vo) {
#pragma unroll
      for (int i = 0; i < N * N; i++) Vo[(b * T + t) * N * N + i] = V[i];
#pragma unroll
      for (int i = 0; i < N; i++) vo[(b * T + t) * N + i] = v[i];
      csto[b * T + t] = cst;
    }
  }
  __device__ __forceinline__ void put_state(int t, const real *x) const {
#pragma unroll
    for (int i = 0; i < N; i++) states[(b * (T + 1) + t) * N + i] = x[i];
  }
  __device__ __forceinline__ void put_action(int t, const real *u) const {
#pragma unroll
    for (int i = 0; i < M; i++) actions[(b * T + t) * M + i] = u[i];
  }
  __device__ __forceinline__ void put_cost(int t, real c) const { costs[b * (T + 1) + t] = c; }
  __device__ __forceinline__ void between() const {}
};

template <int N, int M>
struct TileOut {
  // Per-warp tiles, row = lane.  Region A holds the gains (K, k) for the whole kernel (the rollout reads them back);
  // region B holds V, v, const during the backward sweep and -- once those have been written out -- states, actions,
  // costs during the rollout.  130 reals per row for n = m = 2, T = 10 (16.6 KB per warp -> 12 warps per SM).
  real *base;
  real *Kt, *kt, *Vt, *vt, *ct, *st, *ac, *co;   // this lane's rows
  real *gK, *gk, *gV, *gv, *gc, *gs, *ga, *gco;  // the warp's blocks in the global outputs (nullptr = not wanted)
  int T, lane, nvalid;
  static __host__ __device__ int64_t row_reals(int T) {
    const int64_t a = (int64_t)T * (M * N + M), b1 = (int64_t)T * (N * N + N + 1), b2 = (int64_t)(T + 1) * N + (int64_t)T * M + (T + 1);
    return a + (b1 > b2 ? b1 : b2);
  }
  __device__ __forceinline__ TileOut(real *base_, int lane_, int nvalid_, int T_, int64_t b0, real *states, real *actions, real *costs, real *Ko,
                                     real *ko, real *Vo, real *vo, real *csto)
      : base(base_), T(T_), lane(lane_), nvalid(nvalid_) {
    real *p = base;
    Kt = p + lane * T * M * N; p += 32 * T * M * N;
    kt = p + lane * T * M; p += 32 * T * M;
    real *regionB = p;
    Vt = p + lane * T * N * N; p += 32 * T * N * N;
    vt = p + lane * T * N; p += 32 * T * N;
    ct = p + lane * T;
    p = regionB;
    st = p + lane * (T + 1) * N; p += 32 * (T + 1) * N;
    ac = p + lane * T * M; p += 32 * T * M;
    co = p + lane * (T + 1);
    gK = Ko ? Ko + b0 * T * M * N : nullptr; gk = ko ? ko + b0 * T * M : nullptr;
    gV = Vo ? Vo + b0 * T * N * N : nullptr; gv = vo ? vo + b0 * T * N : nullptr; gc = csto ? csto + b0 * T : nullptr;
    gs = states + b0 * (T + 1) * N; ga = actions + b0 * T * M; gco = costs + b0 * (T + 1);
  }
  __device__ __forceinline__ void put_gain(int t, const real *K, const real *k) const {
#pragma unroll
    for (int i = 0; i < M * N; i++) Kt[t * M * N + i] = K[i];
#pragma unroll
    for (int i = 0; i < M; i++) kt[t * M + i] = k[i];
  }
  __device__ __forceinline__ void get_gain(int t, real *K, real *k) const {
#pragma unroll
    for (int i = 0; i < M * N; i++) K[i] = Kt[t * M * N + i];
#pragma unroll
    for (int i = 0; i < M; i++) k[i] = kt[t * M + i];
  }
  __device__ __forceinline__ void put_value(int t, const real *V, const real *v, real cst) const {
    if (gV) {
#pragma unroll
      for (int i = 0; i < N * N; i++) Vt[t * N * N + i] = V[i];
#pragma unroll
      for (int i = 0; i < N; i++) vt[t * N + i] = v[i];
      ct[t] = cst;
    }
  }
  __device__ __forceinline__ void put_state(int t, const real *x) const {
#pragma unroll
    for (int i = 0; i < N; i++) st[t * N + i] = x[i];
  }
  __device__ __forceinline__ void put_action(int t, const real *u) const {
#pragma unroll
    for (int i = 0; i < M; i++) ac[t * M + i] = u[i];
  }
  __device__ __forceinline__ void put_cost(int t, real c) const { co[t] = c; }

  // the warp's `nvalid` rows of one array (row length rl) are one contiguous block in global memory
  __device__ __forceinline__ void store_block(real *g, const real *lane_row, int rl) const {
    if (g == nullptr) return;
    const real *tile = lane_row - lane * rl;
    const unsigned bytes = (unsigned)(nvalid * rl * (int)sizeof(real));
    if ((bytes & 15u) == 0) {
      if (lane == 0) bulk_store(g, tile, bytes);
    } else {  // ragged last warp: coalesced element copies
      for (int i = lane; i < nvalid * rl; i += 32) g[i] = tile[i];
    }
  }
  __device__ __forceinline__ void drain() const {   // the tiles must outlive the copies that read them
    if (lane == 0) {
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
    __syncwarp();
  }
  // called by all 32 lanes between the backward sweep and the rollout: write out policy and value function, free region B
  __device__ __forceinline__ void between() const {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy tile writes -> visible to the bulk-copy engine
    __syncwarp();
    store_block(gK, Kt, T * M * N); store_block(gk, kt, T * M);
    store_block(gV, Vt, T * N * N); store_block(gv, vt, T * N); store_block(gc, ct, T);
    drain();
  }
  __device__ __forceinline__ void finish() const {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    store_block(gs, st, (T + 1) * N); store_block(ga, ac, T * M); store_block(gco, co, T + 1);
    drain();
  }
};

// LQR.backward (lqr.py:59-129) + LQR.forward (:131-161) of problem b, all matrices in registers
template <int N, int M, class OUT>
__device__ __forceinline__ int lqr_small_solve(int64_t b, int T, const real *__restrict__ Fp, int64_t sF, const real *__restrict__ fp,
                                               int64_t sf, const real *__restrict__ Cp, int64_t sC, const real *__restrict__ cp, int64_t sc,
                                               const real *__restrict__ x0, int terminal_zero, const OUT &out) {
  constexpr int NM = N + M;
  real F[N * NM], f[N], C[NM * NM], c[NM];
#pragma unroll
  for (int i = 0; i < N * NM; i++) F[i] = Fp[b * sF + i];
#pragma unroll
  for (int i = 0; i < N; i++) f[i] = fp[b * sf + i];
#pragma unroll
  for (int i = 0; i < NM * NM; i++) C[i] = Cp[b * sC + i];
#pragma unroll
  for (int i = 0; i < NM; i++) c[i] = cp[b * sc + i];
  real V[N * N], v[N], cst = 0;
  int st = 0;
#pragma unroll
  for (int i = 0; i < N; i++) {  // lqr.py:67-68 (or the README's V_T = v_T = 0)
    v[i] = terminal_zero ? (real)0 : c[i];
#pragma unroll
    for (int j = 0; j < N; j++) V[i * N + j] = terminal_zero ? (real)0 : C[i * NM + j];
  }
  for (int t = T - 1; t >= 0; t--) {
    real FtV[NM * N], Q[NM * NM], q[NM];
#pragma unroll
    for (int i = 0; i < NM; i++)
#pragma unroll
      for (int j = 0; j < N; j++) {  // :74
        real s = 0;
#pragma unroll
        for (int p = 0; p < N; p++) s += F[p * NM + i] * V[p * N + j];
        FtV[i * N + j] = s;
      }
#pragma unroll
    for (int i = 0; i < NM; i++) {
#pragma unroll
      for (int j = 0; j < NM; j++) {  // :75
        real s = 0;
#pragma unroll
        for (int p = 0; p < N; p++) s += FtV[i * N + p] * F[p * NM + j];
        Q[i * NM + j] = C[i * NM + j] + s;
      }
      real s1 = 0, s2 = 0;  // :76-78
#pragma unroll
      for (int p = 0; p < N; p++) { s1 += FtV[i * N + p] * f[p]; s2 += F[p * NM + i] * v[p]; }
      q[i] = c[i] + s1 + s2;
    }
    real Quu[M * M], inv[M * M], K[M * N], k[M];
#pragma unroll
    for (int i = 0; i < M; i++)
#pragma unroll
      for (int j = 0; j < M; j++) Quu[i * M + j] = Q[(N + i) * NM + N + j];
    if (inverse_small<M>(Quu, inv)) st = TFMPC_ST_NONPD;  // :84
#pragma unroll
    for (int i = 0; i < M; i++) {  // :86-87
#pragma unroll
      for (int j = 0; j < N; j++) {
        real s = 0;
#pragma unroll
        for (int p = 0; p < M; p++) s += inv[i * M + p] * Q[(N + p) * NM + j];
        K[i * N + j] = -s;
      }
      real s = 0;
#pragma unroll
      for (int p = 0; p < M; p++) s += inv[i * M + p] * q[N + p];
      k[i] = -s;
    }
    // const recursion with W == Q, w == q built from the OLD V, v  (:107-121)
    real c1 = 0, c2 = 0, c3a = 0, c3b = 0;
#pragma unroll
    for (int i = 0; i < M; i++) {
      real s = 0;
#pragma unroll
      for (int p = 0; p < M; p++) s += Quu[i * M + p] * k[p];
      c1 += k[i] * s;
      c2 += k[i] * q[N + i];
    }
#pragma unroll
    for (int i = 0; i < N; i++) {
      real s = 0;
#pragma unroll
      for (int p = 0; p < N; p++) s += V[i * N + p] * f[p];
      c3a += f[i] * s;
      c3b += f[i] * v[i];
    }
    cst += ((real)0.5 * c1 + c2 + ((real)0.5 * c3a + c3b));
    // V, v update (:97-105)
    real KtQuu[N * M], Vn[N * N], vn[N];
#pragma unroll
    for (int i = 0; i < N; i++)
#pragma unroll
      for (int j = 0; j < M; j++) {
        real s = 0;
#pragma unroll
        for (int p = 0; p < M; p++) s += K[p * N + i] * Quu[p * M + j];
        KtQuu[i * M + j] = s;
      }
#pragma unroll
    for (int i = 0; i < N; i++) {
#pragma unroll
      for (int j = 0; j < N; j++) {
        real s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
        for (int p = 0; p < M; p++) { s1 += Q[i * NM + N + p] * K[p * N + j]; s2 += K[p * N + i] * Q[(N + p) * NM + j]; s3 += KtQuu[i * M + p] * K[p * N + j]; }
        Vn[i * N + j] = Q[i * NM + j] + s1 + s2 + s3;
      }
      real s1 = 0, s2 = 0, s3 = 0;
#pragma unroll
      for (int p = 0; p < M; p++) { s1 += Q[i * NM + N + p] * k[p]; s2 += K[p * N + i] * q[N + p]; s3 += KtQuu[i * M + p] * k[p]; }
      vn[i] = q[i] + s1 + s2 + s3;
    }
#pragma unroll
    for (int i = 0; i < N * N; i++) V[i] = Vn[i];
#pragma unroll
    for (int i = 0; i < N; i++) v[i] = vn[i];
    out.put_gain(t, K, k);
    out.put_value(t, V, v, cst);
  }
  out.between();
  // forward, :131-161
  real x[N], z[NM];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = x0[b * N + i];
  out.put_state(0, x);
  for (int t = 0; t < T; t++) {
    real K[M * N], k[M];
    out.get_gain(t, K, k);
#pragma unroll
    for (int i = 0; i < N; i++) z[i] = x[i];
#pragma unroll
    for (int i = 0; i < M; i++) {  // :143
      real s = 0;
#pragma unroll
      for (int p = 0; p < N; p++) s += K[i * N + p] * x[p];
      z[N + i] = s + k[i];
    }
    out.put_action(t, z + N);
    real quad = 0, lin = 0;  // cost, :41-47
#pragma unroll
    for (int j = 0; j < NM; j++) {
      real s = 0;
#pragma unroll
      for (int p = 0; p < NM; p++) s += z[p] * C[p * NM + j];
      quad += s * z[j];
      lin += z[j] * c[j];
    }
    out.put_cost(t, (real)0.5 * quad + lin);
#pragma unroll
    for (int i = 0; i < N; i++) {  // transition, :36-39
      real s = 0;
#pragma unroll
      for (int p = 0; p < NM; p++) s += F[i * NM + p] * z[p];
      x[i] = s + f[i];
    }
    out.put_state(t + 1, x);
  }
  real quad = 0, lin = 0;  // final_cost, :49-57
#pragma unroll
  for (int j = 0; j < N; j++) {
    real s = 0;
#pragma unroll
    for (int p = 0; p < N; p++) s += x[p] * C[p * NM + j];
    quad += s * x[j];
    lin += x[j] * c[j];
  }
  out.put_cost(T, terminal_zero ? (real)0 : (real)0.5 * quad + lin);
  return st;
}

template <int N, int M>
__global__ void __launch_bounds__(kThreads) k_lqr_small(int64_t B, int T, const real *__restrict__ Fp, int64_t sF, const real *__restrict__ fp,
                                                        int64_t sf, const real *__restrict__ Cp, int64_t sC, const real *__restrict__ cp,
                                                        int64_t sc, const real *__restrict__ x0, int terminal_zero, real *__restrict__ states,
                                                        real *__restrict__ actions, real *__restrict__ costs, real *__restrict__ Ko,
                                                        real *__restrict__ ko, real *__restrict__ Vo, real *__restrict__ vo,
                                                        real *__restrict__ csto, int32_t *__restrict__ status, real *__restrict__ scratch) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  // policy storage: K, k per timestep.  Caller-provided outputs when present, else scratch
  // (struct-of-arrays, problem index fastest: coalesced).
  const DirectOut<N, M> out = {states, actions, costs, Ko, ko, Vo, vo, csto, scratch, b, (B + 31) / 32 * 32, T};
  const int st = lqr_small_solve<N, M>(b, T, Fp, sF, fp, sf, Cp, sC, cp, sc, x0, terminal_zero, out);
  if (status) status[b] = st;
}

constexpr int kStagedThreads = 64;  // 2 warps x 16.6 KB of tiles (C2: T = 10) -> 6 CTAs = 12 warps per SM

template <int N, int M>
__global__ void __launch_bounds__(kStagedThreads) k_lqr_small_staged(int64_t B, int T, const real *__restrict__ Fp, int64_t sF,
                                                                      const real *__restrict__ fp, int64_t sf, const real *__restrict__ Cp,
                                                                      int64_t sC, const real *__restrict__ cp, int64_t sc,
                                                                      const real *__restrict__ x0, int terminal_zero, real *__restrict__ states,
                                                                      real *__restrict__ actions, real *__restrict__ costs, real *__restrict__ Ko,
                                                                      real *__restrict__ ko, real *__restrict__ Vo, real *__restrict__ vo,
                                                                      real *__restrict__ csto, int32_t *__restrict__ status) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  real *base = reinterpret_cast<real *>(smem_raw) + (int64_t)wq * 32 * TileOut<N, M>::row_reals(T);
  const int64_t b0 = ((int64_t)blockIdx.x * (kStagedThreads / 32) + wq) * 32;
  if (b0 >= B) return;
  const int nvalid = (int)min((int64_t)32, B - b0);
  const int64_t b = b0 + min(lane, nvalid - 1);   // padding lanes of the last warp shadow its last problem (their rows are never stored)
  const TileOut<N, M> out(base, lane, nvalid, T, b0, states, actions, costs, Ko, ko, Vo, vo, csto);
  const int st = lqr_small_solve<N, M>(b, T, Fp, sF, fp, sf, Cp, sC, cp, sc, x0, terminal_zero, out);
  if (status && lane < nvalid) status[b] = st;
  out.finish();
}

// ---------------------------------------------------------------- generic warp-per-problem kernel
// Matrices live in shared memory (one warp per block); lanes split the output elements of every
// product.  The m x m inverse is Gauss-Jordan with partial pivoting on an [m][2m] tableau.
__global__ void __launch_bounds__(32) k_lqr_warp(int64_t B, int n, int m, int T, const real *__restrict__ Fp, int64_t sF,
                                                  const real *__restrict__ fp, int64_t sf, const real *__restrict__ Cp, int64_t sC,
                                                  const real *__restrict__ cp, int64_t sc, const real *__restrict__ x0, int terminal_zero,
                                                  real *__restrict__ states, real *__restrict__ actions, real *__restrict__ costs,
                                                  real *__restrict__ Ko, real *__restrict__ ko, real *__restrict__ Vo, real *__restrict__ vo,
                                                  real *__restrict__ csto, int32_t *__restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  real *sm = reinterpret_cast<real *>(smem_raw);
  const int N = n + m, lane = threadIdx.x;
  real *F = sm; real *f = F + n * N; real *C = f + n; real *c = C + N * N; real *V = c + N; real *v = V + n * n;
  real *FtV = v + n; real *Q = FtV + N * n; real *q = Q + N * N; real *W = q + N; real *K = W + 2 * m * m; real *k = K + m * n;
  real *KtQuu = k + m; real *Vn = KtQuu + n * m; real *vn = Vn + n * n; real *z = vn + n; real *x = z + N;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    for (int i = lane; i < n * N; i += 32) F[i] = Fp[b * sF + i];
    for (int i = lane; i < n; i += 32) f[i] = fp[b * sf + i];
    for (int i = lane; i < N * N; i += 32) C[i] = Cp[b * sC + i];
    for (int i = lane; i < N; i += 32) c[i] = cp[b * sc + i];
    __syncwarp();
    for (int i = lane; i < n * n; i += 32) V[i] = terminal_zero ? (real)0 : C[(i / n) * N + (i % n)];
    for (int i = lane; i < n; i += 32) v[i] = terminal_zero ? (real)0 : c[i];
    real cst = 0;
    int st = 0;
    __syncwarp();
    for (int t = T - 1; t >= 0; t--) {
      for (int e = lane; e < N * n; e += 32) {
        int i = e / n, j = e % n;
        real s = 0;
        for (int p = 0; p < n; p++) s += F[p * N + i] * V[p * n + j];
        FtV[e] = s;
      }
      __syncwarp();
      for (int e = lane; e < N * N; e += 32) {
        int i = e / N, j = e % N;
        real s = 0;
        for (int p = 0; p < n; p++) s += FtV[i * n + p] * F[p * N + j];
        Q[e] = C[e] + s;
      }
      for (int i = lane; i < N; i += 32) {
        real s1 = 0, s2 = 0;
        for (int p = 0; p < n; p++) { s1 += FtV[i * n + p] * f[p]; s2 += F[p * N + i] * v[p]; }
        q[i] = c[i] + s1 + s2;
      }
      __syncwarp();
      // tableau W = [Q_uu | I]
      for (int e = lane; e < m * 2 * m; e += 32) {
        int i = e / (2 * m), j = e % (2 * m);
        W[e] = j < m ? Q[(n + i) * N + n + j] : ((j - m) == i ? (real)1 : (real)0);
      }
      __syncwarp();
      for (int kk = 0; kk < m; kk++) {
        int p = kk;
        real best = r_abs(W[kk * 2 * m + kk]);
        for (int i = kk + 1; i < m; i++) { real a = r_abs(W[i * 2 * m + kk]); if (a > best) { best = a; p = i; } }  // uniform across lanes
        if (best == 0) { st = TFMPC_ST_NONPD; best = 1; }
        if (p != kk) for (int j = lane; j < 2 * m; j += 32) { real tv = W[kk * 2 * m + j]; W[kk * 2 * m + j] = W[p * 2 * m + j]; W[p * 2 * m + j] = tv; }
        __syncwarp();
        real piv = W[kk * 2 * m + kk];
        if (piv == 0) piv = 1;
        __syncwarp();
        for (int j = lane; j < 2 * m; j += 32) W[kk * 2 * m + j] /= piv;
        __syncwarp();
        // eliminate row by row; each lane owns columns j = lane, lane+32, ...
        for (int i = 0; i < m; i++) {
          if (i == kk) continue;
          real fct = W[i * 2 * m + kk];
          __syncwarp();
          for (int j = lane; j < 2 * m; j += 32) W[i * 2 * m + j] -= fct * W[kk * 2 * m + j];
          __syncwarp();
        }
      }
      for (int e = lane; e < m * n; e += 32) {
        int i = e / n, j = e % n;
        real s = 0;
        for (int p = 0; p < m; p++) s += W[i * 2 * m + m + p] * Q[(n + p) * N + j];
        K[e] = -s;
      }
      for (int i = lane; i < m; i += 32) {
        real s = 0;
        for (int p = 0; p < m; p++) s += W[i * 2 * m + m + p] * q[n + p];
        k[i] = -s;
      }
      __syncwarp();
      {  // const recursion (every lane computes the same scalars)
        real c1 = 0, c2 = 0, c3a = 0, c3b = 0;
        for (int i = 0; i < m; i++) {
          real s = 0;
          for (int p = 0; p < m; p++) s += Q[(n + i) * N + n + p] * k[p];
          c1 += k[i] * s;
          c2 += k[i] * q[n + i];
        }
        for (int i = 0; i < n; i++) {
          real s = 0;
          for (int p = 0; p < n; p++) s += V[i * n + p] * f[p];
          c3a += f[i] * s;
          c3b += f[i] * v[i];
        }
        cst += ((real)0.5 * c1 + c2 + ((real)0.5 * c3a + c3b));
      }
      for (int e = lane; e < n * m; e += 32) {
        int i = e / m, j = e % m;
        real s = 0;
        for (int p = 0; p < m; p++) s += K[p * n + i] * Q[(n + p) * N + n + j];
        KtQuu[e] = s;
      }
      __syncwarp();
      for (int e = lane; e < n * n; e += 32) {
        int i = e / n, j = e % n;
        real s1 = 0, s2 = 0, s3 = 0;
        for (int p = 0; p < m; p++) { s1 += Q[i * N + n + p] * K[p * n + j]; s2 += K[p * n + i] * Q[(n + p) * N + j]; s3 += KtQuu[i * m + p] * K[p * n + j]; }
        Vn[e] = Q[i * N + j] + s1 + s2 + s3;
      }
      for (int i = lane; i < n; i += 32) {
        real s1 = 0, s2 = 0, s3 = 0;
        for (int p = 0; p < m; p++) { s1 += Q[i * N + n + p] * k[p]; s2 += K[p * n + i] * q[n + p]; s3 += KtQuu[i * m + p] * k[p]; }
        vn[i] = q[i] + s1 + s2 + s3;
      }
      __syncwarp();
      for (int i = lane; i < n * n; i += 32) { V[i] = Vn[i]; if (Vo) Vo[(b * T + t) * n * n + i] = Vn[i]; }
      for (int i = lane; i < n; i += 32) { v[i] = vn[i]; if (vo) vo[(b * T + t) * n + i] = vn[i]; }
      if (csto && lane == 0) csto[b * T + t] = cst;
      for (int i = lane; i < m * n; i += 32) Ko[(b * T + t) * m * n + i] = K[i];
      for (int i = lane; i < m; i += 32) ko[(b * T + t) * m + i] = k[i];
      __syncwarp();
    }
    // forward
    for (int i = lane; i < n; i += 32) { x[i] = x0[b * n + i]; states[b * (T + 1) * n + i] = x[i]; }
    __syncwarp();
    for (int t = 0; t < T; t++) {
      for (int i = lane; i < n; i += 32) z[i] = x[i];
      for (int i = lane; i < m; i += 32) {
        real s = 0;
        for (int p = 0; p < n; p++) s += Ko[(b * T + t) * m * n + i * n + p] * x[p];
        z[n + i] = s + ko[(b * T + t) * m + i];
        actions[(b * T + t) * m + i] = z[n + i];
      }
      __syncwarp();
      real quad = 0, lin = 0;  // every lane: same scalar
      for (int j = 0; j < N; j++) {
        real s = 0;
        for (int p = 0; p < N; p++) s += z[p] * C[p * N + j];
        quad += s * z[j];
        lin += z[j] * c[j];
      }
      if (lane == 0) costs[b * (T + 1) + t] = (real)0.5 * quad + lin;
      for (int i = lane; i < n; i += 32) {
        real s = 0;
        for (int p = 0; p < N; p++) s += F[i * N + p] * z[p];
        vn[i] = s + f[i];
      }
      __syncwarp();
      for (int i = lane; i < n; i += 32) { x[i] = vn[i]; states[(b * (T + 1) + t + 1) * n + i] = vn[i]; }
      __syncwarp();
    }
    {
      real quad = 0, lin = 0;
      for (int j = 0; j < n; j++) {
        real s = 0;
        for (int p = 0; p < n; p++) s += x[p] * C[p * N + j];
        quad += s * x[j];
        lin += x[j] * c[j];
      }
      if (lane == 0) { costs[b * (T + 1) + T] = terminal_zero ? (real)0 : (real)0.5 * quad + lin; if (status) status[b] = st; }
    }
    __syncwarp();
  }
}


// ---------------------------------------------------------------- rollout / single-step operators (any n, m <= 32)
// LQR.forward(policy, x0, T) (lqr.py:131-161) for a caller-supplied policy; thread per problem,
// runtime dimensions (z, x in local memory: this is an API-completeness path, not the hot one).
__global__ void __launch_bounds__(kThreads) k_lqr_forward(int64_t B, int n, int m, int T, const real *__restrict__ Fp, int64_t sF,
                                                          const real *__restrict__ fp, int64_t sf, const real *__restrict__ Cp, int64_t sC,
                                                          const real *__restrict__ cp, int64_t sc, const real *__restrict__ K,
                                                          const real *__restrict__ k, const real *__restrict__ x0, real *__restrict__ states,
                                                          real *__restrict__ actions, real *__restrict__ costs) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const int N = n + m;
  const real *F = Fp + b * sF, *f = fp + b * sf, *C = Cp + b * sC, *c = cp + b * sc;
  real z[2 * MAXD], xn[MAXD];
  for (int i = 0; i < n; i++) { z[i] = x0[b * n + i]; states[b * (T + 1) * n + i] = z[i]; }
  for (int t = 0; t < T; t++) {
    for (int i = 0; i < m; i++) {
      real s = 0;
      for (int p = 0; p < n; p++) s += K[((b * T + t) * m + i) * n + p] * z[p];
      z[n + i] = s + k[(b * T + t) * m + i];
      actions[(b * T + t) * m + i] = z[n + i];
    }
    real quad = 0, lin = 0;
    for (int j = 0; j < N; j++) {
      real s = 0;
      for (int p = 0; p < N; p++) s += z[p] * C[p * N + j];
      quad += s * z[j];
      lin += z[j] * c[j];
    }
    costs[b * (T + 1) + t] = (real)0.5 * quad + lin;
    for (int i = 0; i < n; i++) {
      real s = 0;
      for (int p = 0; p < N; p++) s += F[i * N + p] * z[p];
      xn[i] = s + f[i];
    }
    for (int i = 0; i < n; i++) { z[i] = xn[i]; states[(b * (T + 1) + t + 1) * n + i] = xn[i]; }
  }
  real quad = 0, lin = 0;
  for (int j = 0; j < n; j++) {
    real s = 0;
    for (int p = 0; p < n; p++) s += z[p] * C[p * N + j];
    quad += s * z[j];
    lin += z[j] * c[j];
  }
  costs[b * (T + 1) + T] = (real)0.5 * quad + lin;
}

// LQR.transition / cost / final_cost (lqr.py:36-57) for R rows; any output may be NULL
__global__ void __launch_bounds__(kThreads) k_lqr_step(int64_t R, int n, int m, const real *__restrict__ Fp, int64_t sF, const real *__restrict__ fp,
                                                       int64_t sf, const real *__restrict__ Cp, int64_t sC, const real *__restrict__ cp,
                                                       int64_t sc, const real *__restrict__ x, const real *__restrict__ u, real *__restrict__ xn,
                                                       real *__restrict__ cost, real *__restrict__ fcost) {
  int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= R) return;
  const int N = n + m;
  const real *F = Fp + r * sF, *f = fp + r * sf, *C = Cp + r * sC, *c = cp + r * sc;
  real z[2 * MAXD];
  for (int i = 0; i < n; i++) z[i] = x[r * n + i];
  for (int i = 0; i < m; i++) z[n + i] = u ? u[r * m + i] : (real)0;
  if (xn)
    for (int i = 0; i < n; i++) {
      real s = 0;
      for (int p = 0; p < N; p++) s += F[i * N + p] * z[p];
      xn[r * n + i] = s + f[i];
    }
  if (cost) {
    real quad = 0, lin = 0;
    for (int j = 0; j < N; j++) {
      real s = 0;
      for (int p = 0; p < N; p++) s += z[p] * C[p * N + j];
      quad += s * z[j];
      lin += z[j] * c[j];
    }
    cost[r] = (real)0.5 * quad + lin;
  }
  if (fcost) {
    real quad = 0, lin = 0;
    for (int j = 0; j < n; j++) {
      real s = 0;
      for (int p = 0; p < n; p++) s += z[p] * C[p * N + j];
      quad += s * z[j];
      lin += z[j] * c[j];
    }
    fcost[r] = (real)0.5 * quad + lin;
  }
}

}  // namespace

// scratch for the policy when the caller does not want K, k back: allocated stream-ordered
int lqr_solve_launch(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                     const real *c, int64_t sc, const real *x0, int terminal_zero, real *states, real *actions, real *costs, real *K, real *k,
                     real *V, real *v, real *cst, int32_t *status, cudaStream_t s) {
  if ((K == nullptr) != (k == nullptr)) return tfmpc_set_error(TFMPC_E_INVALID, "K and k must both be given or both be NULL");
  if (V && !(v && cst)) return tfmpc_set_error(TFMPC_E_INVALID, "V, v and cst must be given together");
  const bool small = (n == 2 && m == 2) || (n == 3 && m == 2);
  // staged variant: tiles of 32 rows per warp must fit shared memory and every output block must be 16-byte aligned
  const int64_t regA = (int64_t)T * (m * n + m), regB1 = (int64_t)T * (n * n + n + 1), regB2 = (int64_t)(T + 1) * n + (int64_t)T * m + (T + 1);
  const int64_t row_reals = regA + std::max(regB1, regB2);   // TileOut::row_reals
  const size_t staged_smem = (size_t)(kStagedThreads / 32) * 32 * row_reals * sizeof(real);
  auto aligned16 = [](const void *p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
  const bool staged = small && staged_smem <= 100 * 1024 && aligned16(states) && aligned16(actions) && aligned16(costs) && aligned16(K) &&
                      aligned16(k) && aligned16(V) && aligned16(v) && aligned16(cst);
  real *scratch = nullptr;
  if (K == nullptr && !staged) {
    int64_t S = (B + 31) / 32 * 32;
    int64_t elems = small ? (int64_t)T * (m * n + m) * S : (int64_t)B * T * (m * n + m);
    CUDA_TRY(cudaMallocAsync((void **)&scratch, (size_t)elems * sizeof(real), s));
  }
  if (staged) {
    unsigned grid = (unsigned)((B + kStagedThreads - 1) / kStagedThreads);
    if (n == 2) {
      cudaFuncSetAttribute(k_lqr_small_staged<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      k_lqr_small_staged<2, 2><<<grid, kStagedThreads, staged_smem, s>>>(B, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, K, k, V, v, cst, status);
    } else {
      cudaFuncSetAttribute(k_lqr_small_staged<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
      k_lqr_small_staged<3, 2><<<grid, kStagedThreads, staged_smem, s>>>(B, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, K, k, V, v, cst, status);
    }
  } else if (small) {
    unsigned grid = (unsigned)((B + kThreads - 1) / kThreads);
    if (n == 2) k_lqr_small<2, 2><<<grid, kThreads, 0, s>>>(B, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, K, k, V, v, cst, status, scratch);
    else k_lqr_small<3, 2><<<grid, kThreads, 0, s>>>(B, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, K, k, V, v, cst, status, scratch);
  } else {
    const int N = n + m;
    real *Kp = K ? K : scratch, *kp = K ? k : scratch + (int64_t)B * T * m * n;
    size_t smem = sizeof(real) * ((size_t)n * N + n + (size_t)N * N + N + (size_t)n * n + n + (size_t)N * n + (size_t)N * N + N + 2 * (size_t)m * m +
                                  (size_t)m * n + m + (size_t)n * m + (size_t)n * n + n + N + n);
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(k_lqr_warp, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
      attr_set = true;
    }
    unsigned grid = (unsigned)(B < 148 * 8 ? B : 148 * 8);
    k_lqr_warp<<<grid, 32, smem, s>>>(B, n, m, T, F, sF, f, sf, C, sC, c, sc, x0, terminal_zero, states, actions, costs, Kp, kp, V, v, cst, status);
  }
  tfmpc_count_launch(1);
  cudaError_t le = cudaGetLastError();
  if (scratch) cudaFreeAsync(scratch, s);
  if (le != cudaSuccess) return tfmpc_set_error(TFMPC_E_CUDA, "LQR kernel launch: %s", cudaGetErrorString(le));
  return TFMPC_OK;
}

int lqr_forward_launch(int64_t B, int n, int m, int T, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC,
                       const real *c, int64_t sc, const real *K, const real *k, const real *x0, real *states, real *actions, real *costs,
                       cudaStream_t s) {
  k_lqr_forward<<<(unsigned)((B + kThreads - 1) / kThreads), kThreads, 0, s>>>(B, n, m, T, F, sF, f, sf, C, sC, c, sc, K, k, x0, states, actions, costs);
  LAUNCH_CHECK();
  return TFMPC_OK;
}

int lqr_step_launch(int64_t R, int n, int m, const real *F, int64_t sF, const real *f, int64_t sf, const real *C, int64_t sC, const real *c,
                    int64_t sc, const real *x, const real *u, real *xn, real *cost, real *fcost, cudaStream_t s) {
  k_lqr_step<<<(unsigned)((R + kThreads - 1) / kThreads), kThreads, 0, s>>>(R, n, m, F, sF, f, sf, C, sC, c, sc, x, u, xn, cost, fcost);
  LAUNCH_CHECK();
  return TFMPC_OK;
}
