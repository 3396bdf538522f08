// Thread-per-problem iLQR kernels for small environments (NavigationLQR n<=4, Navigation n=2).
//
// Why thread-per-problem: the headline workload (BASELINE config C3) has n = m = 2, so one
// problem's Riccati state (V_xx: 3 unique values, V_x: 2, K: 4, k: 2) fits in a handful of
// registers and a warp-per-problem mapping would idle 30 of 32 lanes (SURVEY.md section 7,
// "hard parts").  Data layout: the solve workspace is made of 16-byte vector chunks (R4, small_core.cuh) --
// trajectories [set][t][slot][lane], gains [t][slot][chunk] -- see struct WS below and DESIGN.md section 4.
#include <algorithm>
#include <atomic>
#include <mutex>
#include <vector>
#include <cstring>

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
  StridedTraj<N, M> out = {states + b * (T + 1) * N, actions + b * T * M, 1};
  CostSink Co = {costs + b * (T + 1), 1};
  start_pass<KIND, N, M>(e, T, x, u_init + b * T * M, out, Co);
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_backward(EnvSmall e, int64_t B, int T, const real *__restrict__ states,
                                                       const real *__restrict__ actions, real mu, real *__restrict__ K,
                                                       real *__restrict__ k, real *__restrict__ J, real *__restrict__ dV1,
                                                       real *__restrict__ dV2, int32_t *__restrict__ status) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  StridedTraj<N, M> nom = {const_cast<real *>(states) + b * (T + 1) * N, const_cast<real *>(actions) + b * T * M, 1};
  StridedGain<N, M> gain = {K + b * T * M * N, k + b * T * M, 1};
  real Jb, d1, d2, g;
  int st = backward_pass<KIND, N, M>(e, T, nom, mu, gain, Jb, d1, d2, g);
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
  StridedTraj<N, M> nom = {const_cast<real *>(states) + b * (T + 1) * N, const_cast<real *>(actions) + b * T * M, 1};
  StridedGain<N, M> gain = {const_cast<real *>(K) + b * T * M * N, const_cast<real *>(k) + b * T * M, 1};
  StridedTraj<N, M> out = {xs + b * (T + 1) * N, us + b * T * M, 1};
  CostSink Co = {cs + b * (T + 1), 1};
  real Jb, res;
  forward_pass<KIND, N, M>(e, T, nom, gain, alpha, out, Co, Jb, res);
  J[b] = Jb; residual[b] = res;
}

// ------------------------------------------------------------------ the batched solve, as per-tick kernels
// iLQR.solve for B problems = a fixed launch sequence with no host round trip:
//     k_init, then per tick { k_tick_backward, k_tick_linesearch }, then k_costs + 3 x k_transpose_out.
// A tick is one turn of the reference's inner loop (ilqr.py:238-270) for every still-active problem.
//  * k_tick_backward   one thread per ACTIVE problem (compacted index list): linearise + Riccati sweep + g_norm test.
//  * k_tick_linesearch GA (= 4) lanes per active problem: the lanes roll out GA consecutive line-search
//                      step sizes concurrently, each into its own candidate buffer; the FIRST accepted alpha
//                      (lowest index, ilqr.py:322-353) is found with a warp ballot; further passes of GA
//                      alphas run only for the groups that rejected all of them.  The group leader then
//                      applies the mu/delta schedule and convergence tests (tick_finish) and re-appends the
//                      problem to the next tick's active list (warp-aggregated atomic).
// Accepting a candidate is a buffer-index switch (p.cur), not a copy: each problem owns two sets of GA
// trajectory buffers; the nominal is the accepted member of the set that is not being written.  Converged
// problems drop out of the list, so late ticks touch few problems and cost only their latency.
constexpr int GA = 4;            // line-search lanes (and candidate buffers) per problem
constexpr int NBUF = 2 * GA;      // two candidate sets of GA trajectories; the nominal is one member of the set not being written
constexpr int kExtraTicks = 24;  // ticks beyond max_iterations available to regularisation retries (ilqr.py:267-270)

struct WS {  // carve-up of the workspace
  int *count;                    // [2] active counts (ping-pong)
  int *list[2];                  // [S] active problem slots
  int *iteration, *n_bwd, *n_fwd, *status, *cur, *phase, *guard;
  double *mu, *delta;
  real *J_hat, *dV1, *dV2;
  R4 *traj;                      // [set 0..1][(T+1) records x CHn chunks][S slots][GA lanes]: the GA candidates of one
                                 // (problem, timestep) are 64 contiguous bytes, so a line-search group writes full sectors
  R4 *gain;                      // [T records][S slots][CHg chunks]: one problem-step of gains = one 32-byte sector (n=m=2)
  real *cost;                    // (T+1) rows x S
  int64_t S, traj_chunks;        // traj_chunks = (T+1) * CHn
};

inline int64_t ws_bytes_for(int64_t S, int T, int N, int M) {
  int64_t chn = (N + M + 3) / 4, chg = (M * N + M + 3) / 4;
  int64_t vec = NBUF * (int64_t)(T + 1) * chn + (int64_t)T * chg;
  return 256 + S * (2 * 4 + 7 * 4 + 2 * 8 + 3 * (int64_t)sizeof(real)) + 64 + vec * S * (int64_t)sizeof(R4) + (int64_t)(T + 1) * S * (int64_t)sizeof(real);
}

inline WS carve(void *ws, int64_t S, int T, int N, int M) {
  WS w;
  char *p = (char *)ws;
  w.S = S;
  w.count = (int *)p; p += 256;
  w.mu = (double *)p; p += S * 8;
  w.delta = (double *)p; p += S * 8;
  w.list[0] = (int *)p; p += S * 4;
  w.list[1] = (int *)p; p += S * 4;
  int **ints[7] = {&w.iteration, &w.n_bwd, &w.n_fwd, &w.status, &w.cur, &w.phase, &w.guard};
  for (int i = 0; i < 7; i++) { *ints[i] = (int *)p; p += S * 4; }
  real **reals[3] = {&w.J_hat, &w.dV1, &w.dV2};
  for (int i = 0; i < 3; i++) { *reals[i] = (real *)p; p += S * sizeof(real); }
  p = (char *)(((uintptr_t)p + 63) & ~(uintptr_t)63);
  int64_t chn = (N + M + 3) / 4, chg = (M * N + M + 3) / 4;
  w.traj_chunks = (int64_t)(T + 1) * chn;
  w.traj = (R4 *)p; p += NBUF * w.traj_chunks * S * sizeof(R4);
  w.gain = (R4 *)p; p += (int64_t)T * chg * S * sizeof(R4);
  w.cost = (real *)p;
  return w;
}

// trajectory buffer `buf` = set * GA + lane of problem slot b
template <int N, int M>
__device__ __forceinline__ VecTraj<N, M> buf_traj(const WS &w, int buf, int64_t b) {
  const int set = buf / GA, ln = buf % GA;
  return VecTraj<N, M>{w.traj + ((int64_t)set * w.traj_chunks * w.S + b) * GA + ln, VecTraj<N, M>::CH * w.S * GA, w.S * GA};
}
template <int N, int M>
__device__ __forceinline__ VecGain<N, M> buf_gain(const WS &w, int64_t b) {
  return VecGain<N, M>{w.gain + b * VecGain<N, M>::CH, w.S * VecGain<N, M>::CH, 1};
}

__device__ __forceinline__ void load_prob(const WS &w, int64_t b, Prob &p) {
  p.mu = w.mu[b]; p.delta = w.delta[b]; p.iteration = w.iteration[b]; p.n_bwd = w.n_bwd[b]; p.n_fwd = w.n_fwd[b];
  p.status = w.status[b]; p.cur = w.cur[b]; p.phase = w.phase[b]; p.guard = w.guard[b];
  p.J_hat = w.J_hat[b]; p.dV1 = w.dV1[b]; p.dV2 = w.dV2[b];
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_init(EnvSmall e, int64_t B, int T, const real *__restrict__ x0,
                                                   const real *__restrict__ u_init, WS w) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b == 0) { w.count[0] = (int)B; w.count[1] = 0; }
  if (b >= B) return;
  Prob p;
  prob_init(p);
  w.mu[b] = p.mu; w.delta[b] = p.delta; w.iteration[b] = 0; w.n_bwd[b] = 0; w.n_fwd[b] = 0; w.status[b] = p.status;
  w.cur[b] = 0; w.phase[b] = PH_SEARCH; w.guard[b] = 0;
  w.list[0][b] = (int)b;
  real x[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = x0[b * N + i];
  const CostSink none = {nullptr, 0};
  start_pass<KIND, N, M>(e, T, x, u_init + b * T * M, buf_traj<N, M>(w, 0, b), none);
}

// COOP: warp-cooperative box-QP backtracking (bounded environments).  The loop is warp-uniform: lanes past the end of
// the active list shadow the last active problem (same nominal, gains written to the spare slot S-1, no state written),
// so that all 32 lanes stay in lock step through the cooperative phases.
template <int KIND, int N, int M, int QP>
__global__ void __launch_bounds__(kThreads) k_tick_backward(EnvSmall e, IlqrOpts o, int T, WS w, int parity) {
  const int cnt = w.count[parity];
  if (blockIdx.x == 0 && threadIdx.x == 0) w.count[parity ^ 1] = 0;  // filled by this tick's line-search kernel
  const int *__restrict__ list = w.list[parity];
  const int lane = threadIdx.x & 31;
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int base = tid - lane; base < cnt; base += nth) {
    const int i = base + lane;
    const bool valid = i < cnt;
    const int64_t b = list[valid ? i : cnt - 1];
    Prob p;
    load_prob(w, b, p);
    tick_backward<KIND, N, M, QP>(e, o, T, buf_traj<N, M>(w, p.cur, b), buf_gain<N, M>(w, valid ? b : w.S - 1), p);
    if (valid) {
      w.n_bwd[b] = p.n_bwd; w.status[b] = p.status; w.phase[b] = p.phase;
      w.J_hat[b] = p.J_hat; w.dV1[b] = p.dV1; w.dV2[b] = p.dV2;
    }
  }
}

// ---- line-search rollout with the nominal / gain records staged through shared memory
// The rollout is a dependent chain (x_{t+1} needs x_t), one step is ~100 instructions, and every step needs 48 bytes
// (n = m = 2) of nominal + gains that live in HBM: ncu showed the register-ring prefetch of forward_pass() reaching
// only ~1.5 steps ahead and >50% of the kernel's stall samples waiting on those loads.  Here the GA lanes of a problem
// share one copy of each record, fetched Ring::depth (8) steps ahead by cp.async (LDGSTS) into a per-warp shared-memory ring, so
// the loads hold no registers and the prefetch distance covers a DRAM round trip.
template <int N, int M>
struct Ring {   // depth: 8 steps, fewer when the records are large (static shared memory budget of 32 KB per CTA)
  static constexpr int CHT = VecTraj<N, M>::CH + VecGain<N, M>::CH;
  static constexpr int slot_bytes = (kThreads / 32) * (32 / 4) * CHT * (int)sizeof(R4);
  static constexpr int depth = 32768 / slot_bytes >= 8 ? 8 : (32768 / slot_bytes >= 4 ? 4 : 2);
};

__device__ __forceinline__ void cp_async_r4(R4 *dst_smem, const R4 *src) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
#pragma unroll
  for (int o = 0; o < (int)sizeof(R4); o += 16)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d + o), "l"((const char *)src + o) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int PENDING>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory"); }

// One rollout per lane (forward_pass semantics, ilqr.py:174-212); the GA lanes of a group share ring[..][sub][..].
// Called by all 32 lanes of the warp; `fetch`: this lane's group still searches (its records are staged), `run`: this
// lane has a step size to try.
template <int KIND, int N, int M>
__device__ __forceinline__ void rollout_staged(const EnvSmall &e, int T, const VecTraj<N, M> &nom, const VecGain<N, M> &gain, real alpha,
                                               const VecTraj<N, M> &out, R4 (*ring)[32 / GA][VecTraj<N, M>::CH + VecGain<N, M>::CH],
                                               int sub, int la, bool fetch, bool run, real &J, real &residual) {
  constexpr int CHn = VecTraj<N, M>::CH, CHg = VecGain<N, M>::CH, CHT = CHn + CHg, kRing = Ring<N, M>::depth;
  constexpr unsigned FULL = 0xffffffffu;
  const CostSink none = {nullptr, 0};
  auto issue = [&](int t) {
    if (fetch && t < T) {
#pragma unroll
      for (int c = la; c < CHT; c += GA) {
        const R4 *src = c < CHn ? nom.base + (int64_t)t * nom.ts + c * nom.cs : gain.base + (int64_t)t * gain.ts + (c - CHn) * gain.cs;
        cp_async_r4(&ring[t % kRing][sub][c], src);
      }
    }
    cp_async_commit();   // one group per step on every lane, empty or not, so that wait_group counts steps
  };
#pragma unroll
  for (int d = 0; d < kRing; d++) issue(d);
  real x[N];
  J = 0; residual = 0;
  for (int t = 0; t < T; t++) {
    cp_async_wait<kRing - 1>();
    __syncwarp(FULL);                      // every lane's share of step t has landed
    NomRec<N, M> r;
    {
      real v[4 * CHT];
      const R4 *rec = ring[t % kRing][sub];
#pragma unroll
      for (int c = 0; c < CHT; c++) {
        const R4 q = rec[c];
#pragma unroll
        for (int j = 0; j < 4; j++) v[4 * c + j] = q.v[j];
      }
#pragma unroll
      for (int i = 0; i < N; i++) r.xh[i] = v[i];
#pragma unroll
      for (int i = 0; i < M; i++) r.uh[i] = v[N + i];
#pragma unroll
      for (int i = 0; i < M * N; i++) r.K[i] = v[4 * CHn + i];
#pragma unroll
      for (int i = 0; i < M; i++) r.k[i] = v[4 * CHn + M * N + i];
    }
    __syncwarp(FULL);                      // slot t % kRing is free again
    issue(t + kRing);
    if (t == 0) {
#pragma unroll
      for (int i = 0; i < N; i++) x[i] = r.xh[i];
    }
    if (run) forward_step<KIND, N, M>(e, alpha, r, t, x, out, none, J, residual);
  }
  cp_async_wait<0>();
  if (run) {
    out.store_x(T, x);
    J += env_final_cost<KIND, N, M>(e, x);
  }
}

template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_tick_linesearch(EnvSmall e, IlqrOpts o, int T, WS w, int parity) {
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int PASSES = (N_ALPHA + GA - 1) / GA;
  const int cnt = w.count[parity];
  const int *__restrict__ list = w.list[parity];
  int *__restrict__ next = w.list[parity ^ 1];
  const int lane = threadIdx.x & 31, la = lane % GA, sub = lane / GA;
  const int groups_per_warp = 32 / GA;
  const int warp_global = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nwarps = (gridDim.x * blockDim.x) >> 5;
  __shared__ R4 ring[kThreads / 32][Ring<N, M>::depth][32 / GA][Ring<N, M>::CHT];
  for (int base = warp_global * groups_per_warp; base < cnt; base += nwarps * groups_per_warp) {  // warp-uniform trip count
    const int gi = base + sub;
    const bool valid = gi < cnt;
    const int64_t b = valid ? list[gi] : list[0];
    Prob p;
    load_prob(w, b, p);
    const bool was_search = valid && p.phase == PH_SEARCH;
    bool searching = was_search, accept = false;
    real residual = 0;
    int rollouts = 0, cand = 0;
    const VecTraj<N, M> nom = buf_traj<N, M>(w, p.cur, b);
    const VecGain<N, M> gain = buf_gain<N, M>(w, b);
    const int cset = 1 - p.cur / GA;               // candidates go to the set that does not hold the nominal
    const int mybuf = cset * GA + la;
    const VecTraj<N, M> mine = buf_traj<N, M>(w, mybuf, b);
    for (int pass = 0; pass < PASSES; pass++) {
      if (!__any_sync(FULL, searching)) break;
      const int ai = pass * GA + la;
      const bool run = searching && ai < N_ALPHA;
      real J = 0, res = 0;
      bool acc = false;
      const real alpha = o.alphas[run ? ai : 0];
      rollout_staged<KIND, N, M>(e, T, nom, gain, alpha, mine, ring[threadIdx.x >> 5], sub, la, searching, run, J, res);
      if (run) acc = ls_accepts(o, alpha, p.J_hat, p.dV1, p.dV2, J);
      const unsigned gm = (__ballot_sync(FULL, acc) >> (sub * GA)) & ((1u << GA) - 1u);
      const int last = min(GA - 1, N_ALPHA - 1 - pass * GA);      // last alpha lane of this pass
      const int src = gm ? (__ffs(gm) - 1) : last;                 // first accepted lane, else the last candidate
      const real r_src = __shfl_sync(FULL, res, src, GA);
      if (searching) {
        residual = r_src;
        cand = cset * GA + src;
        if (gm) { accept = true; rollouts = pass * GA + src + 1; searching = false; }
        else { rollouts = pass * GA + last + 1; if (pass == PASSES - 1) searching = false; }
      }
    }
    bool keep = false;
    if (valid && la == 0) {
      if (was_search) {
        if (tick_finish(o, accept, residual, rollouts, p)) p.cur = cand;
        w.mu[b] = p.mu; w.delta[b] = p.delta; w.iteration[b] = p.iteration; w.n_fwd[b] = p.n_fwd; w.status[b] = p.status;
        w.cur[b] = p.cur; w.phase[b] = p.phase; w.guard[b] = p.guard;
      }
      keep = p.phase != PH_DONE;
    }
    // warp-aggregated append to the next tick's active list (keeps slot order within a warp)
    const unsigned km = __ballot_sync(FULL, keep);
    if (km) {
      int pos = 0;
      if (lane == __ffs(km) - 1) pos = atomicAdd(&w.count[parity ^ 1], __popc(km));
      pos = __shfl_sync(FULL, pos, __ffs(km) - 1);
      if (keep) next[pos + __popc(km & ((1u << lane) - 1u))] = (int)b;
    }
  }
}

// cost(x_t, u_t) of the final nominal (what the accepting forward pass / start computed, ilqr.py:199,208) + stats
template <int KIND, int N, int M>
__global__ void __launch_bounds__(kThreads) k_costs(EnvSmall e, int64_t B, int T, WS w, int32_t *__restrict__ stats) {
  int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  const VecTraj<N, M> nom = buf_traj<N, M>(w, w.cur[b], b);
  real x[N], u[M];
  for (int t = 0; t < T; t++) {
    nom.load_xu(t, x, u);
    w.cost[(int64_t)t * w.S + b] = env_cost<KIND, N, M>(e, x, u);
  }
  nom.load_x(T, x);
  w.cost[(int64_t)T * w.S + b] = env_final_cost<KIND, N, M>(e, x);
  int st = w.status[b];
  if (w.phase[b] != PH_DONE) st = TFMPC_ST_TICKS;  // ran out of ticks (more than kExtraTicks rejected line searches in total)
  reinterpret_cast<int4 *>(stats)[b] = make_int4(w.iteration[b], w.n_bwd[b], w.n_fwd[b], st);
}

// Workspace -> reference layouts.  out[b][r] for r < nrows: 32 x 32 tiles through shared memory so that both the
// slot-fastest reads and the problem-major writes are coalesced.
// which: 0 = states (r = t*N + i), 1 = actions (r = t*M + i) -- from the nominal buffer of each problem; 2 = costs.
template <int N, int M>
__global__ void __launch_bounds__(128) k_transpose_out(int64_t B, int T, WS w, int which, int nrows, real *__restrict__ out) {
  __shared__ real tile[4][32][33];
  constexpr int CH = (N + M + 3) / 4;
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int64_t slot0 = ((int64_t)blockIdx.x * 4 + wq) * 32;
  if (slot0 >= B) return;
  const int64_t b = slot0 + lane;
  const bool vb = b < B;
  const real *rec = nullptr;
  if (which < 2 && vb) rec = reinterpret_cast<const real *>(buf_traj<N, M>(w, w.cur[b], b).base);
  for (int r0 = 0; r0 < nrows; r0 += 32) {
#pragma unroll 4
    for (int rr = 0; rr < 32; rr++) {
      const int r = r0 + rr;
      if (vb && r < nrows) {
        real v;
        if (which == 2) v = w.cost[(int64_t)r * w.S + b];
        else {
          const int t = which == 0 ? r / N : r / M, j = which == 0 ? r % N : N + r % M;
          v = rec[((int64_t)(t * CH + (j >> 2)) * w.S * GA) * 4 + (j & 3)];
        }
        tile[wq][rr][lane] = v;
      }
    }
    __syncwarp();
#pragma unroll 4
    for (int ss = 0; ss < 32; ss++)
      if (slot0 + ss < B && r0 + lane < nrows) out[(slot0 + ss) * nrows + r0 + lane] = tile[wq][lane][ss];
    __syncwarp();
  }
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

static int64_t padded_slots(int64_t B) { return (B + 1 + 31) / 32 * 32; }  // >= B + 1: slot S-1 is a spare (see k_tick_backward)

int64_t small_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T) { return ws_bytes_for(padded_slots(B), T, e->n, e->m); }

static int device_sms(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

// Straggler streams.  After kHeadTicks ticks most problems have converged and what remains is latency-bound: a few
// problems, each tick a strictly sequential H-step sweep.  The tail of the launch sequence is therefore forked onto an
// internal HIGH-PRIORITY stream (event fork / join, so the caller's stream still observes one ordered operation): when the
// caller keeps several batches in flight on different streams, the tiny high-priority tail kernels of one batch slip in
// between the large head kernels of the next instead of queueing behind them (measured: without priorities concurrent
// full-size solves do not overlap at all, because the head grids monopolise CTA dispatch).
constexpr int kHeadTicks = 24;
constexpr int kAsyncHeadTicks = 4;
constexpr int kTailStreams = 8;

static cudaStream_t tail_stream(int device) {
  static cudaStream_t pool[16][kTailStreams] = {};
  static std::atomic<unsigned> rr[16];
  static std::mutex mu;
  if (device < 0 || device >= 16) return nullptr;
  unsigned i = rr[device].fetch_add(1) % kTailStreams;
  std::lock_guard<std::mutex> g(mu);
  if (!pool[device][i]) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);  // hi = numerically lowest = greatest priority
    if (cudaStreamCreateWithPriority(&pool[device][i], cudaStreamNonBlocking, hi) != cudaSuccess) { cudaGetLastError(); pool[device][i] = nullptr; }
  }
  return pool[device][i];
}

// ---- CUDA-graph cache of the launch sequence
// One solve is 253 launches; enqueueing them costs the host ~1.4 ms (5.6 us each with a 1.3 KB parameter block), and on
// a shared host that is what limits a pipeline of batches (measured: four ranks on one 24-vCPU box spend 6 ms per batch
// enqueueing).  The sequence is static -- same kernels, grids and arguments for a given (environment, B, T, options,
// buffers) -- so the second time a key is seen the two halves (ticks before the fork, ticks after it + result kernels)
// are captured once and replayed with two cudaGraphLaunch calls (tfmpc_set_graph_mode(1) or TFMPC_GRAPH=1).
struct GraphKey {
  unsigned long long uid;
  int64_t B;
  int T, fork_at;
  const void *ptr[7];
  IlqrOpts o;
};
static bool same_key(const GraphKey &a, const GraphKey &b) {
  if (a.uid != b.uid || a.B != b.B || a.T != b.T || a.fork_at != b.fork_at) return false;
  for (int i = 0; i < 7; i++) if (a.ptr[i] != b.ptr[i]) return false;
  if (a.o.atol != b.o.atol || a.o.c1 != b.o.c1 || a.o.max_iterations != b.o.max_iterations || a.o.mu_min != b.o.mu_min || a.o.delta_0 != b.o.delta_0) return false;
  for (int i = 0; i < N_ALPHA; i++) if (a.o.alphas[i] != b.o.alphas[i]) return false;
  return true;
}
struct GraphEntry {
  GraphKey key;
  cudaGraphExec_t head = nullptr, tail = nullptr;
  bool captured = false;
  unsigned long long stamp = 0;
};
// off by default: replaying the sequence from graphs shortens a lone solve by 3 % (15.8 -> 15.3 ms) and cuts the host
// cost per solve from 1.4 to 0.3 ms, but with several batches in flight the graph launches made the overlap of one
// batch's stragglers with the next batch's head erratic (pipelined throughput 115-194 M/s against 201-208 M/s direct).
static std::atomic<int> g_graph_mode{[] { const char *v = getenv("TFMPC_GRAPH"); return (v && atoi(v) != 0) ? 1 : 0; }()};
static std::mutex g_graph_mu;
static std::vector<GraphEntry> g_graphs;
static unsigned long long g_graph_clock = 0;
constexpr size_t kGraphCacheEntries = 64;

static void destroy_entry(GraphEntry &g) {
  if (g.head) cudaGraphExecDestroy(g.head);
  if (g.tail) cudaGraphExecDestroy(g.tail);
  g.head = g.tail = nullptr;
}

static cudaStream_t capture_stream(int device, bool high) {   // called with g_graph_mu held
  static cudaStream_t pool[16][2] = {};
  if (device < 0 || device >= 16) return nullptr;
  if (!pool[device][high]) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    if (cudaStreamCreateWithPriority(&pool[device][high], cudaStreamNonBlocking, high ? hi : lo) != cudaSuccess) { cudaGetLastError(); pool[device][high] = nullptr; }
  }
  return pool[device][high];
}

template <int KIND, int N, int M>
static int solve_launch(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                        real *actions, real *costs, int32_t *stats, void *ws, cudaStream_t s, cudaEvent_t done) {
  const WS w = carve(ws, padded_slots(B), T, N, M);
  const int sms = device_sms(e->device);
  const unsigned gB = grid_for(B);
  // persistent-style grids: enough threads for every active problem (x GA lanes), capped at a few waves
  const unsigned g_bwd = (unsigned)std::min<int64_t>(gB, (int64_t)sms * 16);
  const unsigned g_ls = (unsigned)std::min<int64_t>((B * GA + kThreads - 1) / kThreads, (int64_t)sms * 16);
  // tail: one CTA per SM is plenty for the stragglers (grid-stride loops keep it correct for any count)
  const unsigned g_bwd_t = std::min<unsigned>(g_bwd, (unsigned)sms), g_ls_t = std::min<unsigned>(g_ls, (unsigned)sms * 2);
  const int ticks = o.max_iterations + kExtraTicks;
  // Ticks before `fork_at` run on the caller's stream, the rest on a priority stream.  Synchronous form: fork where the
  // kernels stop filling the GPU (kHeadTicks).  Asynchronous form: fork early (kAsyncHeadTicks) -- the caller's stream then
  // carries only the GPU-filling first ticks of each batch, back to back, and everything latency-bound overlaps them.
  static const int sync_head = [] { const char *v = getenv("TFMPC_HEAD_TICKS"); return v ? atoi(v) : kHeadTicks; }();          // tuning knobs
  static const int async_head = [] { const char *v = getenv("TFMPC_ASYNC_HEAD_TICKS"); return v ? atoi(v) : kAsyncHeadTicks; }();
  const bool no_graph = g_graph_mode.load() == 0;
  const int fork_at = std::min(ticks, std::max(0, done ? async_head : sync_head));

  // the two halves of the sequence, enqueued on whatever stream they are given (directly, or under capture)
  auto enqueue_ticks = [&](int t0, int t1, cudaStream_t q) {
    for (int t = t0; t < t1; t++) {
      const bool big = t < kHeadTicks;   // grid size follows the expected active count, whichever stream the tick runs on
      if (e->bounded) k_tick_backward<KIND, N, M, QP_COOP><<<big ? g_bwd : g_bwd_t, kThreads, 0, q>>>(e->es, o, T, w, t & 1);
      else k_tick_backward<KIND, N, M, QP_NEWTON><<<big ? g_bwd : g_bwd_t, kThreads, 0, q>>>(e->es, o, T, w, t & 1);
      k_tick_linesearch<KIND, N, M><<<big ? g_ls : g_ls_t, kThreads, 0, q>>>(e->es, o, T, w, t & 1);
    }
  };
  auto enqueue_head = [&](cudaStream_t q) {
    k_init<KIND, N, M><<<gB, kThreads, 0, q>>>(e->es, B, T, x0, u_init, w);
    enqueue_ticks(0, fork_at, q);
  };
  auto enqueue_tail = [&](cudaStream_t q) {
    enqueue_ticks(fork_at, ticks, q);
    k_costs<KIND, N, M><<<gB, kThreads, 0, q>>>(e->es, B, T, w, stats);
    const unsigned gT = (unsigned)((B + 127) / 128);
    k_transpose_out<N, M><<<gT, 128, 0, q>>>(B, T, w, 0, (T + 1) * N, states);
    k_transpose_out<N, M><<<gT, 128, 0, q>>>(B, T, w, 1, T * M, actions);
    k_transpose_out<N, M><<<gT, 128, 0, q>>>(B, T, w, 2, T + 1, costs);
  };

  // graph lookup: first sight of a key -> remember it and launch directly; second sight -> capture; then replay
  cudaGraphExec_t gh = nullptr, gt = nullptr;
  std::unique_lock<std::mutex> lock(g_graph_mu, std::defer_lock);   // held until the graphs are launched: an eviction by another thread would destroy them
  if (!no_graph) {
    GraphKey key;
    memset(&key, 0, sizeof(key));
    key.uid = e->uid; key.B = B; key.T = T; key.fork_at = fork_at; key.o = o;
    const void *ptrs[7] = {x0, u_init, states, actions, costs, stats, ws};
    for (int i = 0; i < 7; i++) key.ptr[i] = ptrs[i];
    lock.lock();
    GraphEntry *hit = nullptr;
    for (auto &g : g_graphs) if (same_key(g.key, key)) { hit = &g; break; }
    if (!hit) {
      if (g_graphs.size() >= kGraphCacheEntries) {   // evict the least recently used entry
        size_t victim = 0;
        for (size_t i = 1; i < g_graphs.size(); i++) if (g_graphs[i].stamp < g_graphs[victim].stamp) victim = i;
        destroy_entry(g_graphs[victim]);
        g_graphs.erase(g_graphs.begin() + victim);
      }
      GraphEntry g;
      g.key = key; g.stamp = ++g_graph_clock;
      g_graphs.push_back(g);
    } else {
      hit->stamp = ++g_graph_clock;
      if (!hit->captured) {
        hit->captured = true;   // one attempt; on any failure the entry stays without graphs and the launches stay direct
        cudaStream_t ch = capture_stream(e->device, false), ct = capture_stream(e->device, true);
        cudaGraph_t graph = nullptr;
        bool ok = ch && ct;
        if (ok && cudaStreamBeginCapture(ch, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
          enqueue_head(ch);
          ok = cudaStreamEndCapture(ch, &graph) == cudaSuccess && graph && cudaGraphInstantiate(&hit->head, graph, 0) == cudaSuccess;
          if (graph) cudaGraphDestroy(graph);
        } else ok = false;
        graph = nullptr;
        if (ok && cudaStreamBeginCapture(ct, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {   // captured on a priority stream: the nodes keep it
          enqueue_tail(ct);
          ok = cudaStreamEndCapture(ct, &graph) == cudaSuccess && graph && cudaGraphInstantiate(&hit->tail, graph, 0) == cudaSuccess;
          if (graph) cudaGraphDestroy(graph);
        } else ok = false;
        if (!ok) { cudaGetLastError(); destroy_entry(*hit); }
      }
      gh = hit->head; gt = hit->tail;
    }
    if (!(gh && gt)) lock.unlock();
  }

  cudaStream_t ts = (ticks > fork_at) ? tail_stream(e->device) : nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
  if (ts) {
    if (cudaEventCreateWithFlags(&fork, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      if (fork) cudaEventDestroy(fork);
      fork = join = nullptr;
      ts = nullptr;
    }
  }
  if (gh && gt) cudaGraphLaunch(gh, s); else enqueue_head(s);
  cudaStream_t q = s;
  if (ts) { cudaEventRecord(fork, s); cudaStreamWaitEvent(ts, fork, 0); q = ts; }
  if (gh && gt) { cudaGraphLaunch(gt, q); lock.unlock(); } else enqueue_tail(q);
  if (ts) {
    // async form: the caller's stream does NOT wait for the stragglers -- it is free for the head of the next batch --
    // and `done` completes when the results are in place
    if (done) cudaEventRecord(done, ts);
    else { cudaEventRecord(join, ts); cudaStreamWaitEvent(s, join, 0); }
    cudaEventDestroy(fork);   // released by the runtime once the recorded work has completed
    cudaEventDestroy(join);
  } else if (done) cudaEventRecord(done, s);
  // kernels executed by this solve, launched directly or replayed from the two captured graphs: k_init, two per tick, k_costs
  // and three k_transpose_out (the same sequence in both cases; a graph replay runs its kernel nodes, it does not skip any)
  tfmpc_count_launch(1 + 2 * ticks + 1 + 3);
  return TFMPC_OK;
}

int small_ilqr_graph_mode(int on) { return g_graph_mode.exchange(on ? 1 : 0); }

void small_ilqr_forget(const tfmpc_env *e) {
  std::lock_guard<std::mutex> lock(g_graph_mu);
  for (size_t i = 0; i < g_graphs.size();) {
    if (g_graphs[i].key.uid == e->uid) { destroy_entry(g_graphs[i]); g_graphs.erase(g_graphs.begin() + i); }
    else i++;
  }
}

int small_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, cudaStream_t s, cudaEvent_t done) {
  if (ws_bytes < small_ilqr_workspace_bytes(e, B, T)) return tfmpc_set_error(TFMPC_E_WORKSPACE, "workspace too small");
  if (B > 0x7fffffff) return tfmpc_set_error(TFMPC_E_INVALID, "batch too large");
#define CALL(KD, N, M) { int rc_ = solve_launch<KD, N, M>(e, B, T, x0, u_init, o, states, actions, costs, stats, ws, s, done); if (rc_) return rc_; }
  SMALL_DISPATCH(e, CALL);
#undef CALL
  {   // (the launches were counted by solve_launch)
    cudaError_t err_ = cudaGetLastError();
    if (err_ != cudaSuccess) return tfmpc_set_error(TFMPC_E_CUDA, "CUDA launch failed: %s", cudaGetErrorString(err_));
  }
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
