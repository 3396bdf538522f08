// Persistent work-queue iLQR solver for SMALL environments (thread-per-problem arithmetic of small_core.cuh), written
// against the warp runtime shim (warp_rt.cuh) so that the host-emulation tests can execute it on the CPU.
//
// iLQR.solve (reference tfmpc/solvers/ilqr.py:214-283) for B problems in ONE launch.  Iteration counts vary 7-100 across
// problems, so a fixed problem -> thread mapping wastes most of a warp on its slowest lane.  Instead every warp loops:
//
//   acquire   pop up to 32 problem tickets from a global queue (fresh problems first, then re-queued survivors)
//   backward  one lane per problem: linearise + Riccati sweep + box-QP (ilqr.py:84-172, :285-315), gains -> per-warp scratch
//   search    the line search (ilqr.py:317-355) in ROUNDS of 32 concurrent rollouts that only accumulate J: round 1 tries
//             alpha_0 of every problem (one lane each); the problems that rejected it share the 32 lanes for their next
//             step sizes -- as many per problem as fit (the accepted index is 0/1/2/3 for 23/38/20/18 % of the C3
//             iterations, so a full warp needs ~3 rounds; a warp with <= 2 problems evaluates all 11 at once).  Then ONE
//             store pass replays the step size each problem settled on and leaves the candidate in the problem's other
//             trajectory buffer: the line search writes exactly one trajectory per iteration (the tick kernels wrote 4).
//   finish    mu / delta schedule and convergence tests (ilqr.py:245-270) per problem; finished problems write their
//             results in the reference layouts, the others are pushed back onto the queue
//
// No grid-wide barrier and no relaunch: a problem's next iteration starts as soon as any warp is free, warps that find
// the queue empty retire (freeing their SM slots for the next batch's kernel), and near the end the pop size shrinks so
// that the last stragglers each own a warp (all 11 step sizes in one round).
//
// Memory traffic.  Trajectories are PROBLEM-major (one problem = one contiguous, 64-byte aligned row) and move between
// HBM/L2 and shared memory only as 64-byte half-lines (4 steps for n = m = 2): 4 lanes fetch one half-line with cp.async
// (8 rows per instruction instead of 32 scattered sectors), the store pass overwrites the staged nominal records in place
// with the candidate, and the same 4-lane pattern writes the half-line back.  Gains never leave the warp:
// [warp][t][pair][lane] scratch written by the backward sweep and streamed back by the rollouts, again with cp.async, one
// contiguous block per half-line of steps.  Everything a rollout consumes is requested 4-8 steps ahead of its use.
#pragma once
#include "small_core.cuh"
#include "warp_rt.cuh"

namespace tq {

// control block (ints, one counter per 128-byte line)
enum { C_HEAD = 0, C_TAIL = 32, C_DONE = 64, C_ALIVE = 96, C_ERR = 128, C_WITER = 160, C_LANES = 192, C_ROUNDS = 224, C_REPLAYS = 256, C_COUNT = 288, C_TRACE = 320, C_INTS = 352 };

struct alignas(16) QProb {   // per-problem solver state carried between iterations (one 32-byte sector)
  double mu, delta;
  int iteration, n_bwd, n_fwd, cg;   // cg = cur | guard << 1
};

struct alignas(2 * sizeof(real)) R2 { real v[2]; };

struct QParams {
  int *ctrl;                    // [C_INTS]
  unsigned long long *ring;     // [ring_mask + 1] re-queue ring: ticket << 32 | problem
  unsigned ring_mask;
  QProb *prob;                  // [B]
  R4 *traj;                     // [2][B][row_r4]
  R2 *gain;                     // [nwarps][T][Gain2::CH2][32]
  int B, T, row_r4;
  int w_target;                 // warps the pop size is planned for: pop = clamp(ceil(outstanding / w_target), 1, 32)
  int patience;                 // idle polls before a warp settles for fewer problems than the planned pop size
  unsigned long long watchdog_ns;
  const real *x0, *u_init;
  real *states, *actions, *costs;
  int32_t *stats;
  // optional scheduling trace (diagnostics, option "queue_trace"): one record per warp iteration
  //   {acquire start [ns, low 32 bits of %globaltimer], wait for tickets [ns], work [ns], lanes | rounds << 8 | warp slot << 16}
  unsigned *trace;
  int trace_cap;
};

constexpr int CPH = 4;         // 16-byte chunks (R4) per staged half-line of a trajectory row
template <int GAIN_HALF_R2>    // R2 elements of one staged half-line of gains: steps per half-line * Gain2::CH2 * 32
struct WarpSmemT {
  R4 buf[2][32][CPH + 1];       // two half-line buffers: [row = lane of the owning problem][4 chunks + 1 pad (bank spread)]
  const R4 *inrow[32];          // nominal trajectory row of each lane's problem
  R4 *outrow[32];               // candidate trajectory row (the problem's other buffer)
  R2 gbuf[2][GAIN_HALF_R2];     // rollouts: gains of two half-lines of steps, [step][pair][lane] as in the scratch
};

HD int popc32(unsigned m) {
#ifdef __CUDA_ARCH__
  return __popc(m);
#else
  return __builtin_popcount(m);
#endif
}
HD int nth_set_bit(unsigned m, int n) {   // index of the n-th (0-based) set bit; m must have more than n bits set
  for (int i = 0; i < n; i++) m &= m - 1;
#ifdef __CUDA_ARCH__
  return __ffs((int)m) - 1;
#else
  return __builtin_ffs((int)m) - 1;
#endif
}

HD R4 *traj_row(const QParams &q, int buf, int b) { return q.traj + ((int64_t)buf * q.B + b) * q.row_r4; }

// Gains (K_t, k_t) of the problem held by one lane: per-warp scratch [t][pair][lane] in 2-real units, so a warp's store
// or load of one pair is a single contiguous 256-byte (fp32) row and a problem-step costs M*N+M reals with no padding
// (24 bytes for n = m = 2).  `base` already points at the lane's column.
// L2 residency hint (TFMPC_QUEUE_L2HINT, fp32 build): gain accesses carry an evict_last policy so that the scratch a warp
// re-reads within the same iteration outlives the streaming trajectory lines in the 126 MB L2.
#if defined(__CUDA_ARCH__) && defined(TFMPC_QUEUE_L2HINT) && !defined(TFMPC_F64)
#define TQ_GAIN_HINT 1
__device__ __forceinline__ unsigned long long tq_policy_evict_last() {
  unsigned long long p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
#endif

template <int N, int M>
struct Gain2 {
  static constexpr int G = M * N + M, CH2 = (G + 1) / 2;
  R2 *base;
  HD R2 ld(const R2 *p) const {
#ifdef TQ_GAIN_HINT
    R2 v;
    asm volatile("ld.global.L2::cache_hint.v2.f32 {%0, %1}, [%2], %3;\n" : "=f"(v.v[0]), "=f"(v.v[1]) : "l"(p), "l"(tq_policy_evict_last()));
    return v;
#else
    return *p;
#endif
  }
  HD void st(R2 *p, const R2 &v) const {
#ifdef TQ_GAIN_HINT
    asm volatile("st.global.L2::cache_hint.v2.f32 [%0], {%1, %2}, %3;\n" ::"l"(p), "f"(v.v[0]), "f"(v.v[1]), "l"(tq_policy_evict_last()) : "memory");
#else
    *p = v;
#endif
  }
  HD void load(int t, real *Kt, real *kt) const {
    real r[2 * CH2];
#pragma unroll
    for (int c = 0; c < CH2; c++) {
      const R2 v = ld(base + ((int64_t)t * CH2 + c) * 32);
      r[2 * c] = v.v[0]; r[2 * c + 1] = v.v[1];
    }
#pragma unroll
    for (int i = 0; i < M * N; i++) Kt[i] = r[i];
#pragma unroll
    for (int i = 0; i < M; i++) kt[i] = r[M * N + i];
  }
  HD void store(int t, const real *Kt, const real *kt) const {
    real r[2 * CH2];
#pragma unroll
    for (int i = 0; i < 2 * CH2; i++) r[i] = 0;
#pragma unroll
    for (int i = 0; i < M * N; i++) r[i] = Kt[i];
#pragma unroll
    for (int i = 0; i < M; i++) r[M * N + i] = kt[i];
#pragma unroll
    for (int c = 0; c < CH2; c++) {
      R2 v;
      v.v[0] = r[2 * c]; v.v[1] = r[2 * c + 1];
      st(base + ((int64_t)t * CH2 + c) * 32, v);
    }
  }
};

// ---- cooperative half-line staging: lanes 4i..4i+3 move the 4 chunks of one row's half-line, 8 rows per instruction.
// The fetches only ISSUE the copies; the caller closes a group with cp_commit().
template <class WarpSmem>
WD void fetch_half(WarpRT &rt, WarpSmem &sm, int p, int h, int NH, unsigned rows) {
  if (h >= 0 && h < NH) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const int row = 8 * i + (rt.lane >> 2), c = rt.lane & 3;
      if ((rows >> row) & 1u) {
        const char *src = (const char *)(sm.inrow[row] + h * CPH + c);
        char *dst = (char *)&sm.buf[p][row][c];
#pragma unroll
        for (int o = 0; o < (int)sizeof(R4); o += 16) rt.cp_async16(dst + o, src + o);   // one chunk: 16 bytes (fp32) / 32 bytes (fp64 build)
      }
    }
  }
}
template <class WarpSmem>
WD void writeout_half(WarpRT &rt, WarpSmem &sm, int p, int h, unsigned rows) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int row = 8 * i + (rt.lane >> 2), c = rt.lane & 3;
    if ((rows >> row) & 1u) sm.outrow[row][h * CPH + c] = sm.buf[p][row][c];
  }
}

template <int N, int M>
struct SlotOut {   // forward_step's trajectory sink: overwrite the staged record in place (storing lanes only)
  R4 *rec;
  bool on;
  HD void store_xu(int, const real *x, const real *u) const {
    if (on) VecTraj<N, M>{rec, 0, 1}.store_xu(0, x, u);
  }
  HD void store_x(int, const real *x) const {
    if (on) VecTraj<N, M>{rec, 0, 1}.store_x(0, x);
  }
};

// ---- iLQR.backward (ilqr.py:94-172) + derivatives (:84-92) for the lanes with `act`, nominal streamed last half-line
// first.  Every lane of the warp calls this (the staging is cooperative).  Returns backward_pass()'s status for `act` lanes.
template <int KIND, int N, int M, int QP, class WarpSmem>
WD int backward_staged(WarpRT &rt, WarpSmem &sm, const EnvSmall &e, int T, int NH, bool act, real mu, R2 *gain_lane, real &J, real &dV1,
                       real &dV2, real &gsum) {
  constexpr int CHn = VecTraj<N, M>::CH, HS = CPH / CHn;   // records per half-line
  const unsigned rows = rt.ballot(act);
  const Gain2<N, M> gain = {gain_lane};
  fetch_half(rt, sm, (NH - 1) & 1, NH - 1, NH, rows);
  rt.cp_commit();
  fetch_half(rt, sm, (NH - 2) & 1, NH - 2, NH, rows);
  rt.cp_commit();
  real V_x[N], V_xx[N * N];
  int status = 0;
  bool live = act;
  J = 0; dV1 = 0; dV2 = 0; gsum = 0;
  for (int h = NH - 1; h >= 0; h--) {
    rt.template cp_wait<1>();
    rt.syncwarp();
    if (live) {
#pragma unroll 1
      for (int s = HS - 1; s >= 0; s--) {
        const int t = h * HS + s;
        if (t > T || !live) continue;
        real x[N], u[M];
        VecTraj<N, M>{&sm.buf[h & 1][rt.lane][s * CHn], 0, 1}.load_xu(0, x, u);
        if (t == T) {
          env_final_quad<KIND, N, M>(e, x, J, V_x, V_xx);  // :101-104
          continue;
        }
        Lin<N, M> L;
        env_linearize<KIND, N, M>(e, x, u, L);
        real K[M * N], k[M];
        const int st = backward_step<KIND, N, M, QP>(e, L, u, mu, V_x, V_xx, J, dV1, dV2, K, k);
        if (st == 1) { status = 1; live = false; continue; }   // unconstrained Cholesky failed: the caller retries (ilqr.py:305-309)
        if (st) status = st;
        real mx = 0;
#pragma unroll
        for (int i = 0; i < M; i++) {
          real v = r_abs(k[i]) / (r_abs(u[i]) + (real)1.0);
          mx = (i == 0 || v > mx) ? v : mx;
        }
        gsum += mx;
        gain.store(t, K, k);
      }
    }
    rt.syncwarp();
    fetch_half(rt, sm, h & 1, h - 2, NH, rows);
    rt.cp_commit();
  }
  rt.template cp_wait<0>();
  return status;
}

// ---- one line-search round: up to 32 concurrent rollouts (iLQR.forward, ilqr.py:174-212).  Lane = (problem of lane
// `src`, step size alpha); `rows` = lanes whose problem takes part.  STORE = false: search round, only J and the
// residual are produced.  STORE = true: store pass (src == lane): the candidate overwrites the staged nominal records in
// place and every half-line is written to the problem's other trajectory buffer.
// Gains are staged like the nominal: the warp's scratch is [step][pair][lane], so the gains of one half-line of steps are
// one contiguous block that the 32 lanes copy together.  cp.async group H(h) = {nominal half-line h of every row, gains of
// its steps}, issued two half-lines ahead (the 126 MB L2 does not hold the ~190 MB of gains and trajectories in flight, so
// most of these reads come from HBM: with gains loaded into registers one step ahead, 45 % of the rollouts' stall samples
// were the first use of a gain -- ncu, bulk phase).
template <int N, int M>
struct GainStage {
  static constexpr int CH2 = Gain2<N, M>::CH2, HS = CPH / VecTraj<N, M>::CH, HALF_R2 = HS * CH2 * 32;
  static constexpr int HALF_BYTES = HALF_R2 * (int)sizeof(R2), STEP_BYTES = CH2 * 32 * (int)sizeof(R2);
};
template <int N, int M, class WarpSmem>
WD void fetch_gain_half(WarpRT &rt, WarpSmem &sm, const R2 *gain_warp, int h, int T) {
  typedef GainStage<N, M> GS;
  const int t0 = h * GS::HS;
  if (t0 < T) {
    const char *src = (const char *)(gain_warp + (int64_t)t0 * GS::CH2 * 32);
    char *dst = (char *)sm.gbuf[h & 1];
    const int valid_bytes = (T - t0 < GS::HS ? T - t0 : GS::HS) * GS::STEP_BYTES;
#pragma unroll
    for (int o = 0; o < GS::HALF_BYTES; o += 32 * 16) {
      const int off = o + rt.lane * 16;
      if (off < valid_bytes) rt.cp_async16(dst + off, src + off);
    }
  }
}

template <int KIND, int N, int M, bool STORE, class WarpSmem>
WD void rollout_round(WarpRT &rt, WarpSmem &sm, const EnvSmall &e, int T, int NH, unsigned rows, int src, bool run, real alpha,
                      const R2 *gain_warp, real &J, real &residual) {
  typedef GainStage<N, M> GS;
  constexpr int CHn = VecTraj<N, M>::CH, HS = GS::HS;
  const CostSink none = {nullptr, 0};
#pragma unroll
  for (int h = 0; h < 2; h++) {
    fetch_half(rt, sm, h, h, NH, rows);
    fetch_gain_half<N, M>(rt, sm, gain_warp, h, T);
    rt.cp_commit();
  }
  real x[N];
#pragma unroll
  for (int i = 0; i < N; i++) x[i] = 0;
  J = 0; residual = 0;
  for (int h = 0; h < NH; h++) {
    rt.template cp_wait<1>();   // H(h) has landed (H(h+1) may still be in flight)
    rt.syncwarp();
    if (run) {
      const Gain2<N, M> gain = {sm.gbuf[h & 1] + src};
      R4 *half = &sm.buf[h & 1][src][0];
#pragma unroll 1   // code size: the warps of an SM sit in different phases and share one instruction cache
      for (int s = 0; s < HS; s++) {
        const int t = h * HS + s;
        if (t > T) continue;
        R4 *rec = half + s * CHn;
        if (t < T) {
          NomRec<N, M> use;
          gain.load(s, use.K, use.k);
          VecTraj<N, M>{rec, 0, 1}.load_xu(0, use.xh, use.uh);
          if (t == 0) {
#pragma unroll
            for (int i = 0; i < N; i++) x[i] = use.xh[i];
          }
          forward_step<KIND, N, M>(e, alpha, use, t, x, SlotOut<N, M>{rec, STORE}, none, J, residual);
        } else {
          SlotOut<N, M>{rec, STORE}.store_x(T, x);
          J += env_final_cost<KIND, N, M>(e, x);
        }
      }
    }
    rt.syncwarp();
    if (STORE) {
      writeout_half(rt, sm, h & 1, h, rows);
      rt.syncwarp();
    }
    fetch_half(rt, sm, h & 1, h + 2, NH, rows);
    fetch_gain_half<N, M>(rt, sm, gain_warp, h + 2, T);
    rt.cp_commit();
  }
  rt.template cp_wait<0>();
}

// ---- queue: lane 0 claims `take` consecutive tickets starting at h (0 = this warp retires).
// C_COUNT is a semaphore of queued tickets: a claim is one atomic subtraction (over-draws are handed back), then one
// atomic add on C_HEAD numbers the tickets.  No compare-and-swap loop: with ~2,400 warps popping every ~100 us an
// optimistic read-then-CAS never sees an unchanged head (measured: 60 ms per batch instead of 5).  A ticket may be numbered
// before its producer has published it (the credit came from a later producer); the consumer then spins on the slot tag.
WD int q_acquire(WarpRT &rt, const QParams &q, int &h_out) {
  int *ctrl = q.ctrl;
  unsigned long long t_start = 0;
  unsigned backoff = 64;
  int idle = 0;
  for (;;) {
    if (rt.ld_relaxed(ctrl + C_ERR)) return 0;
    const int P = q.B - rt.ld_relaxed(ctrl + C_DONE);   // problems not finished yet
    if (P <= 0) return 0;
    int g = (P + q.w_target - 1) / q.w_target;
    g = g < 1 ? 1 : (g > 32 ? 32 : g);
    const int avail = rt.ld_relaxed(ctrl + C_COUNT);
    const bool settle = idle >= q.patience;   // waited long enough: take whatever is there
    if (avail >= g || (avail > 0 && settle)) {
      const int c = rt.atomic_add(ctrl + C_COUNT, -g);
      const int take = c >= g ? g : (c > 0 ? c : 0);
      if (take == g || (take > 0 && settle)) {
        if (take < g) rt.atomic_add(ctrl + C_COUNT, g - take);
        h_out = rt.atomic_add(ctrl + C_HEAD, take);
        return take;
      }
      rt.atomic_add(ctrl + C_COUNT, g);   // not enough yet: hand everything back and wait
    } else if (avail <= 0) {   // nothing queued (everything outstanding is being worked on): is this warp still needed?
      const int want = (P + g - 1) / g;
      if (rt.ld_relaxed(ctrl + C_ALIVE) > want) {
        if (rt.atomic_add(ctrl + C_ALIVE, -1) > want) return 0;   // retired
        rt.atomic_add(ctrl + C_ALIVE, 1);                         // lost the race, stay
      }
    }
    const unsigned long long now = rt.now_ns();
    if (!t_start) t_start = now;
    else if (now - t_start > q.watchdog_ns) { rt.st_relaxed(ctrl + C_ERR, 1); return 0; }
    rt.sleep_ns(backoff);
    if (backoff < 2048) backoff *= 2;
    idle++;
  }
}

// ---- the warp's main loop
template <int N, int M>
using WarpSmem = WarpSmemT<GainStage<N, M>::HALF_R2>;

template <int KIND, int N, int M, int QP>
WD void queue_warp_main(WarpRT &rt, const EnvSmall &e, const IlqrOpts &o, const QParams &q, WarpSmem<N, M> &sm, int warp_slot) {
  constexpr int CHn = VecTraj<N, M>::CH;
  static_assert(CHn == 1 || CHn == 2, "trajectory records are 1 or 2 chunks");
  const int lane = rt.lane, T = q.T, B = q.B, NH = q.row_r4 / CPH;
  const CostSink none = {nullptr, 0};
  R2 *gain_ws = q.gain + (int64_t)warp_slot * T * Gain2<N, M>::CH2 * 32;
  int n_witer = 0, n_lanes = 0, n_rounds = 0, n_stores = 0;   // scheduling statistics of this warp (flushed once, at exit)
  for (;;) {
    // ------------------------------------------------ acquire
    int h = 0, take = 0;
    unsigned long long tr0 = 0, tr1 = 0;
    if (q.trace && lane == 0) tr0 = rt.now_ns();
    if (lane == 0) take = q_acquire(rt, q, h);
    if (q.trace && lane == 0) tr1 = rt.now_ns();
    take = rt.shfl(take, 0);
    h = rt.shfl(h, 0);
    if (take <= 0) break;
    bool valid = lane < take, fresh = false;
    int b = 0;
    if (valid) {
      const unsigned ticket = (unsigned)h + (unsigned)lane;
      if (ticket < (unsigned)B) { b = (int)ticket; fresh = true; }   // first visit: tickets 0..B-1 are the problems themselves
      else {
        const unsigned long long *slot = q.ring + ((ticket - (unsigned)B) & q.ring_mask);
        unsigned long long v = rt.ld_acquire64(slot), t0 = 0;
        while ((unsigned)(v >> 32) != ticket) {   // the producer reserved the ticket but has not published it yet
          const unsigned long long now = rt.now_ns();
          if (!t0) t0 = now;
          else if (now - t0 > q.watchdog_ns) { rt.st_relaxed(q.ctrl + C_ERR, 2); valid = false; break; }
          rt.sleep_ns(32);
          v = rt.ld_acquire64(slot);
        }
        b = (int)(unsigned)(v & 0xffffffffull);
      }
    }
    // ------------------------------------------------ state
    Prob p;
    prob_init(p);
    if (valid) {
      if (fresh) {   // iLQR.start (ilqr.py:53-82) with the supplied initial actions
        real xs[N];
#pragma unroll
        for (int i = 0; i < N; i++) xs[i] = q.x0[(int64_t)b * N + i];
        start_pass<KIND, N, M>(e, T, xs, q.u_init + (int64_t)b * T * M, VecTraj<N, M>{traj_row(q, 0, b), CHn, 1}, none);
      } else {
        const QProb s = q.prob[b];
        p.mu = s.mu; p.delta = s.delta; p.iteration = s.iteration; p.n_bwd = s.n_bwd; p.n_fwd = s.n_fwd;
        p.cur = s.cg & 1; p.guard = s.cg >> 1;
      }
      sm.inrow[lane] = traj_row(q, p.cur, b);
      sm.outrow[lane] = traj_row(q, p.cur ^ 1, b);
    }
    if (rt.any(valid && fresh)) rt.fence();   // start_pass used plain stores; the staged reads are asynchronous copies issued by other lanes
    rt.syncwarp();

    // ------------------------------------------------ backward (+ the retry wrapper _backward, ilqr.py:285-315)
    {
      double mu_l = p.mu, delta_l = p.delta;   // the retry bump is local (:308-309,315)
      int tries = 0, bst = 0;
      bool need_pass = valid;
      real gsum = 0;
      while (rt.any(need_pass)) {
        real J, d1, d2, gs;
        const int st = backward_staged<KIND, N, M, QP>(rt, sm, e, T, NH, need_pass, (real)mu_l, gain_ws + lane, J, d1, d2, gs);
        if (need_pass) {
          p.n_bwd++;
          bst = st; p.J_hat = J; p.dV1 = d1; p.dV2 = d2; gsum = gs;
          if (bst != 1 || ++tries > 200) need_pass = false;
          else {
            delta_l = fmax(o.delta_0, delta_l * o.delta_0);
            mu_l = fmax(o.mu_min, mu_l * delta_l);
          }
        }
      }
      if (valid) {   // g_norm test, ilqr.py:243-248 (same rules as tick_backward)
        p.phase = PH_SEARCH;
        const real g = gsum / (real)T;
        if (bst) { p.status = TFMPC_ST_NONPD; p.phase = PH_DONE; }
        else if (!(g == g)) { p.status = TFMPC_ST_NAN; p.phase = PH_DONE; }
        else if (g < o.atol) { p.status = TFMPC_ST_CONVERGED; p.phase = PH_DONE; }
      }
    }
    rt.syncwarp();   // the gains of every lane are visible to the whole warp

    // ------------------------------------------------ line search in rounds (_forward, ilqr.py:317-355)
    bool searching = valid && p.phase == PH_SEARCH, accept = false;
    int next_ai = 0, chosen = 0, rollouts = 0, nrounds = 0;
    real residual = 0;
    for (int round = 0; round < N_ALPHA + 1; round++) {
      const unsigned act_mask = rt.ballot(searching);
      if (!act_mask) break;
      nrounds++;
      const int c = popc32(act_mask);
      int per = 32 / c;   // step sizes tried per problem this round: as many as fit in the warp
      per = per > N_ALPHA ? N_ALPHA : per;
      const int g = lane / per, a = lane - g * per;
      const bool in_group = g < c;
      const int src = in_group ? nth_set_bit(act_mask, g) : lane;
      const int my_cnt = per < N_ALPHA - next_ai ? per : N_ALPHA - next_ai;
      const int first = rt.shfl(next_ai, src), cnt = rt.shfl(my_cnt, src);
      const bool run = in_group && a < cnt;
      real J = 0, res = 0;
      rollout_round<KIND, N, M, false>(rt, sm, e, T, NH, act_mask, src, run, o.alphas[run ? first + a : 0], gain_ws, J, res);
      // each owner reads the results of its group in step-size order: first accept wins (:322-353)
      const bool owner = (act_mask >> lane) & 1u;
      const int base = owner ? popc32(act_mask & ((1u << lane) - 1u)) * per : 0;
      int hit = -1;
      real hit_res = 0, last_res = 0;
      for (int j = 0; j < per; j++) {
        const real Jj = rt.shfl(J, base + j), rj = rt.shfl(res, base + j);
        if (owner && j < my_cnt) {
          last_res = rj;
          if (hit < 0 && ls_accepts(o, o.alphas[next_ai + j], p.J_hat, p.dV1, p.dV2, Jj)) { hit = j; hit_res = rj; }
        }
      }
      if (owner) {
        if (hit >= 0) { accept = true; chosen = next_ai + hit; residual = hit_res; rollouts = chosen + 1; searching = false; }
        else {
          next_ai += my_cnt;
          residual = last_res;
          if (next_ai >= N_ALPHA) { searching = false; chosen = N_ALPHA - 1; rollouts = N_ALPHA; }   // all rejected: _forward returns the last candidate
        }
      }
    }
    // store pass: the candidate that becomes the nominal -- accepted, or rejected but converged by residual (ilqr.py:253-257)
    const bool take_cand = valid && p.phase == PH_SEARCH && (accept || residual < o.atol);
    const unsigned take_mask = rt.ballot(take_cand);
    if (take_mask) {
      real J, res;
      rollout_round<KIND, N, M, true>(rt, sm, e, T, NH, take_mask, lane, take_cand, o.alphas[take_cand ? chosen : 0], gain_ws, J, res);
      nrounds++;
    }
    if (valid && p.phase == PH_SEARCH && tick_finish(o, accept, residual, rollouts, p)) p.cur ^= 1;   // ilqr.py:253-270

    // ------------------------------------------------ results / re-queue
    rt.fence();      // trajectory lines were written by other lanes of the warp (and must be visible grid-wide before the ticket is)
    rt.syncwarp();
    const bool keep = valid && p.phase != PH_DONE;
    if (valid && !keep) {   // finished: results in the reference layouts (states [B,T+1,n], actions [B,T,m], costs [B,T+1])
      const VecTraj<N, M> nom = {traj_row(q, p.cur, b), CHn, 1};
      real *S = q.states + (int64_t)b * (T + 1) * N, *A = q.actions + (int64_t)b * T * M, *Cc = q.costs + (int64_t)b * (T + 1);
      real x[N], u[M];
      for (int t = 0; t < T; t++) {
        nom.load_xu(t, x, u);
#pragma unroll
        for (int i = 0; i < N; i++) S[t * N + i] = x[i];
#pragma unroll
        for (int i = 0; i < M; i++) A[t * M + i] = u[i];
        Cc[t] = env_cost<KIND, N, M>(e, x, u);
      }
      nom.load_x(T, x);
#pragma unroll
      for (int i = 0; i < N; i++) S[T * N + i] = x[i];
      Cc[T] = env_final_cost<KIND, N, M>(e, x);
      int32_t *st = q.stats + (int64_t)b * 4;
      st[0] = p.iteration; st[1] = p.n_bwd; st[2] = p.n_fwd; st[3] = p.status;
    }
    if (keep) {
      QProb s;
      s.mu = p.mu; s.delta = p.delta; s.iteration = p.iteration; s.n_bwd = p.n_bwd; s.n_fwd = p.n_fwd; s.cg = (p.cur & 1) | (p.guard << 1);
      q.prob[b] = s;
    }
    const unsigned km = rt.ballot(keep), dm = rt.ballot(valid && !keep);
    rt.fence();
    rt.syncwarp();
    int t0 = 0;
    if (lane == 0) {
      if (km) t0 = rt.atomic_add(q.ctrl + C_TAIL, popc32(km));
      if (dm) rt.atomic_add(q.ctrl + C_DONE, popc32(dm));
    }
    t0 = rt.shfl(t0, 0);
    if (keep) {
      const unsigned ticket = (unsigned)t0 + (unsigned)popc32(km & ((1u << lane) - 1u));
      rt.st_release64(q.ring + ((ticket - (unsigned)B) & q.ring_mask), ((unsigned long long)ticket << 32) | (unsigned)b);
    }
    rt.syncwarp();
    if (lane == 0 && km) rt.atomic_add(q.ctrl + C_COUNT, popc32(km));   // credit the semaphore once the tickets are published
    n_witer++; n_lanes += take; n_rounds += nrounds; n_stores += popc32(take_mask);
    if (q.trace && lane == 0) {
      const int slot = rt.atomic_add(q.ctrl + C_TRACE, 1);
      if (slot < q.trace_cap) {
        unsigned *r = q.trace + (int64_t)slot * 4;
        r[0] = (unsigned)tr0; r[1] = (unsigned)(tr1 - tr0); r[2] = (unsigned)(rt.now_ns() - tr1);
        r[3] = (unsigned)take | ((unsigned)nrounds << 8) | ((unsigned)warp_slot << 16);
      }
    }
  }
  if (lane == 0) {
    rt.atomic_add(q.ctrl + C_WITER, n_witer);
    rt.atomic_add(q.ctrl + C_LANES, n_lanes);
    rt.atomic_add(q.ctrl + C_ROUNDS, n_rounds);
    rt.atomic_add(q.ctrl + C_REPLAYS, n_stores);
  }
}

}  // namespace tq
