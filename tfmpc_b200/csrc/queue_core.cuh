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
// The solo engine.  When few problems are left a warp pops only a handful, and a lane-per-problem iteration then costs the
// full serial latency of one thread (~75 us for C3: 36 us backward, 2 x 16 us rollouts) with 31 lanes idle.  A warp that
// popped <= solo_max problems instead runs them one after another with ALL lanes working on one problem, everything in
// shared memory: the linearisation of the whole horizon is computed first, one timestep per lane; the Riccati sweep then
// assembles each timestep's Q block with one matrix entry per lane (same expressions, same summation order as the
// one-thread code, so the results are bit-identical), hands it over through shared memory, and every lane runs the
// controller and the value update redundantly (no divergence, no second hand-over); the 11 step sizes roll out on 11
// lanes at once and keep their candidates in shared memory, so the accepted one is copied, not replayed.  A warp that
// popped a single problem keeps it until it has converged (no queue round trip, no global trajectory traffic at all) -- unless
// tickets are queued with no warp to serve them, in which case it yields after every iteration.  ~30 us per iteration.
// The Q assembly is written as explicit fma chains in both shapes (r_fma, common.cuh), so a problem's result does not depend on
// which shape computed which iteration: results are independent of the schedule, bit for bit, in the fp32 build too.
//
// Which regime a launch runs in (full warps for the stragglers while other batches want the SMs, or spread out with the solo
// engine when the batch has the GPU to itself) is the host's choice per launch: QParams w_target / w_solo / solo_max / bulk,
// set by ilqr_queue.cu from the "queue_mode" option.
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
enum { C_HEAD = 0, C_TAIL = 32, C_DONE = 64, C_ALIVE = 96, C_ERR = 128, C_WITER = 160, C_LANES = 192, C_ROUNDS = 224, C_REPLAYS = 256, C_COUNT = 288, C_TRACE = 320, C_BULK = 352, C_INTS = 384 };

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
  int solo_max;                 // a warp that popped <= solo_max problems runs them one after another on the solo engine (0 = never)
  int w_solo;                   // once <= w_solo problems are unfinished every warp pops ONE (and keeps it on the solo engine)
  // Draining pipeline (throughput mode).  `bulk` counts, per device, the queue solves that still have more than `bulk_thr`
  // unfinished problems, i.e. that can use every warp slot they get: a solve adds itself when its kernel starts and leaves
  // when it drops below the threshold.  While other solves are in their bulk phase a batch's stragglers stay lane-per-problem
  // (the measured optimum for pipelined throughput); once NONE is -- the last batches of a run, or a gap in the submissions --
  // a warp that popped a lone problem takes it on the solo engine: nothing else wants the issue slots.  nullptr = off.
  int *bulk;
  int bulk_thr;
  unsigned long long watchdog_ns;
  const real *x0, *u_init;
  real *states, *actions, *costs;
  int32_t *stats;
  // optional scheduling trace (diagnostics, option "queue_trace"): one record of 8 words per warp iteration
  //   {acquire start [ns, low 32 bits of %globaltimer], wait for tickets [ns], work [ns], lanes | rounds << 8 | warp slot << 16,
  //    state set-up [ns], backward [ns], search rounds [ns], store pass [ns]}   (the rest of `work` is results / re-queue)
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

// Results of one finished problem in the reference layouts (states [B,T+1,n], actions [B,T,m], costs [B,T+1]), written by
// the whole warp: lane t handles timesteps t, t + 32, ... of the nominal `nom` (shared or global memory).
template <int KIND, int N, int M>
WD void write_results(WarpRT &rt, const EnvSmall &e, const QParams &q, int b, const VecTraj<N, M> &nom) {
  const int T = q.T;
  real *Sx = q.states + (int64_t)b * (T + 1) * N, *A = q.actions + (int64_t)b * T * M, *Cc = q.costs + (int64_t)b * (T + 1);
  for (int t = rt.lane; t <= T; t += 32) {
    real x[N], u[M];
    nom.load_xu(t, x, u);
#pragma unroll
    for (int i = 0; i < N; i++) Sx[t * N + i] = x[i];
    if (t < T) {
#pragma unroll
      for (int i = 0; i < M; i++) A[t * M + i] = u[i];
      Cc[t] = env_cost<KIND, N, M>(e, x, u);
    } else {
      Cc[T] = env_final_cost<KIND, N, M>(e, x);
    }
  }
}

// ================================================================== the solo engine (see the header comment)
template <int N, int M>
struct Solo {
  static constexpr int Z = N + M, CHn = VecTraj<N, M>::CH;
#ifdef TFMPC_QUEUE_NO_SOLO
  static constexpr bool supported = false;                // A/B builds
#else
  static constexpr bool supported = Z * Z <= 32;          // one Q entry per lane
#endif
  static constexpr int LIN = N * Z + Z + Z * Z + 1;        // reals per linearised timestep: F = [f_x f_u] [N][Z], l_z [Z], l_zz [Z][Z], l
  static constexpr int G = M * N + M;                      // reals per timestep of gains: K [M][N], k [M]
  static constexpr int QBS = (QB<N, M>::SIZE + 3) / 4 * 4;
  struct Map { int nom, xq, gain, lin, cand, bytes; };     // byte offsets inside the warp's shared memory
  HD static Map map(int T, int row_r4) {
    Map m;
    int o = 0;
    m.nom = o; o += row_r4 * (int)sizeof(R4);
    m.xq = o; o += 2 * QBS * (int)sizeof(real);
    m.gain = o; o += (T * G * (int)sizeof(real) + 15) / 16 * 16;
    m.lin = m.cand = o;                                     // the linearisation is dead once the backward sweep is over
    const int lin_b = T * LIN * (int)sizeof(real), cand_b = N_ALPHA * row_r4 * (int)sizeof(R4);
    o += lin_b > cand_b ? lin_b : cand_b;
    m.bytes = o;
    return m;
  }
};

// derivatives (ilqr.py:84-92) of the whole nominal trajectory, one timestep per lane, into shared memory
template <int KIND, int N, int M>
WD void solo_linearize(WarpRT &rt, char *smb, const typename Solo<N, M>::Map &mp, const EnvSmall &e, int T) {
  typedef Solo<N, M> S;
  constexpr int Z = S::Z;
  const VecTraj<N, M> nom = {(R4 *)(smb + mp.nom), S::CHn, 1};
  real *lin = (real *)(smb + mp.lin);
  for (int t = rt.lane; t < T; t += 32) {
    real x[N], u[M];
    nom.load_xu(t, x, u);
    Lin<N, M> L;
    env_linearize<KIND, N, M>(e, x, u, L);
    real *r = lin + t * S::LIN;
#pragma unroll
    for (int p = 0; p < N; p++) {
#pragma unroll
      for (int c = 0; c < N; c++) r[p * Z + c] = L.f_x[p * N + c];
#pragma unroll
      for (int c = 0; c < M; c++) r[p * Z + N + c] = L.f_u[p * M + c];
    }
#pragma unroll
    for (int i = 0; i < N; i++) r[N * Z + i] = L.l_x[i];
#pragma unroll
    for (int i = 0; i < M; i++) r[N * Z + N + i] = L.l_u[i];
    real *zz = r + N * Z + Z;
#pragma unroll
    for (int a = 0; a < Z; a++)
#pragma unroll
      for (int b = 0; b < Z; b++)
        zz[a * Z + b] = (a < N && b < N) ? L.l_xx[(a < N ? a : 0) * N + (b < N ? b : 0)]
                        : (a < N) ? L.l_xu[(a < N ? a : 0) * M + (b >= N ? b - N : 0)]
                        : (b < N) ? L.l_xu[(b < N ? b : 0) * M + (a >= N ? a - N : 0)]     // l_ux = l_xu^T (ilqr.py:131)
                                  : L.l_uu[(a >= N ? a - N : 0) * M + (b >= N ? b - N : 0)];
    r[S::LIN - 1] = L.l;
  }
  rt.syncwarp();
}

// iLQR.backward (ilqr.py:94-172) of ONE problem by the whole warp.  Stage 1 of every timestep (assemble_q) is spread over
// the lanes -- lane a * Z + b computes entry (a, b) of [Q_xx Q_xu; Q_ux Q_uu] and its regularised twin, the lanes with
// b == 0 also Q_z[a] -- with exactly the expressions of assemble_q (the structural zeros it skips add +0 here); stages
// 2 and 3 run redundantly on every lane.  Gains go to shared memory.  Returns backward_pass()'s status.
template <int KIND, int N, int M, int QP>
WD int solo_backward(WarpRT &rt, char *smb, const typename Solo<N, M>::Map &mp, const EnvSmall &e, int T, real mu, real &J, real &dV1,
                     real &dV2, real &gsum) {
  typedef Solo<N, M> S;
  typedef QB<N, M> Q;
  constexpr int Z = S::Z;
  const VecTraj<N, M> nom = {(R4 *)(smb + mp.nom), S::CHn, 1};
  const real *lin = (const real *)(smb + mp.lin);
  real *xq = (real *)(smb + mp.xq), *gain = (real *)(smb + mp.gain);
  const int lane = rt.lane;
  const bool worker = lane < Z * Z;
  const int a = worker ? lane / Z : 0, b = worker ? lane % Z : 0;
  int off_q = -1, off_qr = -1, off_z = -1;
  if (worker) {
    if (a < N && b < N) off_q = Q::OXX + a * N + b;
    else if (a >= N && b < N) { off_q = Q::OUX + (a - N) * N + b; off_qr = Q::OUXR + (a - N) * N + b; }
    else if (a >= N && b >= N) { off_q = Q::OUU + (a - N) * M + (b - N); off_qr = Q::OUUR + (a - N) * M + (b - N); }
    if (b == 0) off_z = a < N ? Q::OX + a : Q::OU + (a - N);
  }
  real V_x[N], V_xx[N * N], x[N], u[M];
  nom.load_x(T, x);
  env_final_quad<KIND, N, M>(e, x, J, V_x, V_xx);  // :101-104
  dV1 = 0; dV2 = 0; gsum = 0;
  int status = 0;
  // this lane's slice of the linearised timestep, loaded one step ahead (the loads do not depend on the value function)
  real Fa[N], Fb[N], lzz, lz, lt, Fa_n[N], Fb_n[N], lzz_n, lz_n, lt_n;
  {
    const real *r = lin + (T - 1) * S::LIN;
#pragma unroll
    for (int p = 0; p < N; p++) { Fa_n[p] = r[p * Z + a]; Fb_n[p] = r[p * Z + b]; }
    lzz_n = r[N * Z + Z + a * Z + b]; lz_n = r[N * Z + a]; lt_n = r[S::LIN - 1];
  }
  for (int t = T - 1; t >= 0; t--) {
#pragma unroll
    for (int p = 0; p < N; p++) { Fa[p] = Fa_n[p]; Fb[p] = Fb_n[p]; }
    lzz = lzz_n; lz = lz_n; lt = lt_n;
    real qz, qe, qr;
    {
      real s = 0;   // explicit fma chains, as in assemble_q: the two code shapes round identically (common.cuh, r_fma)
#pragma unroll
      for (int p = 0; p < N; p++) s = r_fma(Fa[p], V_x[p], s);
      qz = lz + s;
      real sq = 0, sqr = 0;
#pragma unroll
      for (int p = 0; p < N; p++) {
        real tp = 0, tr = 0;
#pragma unroll
        for (int c = 0; c < N; c++) {
          tp = r_fma(Fa[c], V_xx[c * N + p], tp);
          tr = r_fma(Fa[c], (c == p ? V_xx[c * N + p] + mu * (real)1 : V_xx[c * N + p]), tr);
        }
        sq = r_fma(tp, Fb[p], sq);
        sqr = r_fma(tr, Fb[p], sqr);
      }
      qe = lzz + sq; qr = lzz + sqr;
    }
    real *X = xq + (t & 1) * S::QBS;   // two hand-over blocks: the reads of step t + 2 are ordered before these writes by step t + 1's barrier
    if (off_q >= 0) X[off_q] = qe;
    if (off_qr >= 0) X[off_qr] = qr;
    if (off_z >= 0) X[off_z] = qz;
    rt.syncwarp();
    Q q;
#pragma unroll
    for (int i = 0; i < Q::SIZE; i++) q.v[i] = X[i];
    nom.load_xu(t, x, u);
    {
      const real *r = lin + (t > 0 ? t - 1 : 0) * S::LIN;
#pragma unroll
      for (int p = 0; p < N; p++) { Fa_n[p] = r[p * Z + a]; Fb_n[p] = r[p * Z + b]; }
      lzz_n = r[N * Z + Z + a * Z + b]; lz_n = r[N * Z + a]; lt_n = r[S::LIN - 1];
    }
    real K[M * N], k[M];
    const int st = controller<KIND, N, M, QP>(e, q, V_xx, u, K, k);
    if (st == 1) { status = 1; break; }   // unconstrained Cholesky failed: the caller retries (ilqr.py:305-309)
    if (st) status = st;
    value_update<N, M>(q, K, k, lt, V_x, V_xx, J, dV1, dV2);   // (spread over the lanes it is no faster: the shuffles that
                                                                            //  re-broadcast V sit on the critical path -- measured 23.8 vs 22.5 us)
    real mx = 0;
#pragma unroll
    for (int i = 0; i < M; i++) {
      real v = r_abs(k[i]) / (r_abs(u[i]) + (real)1.0);
      mx = (i == 0 || v > mx) ? v : mx;
    }
    gsum += mx;
    if (lane == 0) {
      real *g = gain + t * S::G;
#pragma unroll
      for (int i = 0; i < M * N; i++) g[i] = K[i];
#pragma unroll
      for (int i = 0; i < M; i++) g[M * N + i] = k[i];
    }
  }
  rt.syncwarp();
  return status;
}

// the line search (_forward, ilqr.py:317-355): all N_ALPHA step sizes at once, one per lane, candidates kept in shared memory
template <int KIND, int N, int M>
WD void solo_search(WarpRT &rt, char *smb, const typename Solo<N, M>::Map &mp, const EnvSmall &e, const IlqrOpts &o, int T, int row_r4,
                    const Prob &p, bool &accept, int &chosen, int &rollouts, real &residual) {
  typedef Solo<N, M> S;
  const VecTraj<N, M> nom = {(R4 *)(smb + mp.nom), S::CHn, 1};
  const real *gain = (const real *)(smb + mp.gain);
  const CostSink none = {nullptr, 0};
  const bool run = rt.lane < N_ALPHA;
  const real alpha = o.alphas[run ? rt.lane : 0];
  real J = 0, res = 0;
  if (run) {
    const VecTraj<N, M> cand = {(R4 *)(smb + mp.cand) + rt.lane * row_r4, S::CHn, 1};
    real x[N];
#pragma unroll
    for (int i = 0; i < N; i++) x[i] = 0;
#pragma unroll 2   // lets the loads of step t + 1 (independent of x) be scheduled above the arithmetic of step t
    for (int t = 0; t < T; t++) {
      NomRec<N, M> use;
      const real *g = gain + t * S::G;
#pragma unroll
      for (int i = 0; i < M * N; i++) use.K[i] = g[i];
#pragma unroll
      for (int i = 0; i < M; i++) use.k[i] = g[M * N + i];
      nom.load_xu(t, use.xh, use.uh);
      if (t == 0) {
#pragma unroll
        for (int i = 0; i < N; i++) x[i] = use.xh[i];
      }
      forward_step<KIND, N, M>(e, alpha, use, t, x, cand, none, J, res);
    }
    cand.store_x(T, x);
    J += env_final_cost<KIND, N, M>(e, x);
  }
  const unsigned acc = rt.ballot(run && ls_accepts(o, alpha, p.J_hat, p.dV1, p.dV2, J));   // first accept wins (:322-353)
  accept = acc != 0;
  chosen = accept ? nth_set_bit(acc, 0) : N_ALPHA - 1;   // all rejected: _forward returns the last candidate
  rollouts = accept ? chosen + 1 : N_ALPHA;
  residual = rt.shfl(res, chosen);
}

// Runs problem b on the solo engine: one iteration (sticky = false; the problem's state goes back to global memory for
// the re-queue) or all of them (sticky = true).  Returns true when the problem has finished (results written).
template <int KIND, int N, int M, int QP, class WarpSmem>
WD bool solo_run(WarpRT &rt, WarpSmem &sm, const EnvSmall &e, const IlqrOpts &o, const QParams &q, int b, bool fresh, bool sticky,
                 int &n_iter, int &n_taken, unsigned &ns_lin, unsigned &ns_bwd, unsigned &ns_search) {
  typedef Solo<N, M> S;
  const int lane = rt.lane, T = q.T, row_r4 = q.row_r4;
  char *smb = (char *)&sm;
  const typename S::Map mp = S::map(T, row_r4);
  R4 *nom_s = (R4 *)(smb + mp.nom);
  const VecTraj<N, M> nom = {nom_s, S::CHn, 1};
  const CostSink none = {nullptr, 0};
  Prob p;
  prob_init(p);
  rt.syncwarp();   // whatever the warp did in this shared memory before is over
  if (fresh) {     // iLQR.start (ilqr.py:53-82) with the supplied initial actions
    if (lane == 0) {
      real xs[N];
#pragma unroll
      for (int i = 0; i < N; i++) xs[i] = q.x0[(int64_t)b * N + i];
      start_pass<KIND, N, M>(e, T, xs, q.u_init + (int64_t)b * T * M, nom, none);
    }
  } else {
    const QProb s = q.prob[b];
    p.mu = s.mu; p.delta = s.delta; p.iteration = s.iteration; p.n_bwd = s.n_bwd; p.n_fwd = s.n_fwd;
    p.cur = s.cg & 1; p.guard = s.cg >> 1;
    const R4 *row = traj_row(q, p.cur, b);
    for (int i = lane; i < row_r4; i += 32) nom_s[i] = row[i];
  }
  rt.syncwarp();
  bool write_back = fresh;   // the global copy of the nominal is stale (or was never written)
  for (;;) {
    // ---- backward (+ the retry wrapper _backward, ilqr.py:285-315)
    unsigned long long tq0 = 0, tq1 = 0, tq2 = 0;
    if (q.trace) tq0 = rt.now_ns();
    solo_linearize<KIND, N, M>(rt, smb, mp, e, T);
    if (q.trace) tq1 = rt.now_ns();
    {
      double mu_l = p.mu, delta_l = p.delta;   // the retry bump is local (:308-309,315)
      int tries = 0, bst;
      real gsum;
      for (;;) {
        bst = solo_backward<KIND, N, M, QP>(rt, smb, mp, e, T, (real)mu_l, p.J_hat, p.dV1, p.dV2, gsum);
        p.n_bwd++;
        if (bst != 1 || ++tries > 200) break;
        delta_l = fmax(o.delta_0, delta_l * o.delta_0);
        mu_l = fmax(o.mu_min, mu_l * delta_l);
      }
      p.phase = PH_SEARCH;   // g_norm test, ilqr.py:243-248 (same rules as tick_backward)
      const real g = gsum / (real)T;
      if (bst) { p.status = TFMPC_ST_NONPD; p.phase = PH_DONE; }
      else if (!(g == g)) { p.status = TFMPC_ST_NAN; p.phase = PH_DONE; }
      else if (g < o.atol) { p.status = TFMPC_ST_CONVERGED; p.phase = PH_DONE; }
    }
    n_iter++;
    if (q.trace) { tq2 = rt.now_ns(); ns_lin += (unsigned)(tq1 - tq0); ns_bwd += (unsigned)(tq2 - tq1); }
    // ---- line search and schedule (ilqr.py:253-270)
    if (p.phase == PH_SEARCH) {
      bool accept;
      int chosen, rollouts;
      real residual;
      solo_search<KIND, N, M>(rt, smb, mp, e, o, T, row_r4, p, accept, chosen, rollouts, residual);
      if (tick_finish(o, accept, residual, rollouts, p)) {   // the candidate becomes the nominal
        const R4 *cand = (const R4 *)(smb + mp.cand) + chosen * row_r4;
        rt.syncwarp();
        for (int i = lane; i < row_r4; i += 32) nom_s[i] = cand[i];
        p.cur ^= 1;
        write_back = true;
        n_taken++;
      }
      rt.syncwarp();
      if (q.trace) ns_search += (unsigned)(rt.now_ns() - tq2);
    }
    if (p.phase == PH_DONE) break;
    // A lone problem stays with its warp (sticky) only while nobody is waiting: with fewer live warps than unfinished problems
    // (throughput mode retires warps early) a warp that kept its problem to the end would leave the queued ones without any
    // progress for up to 64 iterations, then run them one after the other -- measured -4 % pipelined.  With tickets queued it
    // yields after every iteration and the problems take turns.
    bool yield = !sticky;
    if (sticky && q.bulk) yield = rt.shfl(lane == 0 ? rt.ld_relaxed(q.ctrl + C_COUNT) : 0, 0) > 0;   // (q.bulk set = throughput mode)
    if (yield) {     // state back to global memory, the caller re-queues the problem
      if (write_back) {
        R4 *row = traj_row(q, p.cur, b);
        for (int i = lane; i < row_r4; i += 32) row[i] = nom_s[i];
      }
      if (lane == 0) {
        QProb s;
        s.mu = p.mu; s.delta = p.delta; s.iteration = p.iteration; s.n_bwd = p.n_bwd; s.n_fwd = p.n_fwd; s.cg = (p.cur & 1) | (p.guard << 1);
        q.prob[b] = s;
      }
      return false;
    }
  }
  // ---- finished
  write_results<KIND, N, M>(rt, e, q, b, nom);
  if (lane == 0) {
    int32_t *st = q.stats + (int64_t)b * 4;
    st[0] = p.iteration; st[1] = p.n_bwd; st[2] = p.n_fwd; st[3] = p.status;
  }
  return true;
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
    if (q.bulk && P <= q.bulk_thr && rt.ld_relaxed(ctrl + C_BULK) == 1 && rt.atomic_cas(ctrl + C_BULK, 1, 2) == 1)
      rt.atomic_add(q.bulk, -1);                         // this solve has left its bulk phase (once per solve)
    if (P <= 0) return 0;
    int g = (P + q.w_target - 1) / q.w_target;
    g = g < 1 ? 1 : (g > 32 ? 32 : g);
    if (P <= q.w_solo) g = 1;
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
  const bool solo_ok = Solo<N, M>::supported && Solo<N, M>::map(T, q.row_r4).bytes <= (int)sizeof(sm);
  if (q.bulk && lane == 0 && B > q.bulk_thr && rt.ld_relaxed(q.ctrl + C_BULK) == 0 && rt.atomic_cas(q.ctrl + C_BULK, 0, 1) == 0)
    rt.atomic_add(q.bulk, 1);   // this solve enters its bulk phase (the first warp of the kernel to run)
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
    unsigned long long tp0 = 0, tp1 = 0, tp2 = 0, tp3 = 0;   // phase boundaries (trace only)
    bool keep = false;
    int nrounds = 0, n_taken = 0;
    bool use_solo = false;
    if constexpr (Solo<N, M>::supported) {
      use_solo = solo_ok && take <= q.solo_max;
      if (solo_ok && !use_solo && take == 1 && q.bulk) use_solo = rt.shfl(lane == 0 ? rt.ld_relaxed(q.bulk) : 0, 0) <= 0;   // draining pipeline
    }
    if (use_solo) {
      // ------------------------------------------------ few problems: one after another, the whole warp on each (solo engine)
      int iters = 0;
      unsigned ns_lin = 0, ns_bwd = 0, ns_search = 0;
      if constexpr (Solo<N, M>::supported)
      for (int j = 0; j < take; j++) {
        const int bj = rt.shfl(b, j);
        const bool fj = rt.shfl((int)fresh, j) != 0, vj = rt.shfl((int)valid, j) != 0;
        bool done_j = true;
        if (vj) done_j = solo_run<KIND, N, M, QP>(rt, sm, e, o, q, bj, fj, take == 1, iters, n_taken, ns_lin, ns_bwd, ns_search);
        if (lane == j) keep = valid && !done_j;
      }
      n_witer += iters; n_lanes += iters; n_rounds += iters; n_stores += n_taken;
      // trace columns of a solo visit (rounds = 0): {linearise ns, backward ns, search ns, iterations of this visit}
      tp0 = tr1 + ns_lin; tp1 = tp0 + ns_bwd; tp2 = tp1 + ns_search; tp3 = tp2 + (unsigned)iters;
    } else {
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
    if (q.trace && lane == 0) tp0 = rt.now_ns();

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
    if (q.trace && lane == 0) tp1 = rt.now_ns();

    // ------------------------------------------------ line search in rounds (_forward, ilqr.py:317-355)
    bool searching = valid && p.phase == PH_SEARCH, accept = false;
    int next_ai = 0, chosen = 0, rollouts = 0;
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
    if (q.trace && lane == 0) tp2 = rt.now_ns();
    if (take_mask) {
      real J, res;
      rollout_round<KIND, N, M, true>(rt, sm, e, T, NH, take_mask, lane, take_cand, o.alphas[take_cand ? chosen : 0], gain_ws, J, res);
      nrounds++;
    }
    if (q.trace && lane == 0) tp3 = rt.now_ns();
    if (valid && p.phase == PH_SEARCH && tick_finish(o, accept, residual, rollouts, p)) p.cur ^= 1;   // ilqr.py:253-270

    // ------------------------------------------------ results
    rt.fence();      // trajectory lines were written by other lanes of the warp (and must be visible grid-wide before the ticket is)
    rt.syncwarp();
    keep = valid && p.phase != PH_DONE;
    // finished problems: one after the other, each written by the whole warp (a lane writing its own 51 records one by one
    // costs the warp ~1,500 instructions whenever any of its problems finishes; this is ~60 per finished problem)
    for (unsigned fm = rt.ballot(valid && !keep); fm; fm &= fm - 1) {
      const int src = nth_set_bit(fm, 0);
      const int bs = rt.shfl(b, src), cs = rt.shfl(p.cur, src);
      write_results<KIND, N, M>(rt, e, q, bs, VecTraj<N, M>{traj_row(q, cs, bs), CHn, 1});
    }
    if (valid && !keep) {
      int32_t *st = q.stats + (int64_t)b * 4;
      st[0] = p.iteration; st[1] = p.n_bwd; st[2] = p.n_fwd; st[3] = p.status;
    }
    if (keep) {
      QProb s;
      s.mu = p.mu; s.delta = p.delta; s.iteration = p.iteration; s.n_bwd = p.n_bwd; s.n_fwd = p.n_fwd; s.cg = (p.cur & 1) | (p.guard << 1);
      q.prob[b] = s;
    }
    n_witer++; n_lanes += take; n_rounds += nrounds; n_stores += popc32(take_mask);
    }
    // ------------------------------------------------ re-queue the problems that go on, count the finished ones
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
    if (q.trace && lane == 0) {
      const int slot = rt.atomic_add(q.ctrl + C_TRACE, 1);
      if (slot < q.trace_cap) {
        unsigned *r = q.trace + (int64_t)slot * 8;
        r[0] = (unsigned)tr0; r[1] = (unsigned)(tr1 - tr0); r[2] = (unsigned)(rt.now_ns() - tr1);
        r[3] = (unsigned)take | ((unsigned)nrounds << 8) | ((unsigned)warp_slot << 16);
        r[4] = (unsigned)(tp0 - tr1); r[5] = (unsigned)(tp1 - tp0); r[6] = (unsigned)(tp2 - tp1); r[7] = (unsigned)(tp3 - tp2);
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
