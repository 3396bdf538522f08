// Persistent work-queue iLQR solve for small environments: kernels and launcher around queue_core.cuh.
// One solve = k_queue_init (control block, ring, default stats) + k_queue_solve (one warp per CTA, persistent).
#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>

#include "queue_core.cuh"

namespace {

#if defined(TFMPC_QUEUE_MAXWPS)
constexpr int kMaxWarpsPerSM = TFMPC_QUEUE_MAXWPS;   // A/B builds (scripts/build_variant.py)
#elif defined(TFMPC_F64)
constexpr int kMaxWarpsPerSM = 8;    // WarpSmem is 23 KB in the fp64 build (n = m = 2)
#else
constexpr int kMaxWarpsPerSM = 18;   // 11.8 KB of shared memory per warp (n = m = 2) and 96 registers per thread (no spills): measured
                                     // 337 M problem-iterations/s against 306 at 16 warps / 106 registers (profiles/r02_ab_queue.txt)
#endif
constexpr int kLatencyWarpsPerSM = kMaxWarpsPerSM < 13 ? kMaxWarpsPerSM : 13;

template <int KIND, int N, int M, int QP>
__global__ void __launch_bounds__(32, kMaxWarpsPerSM) k_queue_solve(EnvSmall e, IlqrOpts o, tq::QParams q) {
  __shared__ tq::WarpSmem<N, M> sm;
  WarpRT rt;
  tq::queue_warp_main<KIND, N, M, QP>(rt, e, o, q, sm, (int)blockIdx.x);
}

__global__ void __launch_bounds__(256) k_queue_init(tq::QParams q, int nwarps) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < tq::C_INTS) q.ctrl[i] = (i == tq::C_TAIL || i == tq::C_COUNT) ? q.B : (i == tq::C_ALIVE ? nwarps : 0);
  if (i <= q.ring_mask) q.ring[i] = 0ull;   // ticket field 0 never matches a re-queue ticket (those are >= B >= 1)
  if (i < q.B) reinterpret_cast<int4 *>(q.stats)[i] = make_int4(0, 0, 0, TFMPC_ST_ABORTED);   // overwritten when the problem finishes
}

int env_int(const char *name, int dflt) {
  const char *v = getenv(name);
  return v && *v ? atoi(v) : dflt;
}

std::atomic<int> g_wps{env_int("TFMPC_QUEUE_WPS", 0)};   // 0 = the mode decides
// Scheduling policy.  The same kernel serves two regimes that want opposite things from the last few thousand problems of a
// batch: with several batches in flight (pipelined throughput) the stragglers should occupy as few warps as possible --
// full warps, one per SM, lane-per-problem -- because every other warp slot is doing another batch's bulk work; a batch
// that has the GPU to itself (latency) should spread them over all warp slots and, once there are fewer problems than
// slots, give each problem a whole warp (the solo engine: ~34 us per iteration against ~75 us on one lane).
//   mode 1 = throughput: 18 resident warps per SM, pop-size target = one warp per FOUR SMs (full warps until < 1,184 problems are
//                        left: 8 batches in flight measured 374 / 371 / 367 / 358 / 346 / 334 M/s for targets of 18 / 37 / 74 / 148 /
//                        222 / 296 warps; 20-batch runs peak at 37), solo engine off (on while the pipeline drains, see `bulk`)
//   mode 2 = latency:    13 resident warps per SM (the lone batch is bound by the latency of a warp iteration, which contention
//                        for the issue slots stretches), pop-size target = 15 warps per SM, solo once unfinished <= launched
//                        warps.  Lone C3 batch (profiles/r02_ab_queue.txt, calls 18, 36, 37): 8.14 ms at 18 warps / target 12 per
//                        SM, 7.98 at 16, 7.97 at 14, 7.67 at 13 warps / target 15 per SM, 7.75 at 12; targets of 18+ per SM: > 8.1
//   mode 0 = auto:       latency when no other stream of this device has a queue solve in flight at launch time
// Each of the knobs below overrides the mode's choice when set (> 0; solo_max: anything but 255).
std::atomic<int> g_mode{env_int("TFMPC_QUEUE_MODE", 0)};
std::atomic<int> g_w_target{env_int("TFMPC_QUEUE_WTARGET", 0)};
std::atomic<int> g_w_solo{env_int("TFMPC_QUEUE_WSOLO", 0)};
std::atomic<int> g_solo_max{env_int("TFMPC_QUEUE_SOLO", 255)};   // 255 = the mode decides
std::atomic<int> g_patience{env_int("TFMPC_QUEUE_PATIENCE", 0)};
std::atomic<int> g_trace{env_int("TFMPC_QUEUE_TRACE", 0)};
std::atomic<int> g_last_mode{0};   // what the last launch chose (diagnostics)
std::atomic<int> g_drain_solo{env_int("TFMPC_QUEUE_DRAIN_SOLO", 1)};   // throughput mode: solo engine for lone stragglers once no solve is in its bulk phase

// per-device counter of queue solves in their bulk phase (QParams::bulk); allocated once, never freed
int *device_bulk_counter(int device) {
  static std::mutex mu;
  static int *ptr[64] = {nullptr};
  if (device < 0 || device >= 64) return nullptr;
  std::lock_guard<std::mutex> lock(mu);
  if (!ptr[device]) {
    int *p = nullptr;
    if (cudaMalloc(&p, 256) != cudaSuccess || cudaMemset(p, 0, 256) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    ptr[device] = p;
  }
  return ptr[device];
}

// queue solves launched and not yet known to be complete: (stream, event recorded behind the solve kernel)
struct InFlight { cudaStream_t stream; cudaEvent_t ev; int device; bool active; };
constexpr int kInFlightSlots = 64;
InFlight g_inflight[kInFlightSlots];
std::mutex g_inflight_mu;

// true when another stream of `device` has a queue solve in flight (or when that cannot be determined)
bool others_in_flight(int device, cudaStream_t s) {
  bool busy = false;
  for (auto &f : g_inflight) {
    if (!f.active || f.device != device) continue;
    if (cudaEventQuery(f.ev) == cudaSuccess) { f.active = false; continue; }
    cudaGetLastError();   // cudaErrorNotReady is not an error
    if (f.stream != s) busy = true;
  }
  return busy;
}
void note_in_flight(int device, cudaStream_t s) {
  InFlight *slot = nullptr;
  for (auto &f : g_inflight)
    if (f.active && f.device == device && f.stream == s) { slot = &f; break; }   // a later solve on the same stream supersedes the earlier one
  if (!slot)
    for (auto &f : g_inflight)
      if (!f.active && (!f.ev || f.device == device)) { slot = &f; break; }
  if (!slot) return;   // table full: the next launch then simply sees fewer solves than there are
  if (!slot->ev) {
    if (cudaEventCreateWithFlags(&slot->ev, cudaEventDisableTiming) != cudaSuccess) { cudaGetLastError(); slot->ev = nullptr; return; }
  }
  slot->device = device; slot->stream = s;
  slot->active = cudaEventRecord(slot->ev, s) == cudaSuccess;
  if (!slot->active) cudaGetLastError();
}
constexpr int kTraceCap = 1 << 18;   // warp iterations recorded when the trace is on (8 words each: 8 MB)

int device_sms(int device) {
  static int cached[64] = {0};
  if (device >= 0 && device < 64 && cached[device]) return cached[device];
  int v = 148;
  cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device);
  if (device >= 0 && device < 64) cached[device] = v;
  return v;
}

struct Plan {
  int nwarps, NL, row_r4, ch2;
  unsigned cap;
  int64_t o_ctrl, o_ring, o_prob, o_traj, o_gain, o_trace, bytes;
};

Plan make_plan(const tfmpc_env *e, int64_t B, int T) {
  Plan p;
  const int n = e->n, m = e->m;
  const int chn = (n + m + 3) / 4;
  p.NL = ((T + 1) * chn + tq::CPH - 1) / tq::CPH;   // half-lines per trajectory row
  p.row_r4 = p.NL * tq::CPH;
  p.ch2 = (m * n + m + 1) / 2;
  p.nwarps = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)device_sms(e->device) * kMaxWarpsPerSM, B));   // the layout is sized for the most warps a launch may use
  p.cap = 1024;
  while (p.cap < 2u * (unsigned)B) p.cap <<= 1;
  auto al = [](int64_t v) { return (v + 255) / 256 * 256; };
  int64_t off = 0;
  p.o_ctrl = off; off += al((int64_t)tq::C_INTS * 4);
  p.o_ring = off; off += al((int64_t)p.cap * 8);
  p.o_prob = off; off += al(B * (int64_t)sizeof(tq::QProb));
  p.o_traj = off; off += al(2 * B * p.row_r4 * (int64_t)sizeof(R4));
  p.o_gain = off; off += al((int64_t)p.nwarps * T * p.ch2 * 32 * (int64_t)sizeof(tq::R2));
  p.o_trace = off; off += al((int64_t)kTraceCap * 32);
  p.bytes = off;
  return p;
}

template <int KIND, int N, int M>
int launch(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states, real *actions,
           real *costs, int32_t *stats, void *ws, int qp, cudaStream_t s) {
  const Plan pl = make_plan(e, B, T);
  char *base = (char *)ws;
  tq::QParams q;
  q.ctrl = (int *)(base + pl.o_ctrl);
  q.ring = (unsigned long long *)(base + pl.o_ring);
  q.ring_mask = pl.cap - 1;
  q.prob = (tq::QProb *)(base + pl.o_prob);
  q.traj = (R4 *)(base + pl.o_traj);
  q.gain = (tq::R2 *)(base + pl.o_gain);
  q.B = (int)B; q.T = T; q.row_r4 = pl.row_r4;
  // scheduling policy of this launch (see g_mode above)
  int mode = g_mode.load();
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  const bool capturing = cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone;
  if (capturing) cudaGetLastError();
  std::unique_lock<std::mutex> lock(g_inflight_mu, std::defer_lock);
  if (!capturing) lock.lock();
  if (mode != 1 && mode != 2) mode = (capturing || others_in_flight(e->device, s)) ? 1 : 2;
  g_last_mode.store(mode);
  const int sms = device_sms(e->device), wt = g_w_target.load(), ws_ = g_w_solo.load(), sm_ = g_solo_max.load(), wp_ = g_wps.load();
  const int wps = std::max(1, std::min(wp_ > 0 ? wp_ : (mode == 2 ? kLatencyWarpsPerSM : kMaxWarpsPerSM), kMaxWarpsPerSM));
  const int nwarps = (int)std::max<int64_t>(1, std::min<int64_t>((int64_t)sms * wps, B));   // warps (= CTAs) of this launch
  q.w_target = wt > 0 ? wt : (mode == 2 ? 15 * sms : std::max(1, sms / 4));
  q.patience = std::max(0, g_patience.load());
  q.solo_max = std::max(0, std::min(32, sm_ != 255 ? sm_ : (mode == 2 ? 1 : 0)));
  q.w_solo = q.solo_max > 0 ? (ws_ > 0 ? ws_ : (mode == 2 ? nwarps : q.w_target)) : 0;
  q.bulk = (mode == 1 && q.solo_max == 0 && g_drain_solo.load()) ? device_bulk_counter(e->device) : nullptr;   // throughput mode: solo only while the pipeline drains
  q.bulk_thr = 4 * q.w_target;
  q.watchdog_ns = 4000000000ull;   // 4 s without progress for one warp: give up (status TFMPC_ST_ABORTED) instead of hanging the device
  q.x0 = x0; q.u_init = u_init; q.states = states; q.actions = actions; q.costs = costs; q.stats = stats;
  q.trace = g_trace.load() ? (unsigned *)(base + pl.o_trace) : nullptr;
  q.trace_cap = kTraceCap;
  const int64_t init_items = std::max<int64_t>(std::max<int64_t>(B, (int64_t)pl.cap), tq::C_INTS);
  k_queue_init<<<(unsigned)((init_items + 255) / 256), 256, 0, s>>>(q, nwarps);
  LAUNCH_CHECK();
  constexpr bool can_close = M <= 2;
  if (can_close && qp == QP_CLOSED) k_queue_solve<KIND, N, M, (M <= 2 ? QP_CLOSED : QP_NEWTON)><<<nwarps, 32, 0, s>>>(e->es, o, q);
  else k_queue_solve<KIND, N, M, QP_NEWTON><<<nwarps, 32, 0, s>>>(e->es, o, q);
  LAUNCH_CHECK();
  if (!capturing) note_in_flight(e->device, s);
  return TFMPC_OK;
}

}  // namespace

// ticket counters are 32-bit and a problem takes at most a few hundred tickets: larger batches are solved in slices
constexpr int64_t kMaxSlice = 1 << 21;

int64_t queue_ilqr_workspace_bytes(const tfmpc_env *e, int64_t B, int T) { return make_plan(e, std::min(B, kMaxSlice), T).bytes; }

int queue_ilqr_option(const char *name, int value, int *previous) {
  std::atomic<int> *t = nullptr;
  if (!strcmp(name, "queue_warps_per_sm")) t = &g_wps;
  else if (!strcmp(name, "queue_w_target")) t = &g_w_target;
  else if (!strcmp(name, "queue_patience")) t = &g_patience;
  else if (!strcmp(name, "queue_trace")) t = &g_trace;
  else if (!strcmp(name, "queue_solo_max")) t = &g_solo_max;
  else if (!strcmp(name, "queue_w_solo")) t = &g_w_solo;
  else if (!strcmp(name, "queue_mode")) t = &g_mode;
  else if (!strcmp(name, "queue_drain_solo")) t = &g_drain_solo;
  else if (!strcmp(name, "queue_last_mode")) { *previous = g_last_mode.load(); return 1; }
  if (!t) return 0;
  *previous = t->exchange(value);
  return 1;
}

// control block of the last solve in this workspace (diagnostics: warp iterations, lanes, rounds, ...)
int queue_ilqr_counters(const void *ws, int *out, int n, cudaStream_t s) {
  CUDA_TRY(cudaMemcpyAsync(out, ws, sizeof(int) * std::min(n, (int)tq::C_INTS), cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return TFMPC_OK;
}

// scheduling trace of the last solve in this workspace (option "queue_trace" must have been on): up to max_records records of 8 uint32
int64_t queue_ilqr_trace(const tfmpc_env *e, int64_t B, int T, const void *ws, unsigned *out, int64_t max_records, cudaStream_t s) {
  const Plan pl = make_plan(e, std::min(B, kMaxSlice), T);
  int n = 0;
  CUDA_TRY(cudaMemcpyAsync(&n, (const char *)ws + pl.o_ctrl + 4 * tq::C_TRACE, 4, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  int64_t cnt = std::min<int64_t>(std::min<int64_t>(n, kTraceCap), max_records);
  if (cnt > 0) CUDA_TRY(cudaMemcpyAsync(out, (const char *)ws + pl.o_trace, cnt * 32, cudaMemcpyDeviceToHost, s));
  CUDA_TRY(cudaStreamSynchronize(s));
  return cnt;
}

int queue_ilqr_solve(const tfmpc_env *e, int64_t B, int T, const real *x0, const real *u_init, const IlqrOpts &o, real *states,
                     real *actions, real *costs, int32_t *stats, void *ws, int64_t ws_bytes, int qp, cudaStream_t s) {
  if (ws_bytes < queue_ilqr_workspace_bytes(e, B, T)) return tfmpc_set_error(TFMPC_E_WORKSPACE, "workspace too small");
  const int n = e->n, m = e->m;
  for (int64_t off = 0; off < B; off += kMaxSlice) {   // slices run one after another in stream order and share the workspace
    const int64_t Bs = std::min(kMaxSlice, B - off);
    const real *x0s = x0 + off * n, *u0s = u_init + off * T * m;
    real *ss = states + off * (T + 1) * n, *as = actions + off * T * m, *cs = costs + off * (T + 1);
    int32_t *sts = stats + off * 4;
    int rc;
#define CALL(KD, N, M) rc = launch<KD, N, M>(e, Bs, T, x0s, u0s, o, ss, as, cs, sts, ws, qp, s)
    if (e->kind == TFMPC_ENV_NAVIGATION && n == 2 && e->nz <= 2) CALL(TFMPC_ENV_NAVIGATION_Z2, 2, 2);
    else if (e->kind == TFMPC_ENV_NAVIGATION && n == 2) CALL(TFMPC_ENV_NAVIGATION, 2, 2);
    else if (e->kind == TFMPC_ENV_NAVLQR && n == 1) CALL(TFMPC_ENV_NAVLQR, 1, 1);
    else if (e->kind == TFMPC_ENV_NAVLQR && n == 2) CALL(TFMPC_ENV_NAVLQR, 2, 2);
    else if (e->kind == TFMPC_ENV_NAVLQR && n == 3) CALL(TFMPC_ENV_NAVLQR, 3, 3);
    else if (e->kind == TFMPC_ENV_NAVLQR && n == 4) CALL(TFMPC_ENV_NAVLQR, 4, 4);
    else return tfmpc_set_error(TFMPC_E_UNSUPPORTED, "no thread-per-problem kernel for kind=%d n=%d", e->kind, n);
#undef CALL
    if (rc) return rc;
  }
  return TFMPC_OK;
}
