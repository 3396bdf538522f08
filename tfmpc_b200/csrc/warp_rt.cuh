// Warp runtime shim: the handful of warp-level primitives the persistent queue solver (queue_core.cuh) is written
// against.  Compiled by nvcc it is a zero-cost wrapper over the sm_100a intrinsics (shuffles, votes, cp.async,
// acquire / release accesses, nanosleep).  Compiled by the host compiler (tests/host_emulation only -- TEST
// INFRASTRUCTURE, never part of the shipped library) every lane is an OS thread and every warp collective a barrier, so
// the GPU-less build container can execute the very same scheduling code: queue protocol, line-search rounds,
// cooperative line staging.  A lane that skips a collective deadlocks the emulation exactly as it would hang the GPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define WD __device__ __forceinline__
#define WD_NOINLINE __device__ __noinline__

struct WarpRT {
  int lane;
  __device__ WarpRT() : lane((int)(threadIdx.x & 31u)) {}
  static constexpr unsigned FULL = 0xffffffffu;
  template <class V> WD V shfl(V v, int src) const { return __shfl_sync(FULL, v, src); }
  template <class P> WD P *shfl_ptr(P *p, int src) const { return (P *)(uintptr_t)__shfl_sync(FULL, (unsigned long long)(uintptr_t)p, src); }
  WD unsigned ballot(bool p) const { return __ballot_sync(FULL, p); }
  WD bool any(bool p) const { return __any_sync(FULL, p) != 0; }
  WD void syncwarp() const { __syncwarp(FULL); }
  // 16-byte asynchronous copy global -> shared (LDGSTS, bypasses L1)
  WD void cp_async16(void *smem_dst, const void *gsrc) const {
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gsrc) : "memory");
  }
  WD void cp_commit() const { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
  template <int PENDING> WD void cp_wait() const { asm volatile("cp.async.wait_group %0;\n" ::"n"(PENDING) : "memory"); }
  WD int atomic_add(int *p, int v) const { return atomicAdd(p, v); }
  WD int atomic_cas(int *p, int cmp, int v) const { return atomicCAS(p, cmp, v); }
  WD int ld_relaxed(const int *p) const {
    int v;
    asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
  }
  WD int ld_acquire(const int *p) const {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
    return v;
  }
  WD void st_relaxed(int *p, int v) const { asm volatile("st.relaxed.gpu.global.s32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory"); }
  WD unsigned long long ld_acquire64(const unsigned long long *p) const {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];\n" : "=l"(v) : "l"(p) : "memory");
    return v;
  }
  WD void st_release64(unsigned long long *p, unsigned long long v) const {
    asm volatile("st.release.gpu.global.u64 [%0], %1;\n" ::"l"(p), "l"(v) : "memory");
  }
  WD void fence() const { asm volatile("fence.acq_rel.gpu;\n" ::: "memory"); }
  WD void sleep_ns(unsigned ns) const { __nanosleep(ns); }
  WD unsigned long long now_ns() const {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;\n" : "=l"(t));
    return t;
  }
};

#else  // ------------------------------------------------------------------ host emulation (tests only)
#include <pthread.h>
#include <sched.h>
#include <string.h>
#include <time.h>
#define WD inline
#define WD_NOINLINE inline

struct WarpShared {   // one per emulated warp
  pthread_barrier_t bar;
  unsigned long long xchg[2][32];
};

struct WarpRT {
  int lane;
  WarpShared *ws;
  unsigned gen;
  WarpRT(int lane_, WarpShared *ws_) : lane(lane_), ws(ws_), gen(0) {}
  template <class V> V shfl(V v, int src) {
    static_assert(sizeof(V) <= 8, "shfl: at most 64 bits");
    unsigned long long w = 0;
    memcpy(&w, &v, sizeof(V));
    unsigned long long *x = ws->xchg[gen++ & 1];
    x[lane] = w;
    pthread_barrier_wait(&ws->bar);
    w = x[src & 31];
    V r;
    memcpy(&r, &w, sizeof(V));
    return r;
  }
  template <class P> P *shfl_ptr(P *p, int src) { return (P *)(uintptr_t)shfl((unsigned long long)(uintptr_t)p, src); }
  unsigned ballot(bool p) {
    unsigned long long *x = ws->xchg[gen++ & 1];
    x[lane] = p ? 1 : 0;
    pthread_barrier_wait(&ws->bar);
    unsigned m = 0;
    for (int i = 0; i < 32; i++) m |= (unsigned)(x[i] & 1) << i;
    return m;
  }
  bool any(bool p) { return ballot(p) != 0; }
  void syncwarp() { pthread_barrier_wait(&ws->bar); }
  void cp_async16(void *smem_dst, const void *gsrc) { memcpy(smem_dst, gsrc, 16); }   // completes at once
  void cp_commit() {}
  template <int PENDING> void cp_wait() {}
  int atomic_add(int *p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
  int atomic_cas(int *p, int cmp, int v) { __atomic_compare_exchange_n(p, &cmp, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST); return cmp; }
  int ld_relaxed(const int *p) { return __atomic_load_n(p, __ATOMIC_RELAXED); }
  int ld_acquire(const int *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
  void st_relaxed(int *p, int v) { __atomic_store_n(p, v, __ATOMIC_RELAXED); }
  unsigned long long ld_acquire64(const unsigned long long *p) { return __atomic_load_n(p, __ATOMIC_ACQUIRE); }
  void st_release64(unsigned long long *p, unsigned long long v) { __atomic_store_n(p, v, __ATOMIC_RELEASE); }
  void fence() { __atomic_thread_fence(__ATOMIC_SEQ_CST); }
  void sleep_ns(unsigned) { sched_yield(); }
  unsigned long long now_ns() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (unsigned long long)ts.tv_sec * 1000000000ull + (unsigned long long)ts.tv_nsec;
  }
};
#endif
