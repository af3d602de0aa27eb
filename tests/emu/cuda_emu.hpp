// cuda_emu.hpp — a small SIMT emulator for running this repository's __device__ kernel bodies on the CPU.
//
// TEST INFRASTRUCTURE ONLY (like oracle/): it exists because device code that has not run on a GPU yet still has to be
// exercised somehow — same source, real concurrency.  One OS thread per WARP of a block, its 32 lanes are fibers that
// the warp thread switches between at every collective / barrier / spin-wait (so a warp collective costs a few dozen
// user-level context switches instead of a 32-thread OS barrier); warps run concurrently, blocks one after the other.
// -DCOZO_EMU_LANE_THREADS selects the older, slower mode with one OS thread per lane (every lane truly concurrent).
// __syncthreads() / __syncwarp() are barriers, warp collectives exchange values through a per-warp buffer (all lanes
// of the warp must take part, as with a full mask on the device: a divergent collective never completes here and is a
// bug there), atomics are host atomics, mbarrier + cp.async.bulk are modelled as "copy now, flip the phase when the
// expected bytes have arrived".  A watchdog aborts a launch that does not finish (a hang on the device).
// Nothing under cozo_b200/ includes this file; kernels never see it unless a test defines COZO_CPU_EMU.
#pragma once
#include <atomic>
#include <barrier>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <memory>
#include <mutex>
#include <thread>
#include <vector>

#define COZO_CPU_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static   /* blocks run one at a time: a static is "per block" */
#define __align__(x) alignas(x)

struct emu_dim3 {
  unsigned x = 1, y = 1, z = 1;
  emu_dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
using dim3 = emu_dim3;
struct uint4 {
  uint32_t x, y, z, w;
};
struct float4 {
  float x, y, z, w;
};
inline uint4 make_uint4(uint32_t a, uint32_t b, uint32_t c, uint32_t d) { return uint4{a, b, c, d}; }

#include <sys/mman.h>

namespace emu {
#ifdef COZO_EMU_LANE_THREADS
struct WarpState {
  std::unique_ptr<std::barrier<>> bar;
  unsigned long long buf[32];
};
struct BlockState {
  std::unique_ptr<std::barrier<>> bar;
  std::vector<WarpState> warps;
};
inline thread_local emu_dim3 t_threadIdx, t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockState* t_block = nullptr;
inline WarpState& my_warp() { return t_block->warps[t_threadIdx.x >> 5]; }
inline void warp_sync() { my_warp().bar->arrive_and_wait(); }
inline void block_sync() { t_block->bar->arrive_and_wait(); }
inline void lane_yield() { std::this_thread::yield(); }
#define EMU_TID (emu::t_threadIdx)
#else
// ---- fibers: a warp is one OS thread, its lanes are cooperatively scheduled contexts on stacks of their own ----------
extern "C" void cozo_emu_switch(void** save_sp, void* load_sp);
#if defined(__x86_64__)
asm(R"(
.text
.weak cozo_emu_switch
.hidden cozo_emu_switch
.type cozo_emu_switch,@function
cozo_emu_switch:
  pushq %rbp
  pushq %rbx
  pushq %r12
  pushq %r13
  pushq %r14
  pushq %r15
  movq %rsp, (%rdi)
  movq %rsi, %rsp
  popq %r15
  popq %r14
  popq %r13
  popq %r12
  popq %rbx
  popq %rbp
  ret
.size cozo_emu_switch,.-cozo_emu_switch
)");
#else
#error "the fiber mode of cuda_emu.hpp is written for x86-64; build with -DCOZO_EMU_LANE_THREADS elsewhere"
#endif
struct Fiber {
  void* sp = nullptr;
  uint8_t* stack = nullptr;
  bool done = false;
  unsigned wait_gen = ~0u;  // blocked at the warp barrier of this generation (~0u: runnable)
  emu_dim3 tid;
};
struct WarpState {
  Fiber lane[32];
  unsigned n = 0, alive = 0, arrived = 0, gen = 0;
  void* sched_sp = nullptr;
  unsigned long long buf[32];
  const std::function<void()>* body = nullptr;
};
struct BlockState {
  std::unique_ptr<std::barrier<>> bar;  // one participant per warp thread
  std::vector<WarpState> warps;
};
inline thread_local emu_dim3 t_blockIdx, t_blockDim, t_gridDim;
inline thread_local BlockState* t_block = nullptr;
inline thread_local WarpState* t_warp = nullptr;
inline thread_local Fiber* t_fiber = nullptr;
constexpr size_t kFiberStack = 256 << 10;
inline WarpState& my_warp() { return *t_warp; }
// back to the warp's scheduler, which resumes the next runnable lane
inline void lane_yield() { cozo_emu_switch(&t_fiber->sp, t_warp->sched_sp); }
[[noreturn]] inline void collective_after_exit() {
  std::fprintf(stderr, "emu: a warp collective / barrier was reached after a lane of the warp had exited (undefined on the device)\n");
  std::abort();
}
inline void warp_barrier(bool block_wide) {
  WarpState& w = *t_warp;
  if (w.alive != w.n) collective_after_exit();
  const unsigned my = w.gen;
  if (++w.arrived == w.alive) {
    if (block_wide) t_block->bar->arrive_and_wait();  // the last lane of every warp meets the other warps
    w.arrived = 0;
    ++w.gen;
    return;
  }
  t_fiber->wait_gen = my;
  while (w.gen == my) lane_yield();
  t_fiber->wait_gen = ~0u;
}
inline void warp_sync() { warp_barrier(false); }
inline void block_sync() { warp_barrier(true); }
inline void fiber_main() {
  Fiber* f = t_fiber;
  (*t_warp->body)();
  f->done = true;
  --t_warp->alive;
  if (t_warp->arrived && t_warp->arrived == t_warp->alive) collective_after_exit();
  lane_yield();
  std::abort();  // a finished lane is never resumed
}
inline void fiber_reset(Fiber& f) {
  void** top = reinterpret_cast<void**>((reinterpret_cast<uintptr_t>(f.stack) + kFiberStack) & ~(uintptr_t)15);
  void** sp = top - 8;  // six callee-saved registers, the entry address, one slot of padding (16-byte aligned frame)
  for (int i = 0; i < 6; ++i) sp[i] = nullptr;
  sp[6] = reinterpret_cast<void*>(&fiber_main);
  sp[7] = nullptr;
  f.sp = sp;
  f.done = false;
  f.wait_gen = ~0u;
}
// run one block's worth of this warp: round-robin over the lanes until all of them have returned
inline void warp_run(WarpState& w) {
  t_warp = &w;
  w.alive = w.n;
  w.arrived = 0;
  for (unsigned l = 0; l < w.n; ++l) fiber_reset(w.lane[l]);
  unsigned left = w.n;
  while (left) {
    for (unsigned l = 0; l < w.n; ++l) {
      Fiber& f = w.lane[l];
      if (f.done || (f.wait_gen != ~0u && f.wait_gen == w.gen)) continue;
      t_fiber = &f;
      cozo_emu_switch(&w.sched_sp, f.sp);
      if (f.done) --left;
    }
  }
  t_fiber = nullptr;
}
#define EMU_TID (emu::t_fiber->tid)
#endif
inline std::atomic<long long> g_progress{0};
}  // namespace emu

#define threadIdx EMU_TID
#define blockIdx (emu::t_blockIdx)
#define blockDim (emu::t_blockDim)
#define gridDim (emu::t_gridDim)

inline void __syncthreads() { emu::block_sync(); }
inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_sync(); }

template <class T>
inline T emu_exchange(T v, int src_lane) {
  static_assert(sizeof(T) <= 8, "shuffle payload");
  emu::WarpState& w = emu::my_warp();
  const int lane = threadIdx.x & 31;
  unsigned long long raw = 0;
  std::memcpy(&raw, &v, sizeof(T));
  w.buf[lane] = raw;
  emu::warp_sync();
  T out = v;
  if (src_lane >= 0 && src_lane < 32) {
    unsigned long long r = w.buf[src_lane];
    std::memcpy(&out, &r, sizeof(T));
  }
  emu::warp_sync();
  return out;
}
template <class T>
inline T __shfl_sync(unsigned, T v, int src) { return emu_exchange(v, src & 31); }
template <class T>
inline T __shfl_xor_sync(unsigned, T v, int m) { return emu_exchange(v, (int)((threadIdx.x & 31) ^ m)); }
template <class T>
inline T __shfl_up_sync(unsigned, T v, unsigned d) {
  const int lane = threadIdx.x & 31;
  return emu_exchange(v, lane >= (int)d ? lane - (int)d : lane);
}
inline unsigned __ballot_sync(unsigned, int pred) {
  emu::WarpState& w = emu::my_warp();
  const int lane = threadIdx.x & 31;
  w.buf[lane] = pred ? 1ull : 0ull;
  emu::warp_sync();
  unsigned m = 0;
  const unsigned nl = (blockDim.x - (threadIdx.x & ~31u)) < 32 ? (blockDim.x - (threadIdx.x & ~31u)) : 32;
  for (unsigned l = 0; l < nl; ++l) m |= (unsigned)(w.buf[l] & 1ull) << l;
  emu::warp_sync();
  return m;
}
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __ffs(unsigned x) { return __builtin_ffs((int)x); }
inline float __uint_as_float(uint32_t u) { float f; std::memcpy(&f, &u, 4); return f; }
inline uint32_t __float_as_uint(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }
inline float __int_as_float(int i) { float f; std::memcpy(&f, &i, 4); return f; }
inline unsigned long long __double2ull_rn(double d) { return (unsigned long long)std::llrint(d); }
template <class T> inline T __ldg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <class T> inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_block() { std::atomic_thread_fence(std::memory_order_seq_cst); }
inline void __threadfence_system() { std::atomic_thread_fence(std::memory_order_seq_cst); }
// a polite spin: let the sibling lanes run, and every so often give the core away — a warp whose lanes all poll a flag
// that another rank / warp has yet to write must not starve that writer on a loaded machine
inline void __nanosleep(unsigned) {
  static thread_local unsigned polls = 0;
  emu::lane_yield();
  if ((++polls & 1023u) == 0) std::this_thread::sleep_for(std::chrono::microseconds(50));
}
inline void __trap() { std::fprintf(stderr, "emu: __trap()\n"); std::abort(); }
// device code only uses clock64() for watchdogs ("a peer died"): tick 16x slower than nanoseconds, so that a ~20 s limit
// in GPU cycles becomes ~10 minutes here and a loaded test machine cannot trip it
inline long long clock64() { return std::chrono::steady_clock::now().time_since_epoch().count() / 16; }
using std::isfinite;
using std::isinf;
using std::isnan;
using std::max;
using std::min;

template <class T> inline T atomicAdd(T* p, T v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline double atomicAdd(double* p, double v) {
  unsigned long long* u = reinterpret_cast<unsigned long long*>(p);
  unsigned long long old = __atomic_load_n(u, __ATOMIC_SEQ_CST);
  for (;;) {
    double d; std::memcpy(&d, &old, 8); d += v;
    unsigned long long nw; std::memcpy(&nw, &d, 8);
    if (__atomic_compare_exchange_n(u, &old, nw, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) { double r; std::memcpy(&r, &old, 8); return r; }
  }
}
template <class T> inline T atomicOr(T* p, T v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicAnd(T* p, T v) { return __atomic_fetch_and(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicExch(T* p, T v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
template <class T> inline T atomicCAS(T* p, T cmp, T val) {
  __atomic_compare_exchange_n(p, &cmp, val, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST);
  return cmp;
}

namespace emu {
// Run `body()` for every thread of every block of the grid: `threads` OS threads walk the blocks one after the other
// (a barrier separates consecutive blocks, so a `static` stands in for __shared__).  A watchdog ends the process when a
// launch does not finish — a deadlocked barrier or an endless loop, i.e. a hang on the device.
inline bool launch(dim3 grid, unsigned threads, const std::function<void()>& body, double timeout_s = 60.0, const char* name = "kernel") {
  BlockState bs;
  const unsigned nw = (threads + 31) / 32;
  std::atomic<unsigned> done{0};
  std::atomic<unsigned> cur_bx{0}, cur_by{0};
  std::vector<std::thread> th;
#ifdef COZO_EMU_LANE_THREADS
  bs.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)threads);
  bs.warps.resize(nw);
  for (unsigned w = 0; w < nw; ++w)
    bs.warps[w].bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)std::min(32u, threads - w * 32));
  std::barrier<> between((std::ptrdiff_t)threads);
  const unsigned n_os = threads;
  th.reserve(threads);
  for (unsigned t = 0; t < threads; ++t)
    th.emplace_back([&, t] {
      t_threadIdx = emu_dim3(t, 0, 0);
      t_blockDim = emu_dim3(threads, 1, 1);
      t_gridDim = grid;
      t_block = &bs;
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = emu_dim3(bx, by, 0);
          if (t == 0) {
            cur_bx = bx;
            cur_by = by;
          }
          body();
          between.arrive_and_wait();
        }
      done.fetch_add(1);
    });
#else
  bs.bar = std::make_unique<std::barrier<>>((std::ptrdiff_t)nw);
  bs.warps.resize(nw);
  uint8_t* stacks = static_cast<uint8_t*>(mmap(nullptr, (size_t)threads * kFiberStack, PROT_READ | PROT_WRITE,
                                               MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0));
  if (stacks == MAP_FAILED) {
    std::fprintf(stderr, "emu: cannot map %u lane stacks\n", threads);
    std::abort();
  }
  for (unsigned w = 0; w < nw; ++w) {
    WarpState& ws = bs.warps[w];
    ws.n = std::min(32u, threads - w * 32);
    ws.body = &body;
    for (unsigned l = 0; l < ws.n; ++l) {
      ws.lane[l].stack = stacks + (size_t)(w * 32 + l) * kFiberStack;
      ws.lane[l].tid = emu_dim3(w * 32 + l, 0, 0);
    }
  }
  std::barrier<> between((std::ptrdiff_t)nw);
  const unsigned n_os = nw;
  th.reserve(nw);
  for (unsigned w = 0; w < nw; ++w)
    th.emplace_back([&, w] {
      t_blockDim = emu_dim3(threads, 1, 1);
      t_gridDim = grid;
      t_block = &bs;
      for (unsigned by = 0; by < grid.y; ++by)
        for (unsigned bx = 0; bx < grid.x; ++bx) {
          t_blockIdx = emu_dim3(bx, by, 0);
          if (w == 0) {
            cur_bx = bx;
            cur_by = by;
          }
          warp_run(bs.warps[w]);
          between.arrive_and_wait();
        }
      done.fetch_add(1);
    });
#endif
  const auto t0 = std::chrono::steady_clock::now();
  unsigned naps = 0;
  while (done.load() < n_os) {
    if (++naps < 2000) std::this_thread::yield();
    else std::this_thread::sleep_for(std::chrono::microseconds(200));
    if ((naps & 255) == 0 && std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > timeout_s) {
      std::fprintf(stderr, "emu: %s did not finish within %.0f s (at block (%u,%u)): HANG (deadlocked barrier or endless loop)\n", name,
                   timeout_s, cur_bx.load(), cur_by.load());
      std::_Exit(3);
    }
  }
  for (auto& x : th) x.join();
#ifndef COZO_EMU_LANE_THREADS
  munmap(stacks, (size_t)threads * kFiberStack);
#endif
  return true;
}
}  // namespace emu

// ---- mbarrier + 1-D bulk copy (what common.cuh implements with PTX) ---------------------------------------------------
// bar = { low 32 bits: pending transaction bytes + (1 << 31) while an arrival is outstanding ; bit 32: phase }
inline void mbar_init(uint64_t* bar, uint32_t) { __atomic_store_n(bar, 0ull, __ATOMIC_SEQ_CST); }
inline void mbar_fence_init() {}
inline void mbar_expect_tx(uint64_t* bar, uint32_t bytes) { __atomic_fetch_add(bar, (uint64_t)bytes, __ATOMIC_SEQ_CST); }
inline bool mbar_try_wait(uint64_t* bar, uint32_t parity) { return ((__atomic_load_n(bar, __ATOMIC_SEQ_CST) >> 32) & 1ull) != parity; }
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  unsigned long spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    emu::lane_yield();  // the lane that issues the copy may be a sibling that has not run yet
    if (++spins > (1ul << 28)) __trap();
  }
}
inline void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  if (((uintptr_t)gmem_src & 15) || (bytes & 15)) { std::fprintf(stderr, "emu: cp.async.bulk needs 16-byte aligned source and size\n"); std::abort(); }
  std::memcpy(smem_dst, gmem_src, bytes);
  const uint64_t left = __atomic_sub_fetch(bar, (uint64_t)bytes, __ATOMIC_SEQ_CST) & 0xFFFFFFFFull;
  if (left == 0) __atomic_fetch_xor(bar, 1ull << 32, __ATOMIC_SEQ_CST);  // all expected bytes have landed: phase completes
}
inline void fence_proxy_async_smem() {}
