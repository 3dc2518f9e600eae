// DEVELOPMENT TOOL — NOT A PRODUCT PATH, NOT A CPU FALLBACK.
//
// A tiny single-OS-thread emulation of the CUDA execution model (grid of blocks,
// threads as cooperatively scheduled fibers, __syncthreads, warp shuffles,
// atomics, static/dynamic shared memory).  It exists because the build container
// has no GPU: compiling the plain-CUDA kernels of carca_replication_b200/csrc as
// host C++ against this header lets their indexing / synchronisation logic be
// checked against the oracle (and under -fsanitize=address) before a B200 minute
// is spent.  Only tools/emu/build_emu.py and tests/test_emu_*.py use it; the
// product loader (carca_replication_b200/_native.py) knows nothing about it and
// raises if the real CUDA library is missing.  tcgen05/TMA kernels are not
// emulated (they are compiled out under CARCA_EMU).
#pragma once
#include <ucontext.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#if defined(__SANITIZE_ADDRESS__)
#include <sanitizer/common_interface_defs.h>
#endif

#define CARCA_EMU 1
#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))
#define __constant__ static

struct uint3 { unsigned x, y, z; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct __attribute__((aligned(16))) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct __attribute__((aligned(16))) int4 { int x, y, z, w; };
struct uint2 { unsigned x, y; };
struct __attribute__((aligned(16))) uint4 { unsigned x, y, z, w; };
static inline float2 make_float2(float a, float b) { return {a, b}; }
static inline float4 make_float4(float a, float b, float c, float d) { return {a, b, c, d}; }
static inline int4 make_int4(int a, int b, int c, int d) { return {a, b, c, d}; }
static inline uint4 make_uint4(unsigned a, unsigned b, unsigned c, unsigned d) { return {a, b, c, d}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
static inline cudaError_t cudaGetLastError() { return cudaSuccess; }
static inline cudaError_t cudaPeekAtLastError() { return cudaSuccess; }
static inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
static inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return 0; }
enum cudaMemcpyKind { cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
static inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) {
  memmove(d, s, n);
  return 0;
}
static inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }

namespace emu {

constexpr size_t kStackBytes = 256 * 1024;

struct Warp {
  uint64_t buf[32];
  int nlanes = 0, arrived = 0, gen = 0;
};

struct Fiber {
  ucontext_t ctx;
  char* stack = nullptr;
  uint3 tid{0, 0, 0};
  int linear = 0;
  bool done = false;
  void* fake_stack = nullptr;
};

struct State {
  ucontext_t sched;
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Fiber* cur = nullptr;
  uint3 bid{0, 0, 0};
  dim3 bdim, gdim;
  int block_arrived = 0, block_gen = 0, nthreads = 0;
  long events = 0;  // barrier arrivals + fiber exits; a sweep with none is a deadlock
  std::vector<unsigned char> dyn_smem;
  const std::function<void()>* body = nullptr;
  std::vector<char*> stack_pool;
};

inline State& S() {
  static State s;
  return s;
}

inline void yield() {
  State& s = S();
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_start_switch_fiber(&s.cur->fake_stack, nullptr, 0);
#endif
  swapcontext(&s.cur->ctx, &s.sched);
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(s.cur->fake_stack, nullptr, nullptr);
#endif
}

inline void fiber_entry() {
  State& s = S();
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_finish_switch_fiber(nullptr, nullptr, nullptr);
#endif
  (*s.body)();
  s.cur->done = true;
#if defined(__SANITIZE_ADDRESS__)
  __sanitizer_start_switch_fiber(nullptr, nullptr, 0);
#endif
  swapcontext(&s.cur->ctx, &s.sched);
}

inline void block_barrier() {
  State& s = S();
  int my = s.block_gen;
  s.events++;
  if (++s.block_arrived == s.nthreads) {
    s.block_arrived = 0;
    s.block_gen++;
    return;
  }
  while (s.block_gen == my) yield();
}

inline void warp_barrier() {
  State& s = S();
  Warp& w = s.warps[s.cur->linear / 32];
  int my = w.gen;
  s.events++;
  if (++w.arrived == w.nlanes) {
    w.arrived = 0;
    w.gen++;
    return;
  }
  while (w.gen == my) yield();
}

inline void run_block() {
  State& s = S();
  const int n = s.nthreads;
  s.block_arrived = 0;
  for (auto& w : s.warps) w.arrived = 0;
  for (int i = 0; i < n; ++i) {
    Fiber& f = s.fibers[i];
    f.done = false;
    getcontext(&f.ctx);
    f.ctx.uc_stack.ss_sp = f.stack;
    f.ctx.uc_stack.ss_size = kStackBytes;
    f.ctx.uc_link = &s.sched;
    makecontext(&f.ctx, (void (*)())fiber_entry, 0);
  }
  int remaining = n;
  while (remaining) {
    long before = s.events;
    for (int i = 0; i < n; ++i) {
      Fiber& f = s.fibers[i];
      if (f.done) continue;
      s.cur = &f;
#if defined(__SANITIZE_ADDRESS__)
      void* fake = nullptr;
      __sanitizer_start_switch_fiber(&fake, f.stack, kStackBytes);
#endif
      swapcontext(&s.sched, &f.ctx);
#if defined(__SANITIZE_ADDRESS__)
      __sanitizer_finish_switch_fiber(fake, nullptr, nullptr);
#endif
      if (f.done) { --remaining; s.events++; }
    }
    // a full sweep with no barrier arrival and no exit: every live fiber is parked on a
    // barrier that can never complete (divergent __syncthreads / partial-warp shuffle)
    if (remaining && s.events == before) {
      fprintf(stderr, "cuda_emu: deadlock (divergent __syncthreads / shuffle?) in block (%u,%u,%u)\n", s.bid.x,
              s.bid.y, s.bid.z);
      abort();
    }
  }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  State& s = S();
  s.gdim = grid;
  s.bdim = block;
  s.nthreads = int(block.x * block.y * block.z);
  s.body = &body;
  s.dyn_smem.assign(smem + 64, 0xCD);
  if ((int)s.fibers.size() < s.nthreads) {
    size_t old = s.fibers.size();
    s.fibers.resize(s.nthreads);
    for (size_t i = old; i < s.fibers.size(); ++i) s.fibers[i].stack = (char*)malloc(kStackBytes);
  }
  s.warps.assign((s.nthreads + 31) / 32, Warp());
  for (int i = 0; i < s.nthreads; ++i) {
    Fiber& f = s.fibers[i];
    f.linear = i;
    f.tid = {unsigned(i % block.x), unsigned((i / block.x) % block.y), unsigned(i / (block.x * block.y))};
    s.warps[i / 32].nlanes++;
  }
  for (unsigned bz = 0; bz < grid.z; ++bz)
    for (unsigned by = 0; by < grid.y; ++by)
      for (unsigned bx = 0; bx < grid.x; ++bx) {
        s.bid = {bx, by, bz};
        run_block();
      }
  s.body = nullptr;
}

inline void* dyn_smem_ptr() {
  uintptr_t p = (uintptr_t)S().dyn_smem.data();
  return (void*)((p + 63) & ~uintptr_t(63));
}

}  // namespace emu

#define threadIdx (emu::S().cur->tid)
#define blockIdx (emu::S().bid)
#define blockDim (emu::S().bdim)
#define gridDim (emu::S().gdim)
#define warpSize 32

static inline void __syncthreads() { emu::block_barrier(); }
static inline void __syncwarp(unsigned = 0xffffffffu) { emu::warp_barrier(); }
static inline void __threadfence() {}
static inline void __threadfence_block() {}

template <class T>
static inline T emu_shfl_from(T v, int src_lane) {
  emu::State& s = emu::S();
  emu::Warp& w = s.warps[s.cur->linear / 32];
  int lane = s.cur->linear % 32;
  static_assert(sizeof(T) <= 8, "shuffle payload");
  uint64_t raw = 0;
  memcpy(&raw, &v, sizeof(T));
  w.buf[lane] = raw;
  emu::warp_barrier();
  if (src_lane < 0 || src_lane >= w.nlanes) src_lane = lane;
  uint64_t got = w.buf[src_lane];
  emu::warp_barrier();
  T r;
  memcpy(&r, &got, sizeof(T));
  return r;
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int lane_mask, int width = 32) {
  int lane = emu::S().cur->linear % 32;
  (void)width;
  return emu_shfl_from(v, lane ^ lane_mask);
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src, int width = 32) {
  int lane = emu::S().cur->linear % 32;
  return emu_shfl_from(v, (lane / width) * width + (src % width));
}
template <class T>
static inline T __shfl_down_sync(unsigned, T v, unsigned delta, int width = 32) {
  int lane = emu::S().cur->linear % 32;
  int src = lane + (int)delta;
  if ((src / width) != (lane / width)) src = lane;
  return emu_shfl_from(v, src);
}
static inline unsigned __ballot_sync(unsigned, int pred) {
  unsigned bits = 0;
  int lane = emu::S().cur->linear % 32;
  for (int i = 0; i < 32; ++i) {
    int p = emu_shfl_from(pred, i);
    if (p && i < emu::S().warps[emu::S().cur->linear / 32].nlanes) bits |= (1u << i);
  }
  (void)lane;
  return bits;
}
static inline int __any_sync(unsigned m, int pred) { return __ballot_sync(m, pred) != 0; }
static inline int __all_sync(unsigned m, int pred) {
  unsigned b = __ballot_sync(m, pred);
  int n = emu::S().warps[emu::S().cur->linear / 32].nlanes;
  return b == (n == 32 ? 0xffffffffu : ((1u << n) - 1));
}

template <class T>
static inline T atomicAdd(T* p, T v) {
  T old = *p;
  *p = old + v;
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) {
  unsigned long long o = *p;
  *p = o + v;
  return o;
}
static inline int atomicMax(int* p, int v) {
  int o = *p;
  *p = std::max(o, v);
  return o;
}

template <class T>
static inline T __ldg(const T* p) { return *p; }
#define __expf expf
#define __logf logf
#define __log2f log2f
static inline float rsqrtf(float x) { return 1.0f / sqrtf(x); }
static inline float __fdividef(float a, float b) { return a / b; }
static inline float __frcp_rn(float a) { return 1.0f / a; }
static inline float __fmaf_rn(float a, float b, float c) { return fmaf(a, b, c); }
static inline unsigned __float_as_uint(float f) { unsigned u; memcpy(&u, &f, 4); return u; }
static inline float __uint_as_float(unsigned u) { float f; memcpy(&f, &u, 4); return f; }
static inline int __float_as_int(float f) { int u; memcpy(&u, &f, 4); return u; }
static inline float __int_as_float(int u) { float f; memcpy(&f, &u, 4); return f; }
static inline unsigned __umulhi(unsigned a, unsigned b) { return (unsigned)(((uint64_t)a * b) >> 32); }
static inline int __popc(unsigned x) { return __builtin_popcount(x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline int __clz(int x) { return x ? __builtin_clz((unsigned)x) : 32; }
using std::max;
using std::min;

#define CARCA_LAUNCH(kfn, grid, block, smem, stream, ...) \
  emu::launch((grid), (block), (smem), [=]() { kfn(__VA_ARGS__); })
#define CARCA_DYN_SMEM(type, name) type* name = reinterpret_cast<type*>(emu::dyn_smem_ptr())
