// Shared device/host helpers for the CARCA sm_100a kernels.
#pragma once
#ifndef CARCA_EMU
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#define CARCA_LAUNCH(kfn, grid, block, smem, stream, ...) kfn<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define CARCA_DYN_SMEM(type, name)                                   \
  extern __shared__ __align__(16) unsigned char _carca_dyn_smem[];   \
  type* name = reinterpret_cast<type*>(_carca_dyn_smem)
#endif
#include <cstdarg>

namespace carca {

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;
constexpr float kLeakySlope = 0.01f;   // nn.LeakyReLU default, src/carca.py:285
constexpr float kLnEps = 1e-5f;        // nn.LayerNorm default, src/carca.py:279,283,408

// ---- error reporting: entry points return 0 or a negative code; text via carca_last_error()
char* err_buf();
int fail(int code, const char* fmt, ...);
int check_launch(const char* what);   // call once after every kernel launch (also counts it)
const unsigned long long* seed_source();   // device word set by carca_set_seed_source (or null)
long long launch_count();

#define CARCA_REQUIRE(cond, ...)                      \
  do {                                                \
    if (!(cond)) return carca::fail(-2, __VA_ARGS__); \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(kFull, v, o));
  return v;
}

// ---- Philox4x32-10 dropout stream (restated in oracle/philox.py; keep them in lock-step)
struct Philox4 { unsigned x, y, z, w; };

__host__ __device__ __forceinline__ Philox4 philox4x32_10(unsigned c0, unsigned c1, unsigned c2, unsigned c3,
                                                          unsigned k0, unsigned k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = 0xD2511F53ull * c0;
    const unsigned long long p1 = 0xCD9E8D57ull * c2;
    const unsigned hi0 = (unsigned)(p0 >> 32), lo0 = (unsigned)p0;
    const unsigned hi1 = (unsigned)(p1 >> 32), lo1 = (unsigned)p1;
    const unsigned n0 = hi1 ^ c1 ^ k0;
    const unsigned n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return Philox4{c0, c1, c2, c3};
}

struct DropCfg {
  float p;                 // drop probability; 0 disables
  float scale;             // 1/(1-p)
  unsigned seed_lo, seed_hi;
  unsigned site;
  const unsigned long long* seed_dev;   // optional device word XOR-ed into the seed when the kernel runs, so a
                                        // captured CUDA graph draws a fresh mask on every replay (carca_set_seed_source)
};

__host__ __device__ __forceinline__ DropCfg make_drop(float p, unsigned long long seed, unsigned site) {
  DropCfg c;
  c.p = p;
  c.scale = p > 0.f ? 1.0f / (1.0f - p) : 1.0f;
  c.seed_lo = (unsigned)seed;
  c.seed_hi = (unsigned)(seed >> 32);
  c.site = site;
  c.seed_dev = nullptr;
  return c;
}

// keep-scale for flat element index `elem` of the site's tensor: 0 (dropped) or 1/(1-p)
__host__ __device__ __forceinline__ float drop_factor(const DropCfg& c, unsigned long long elem) {
  if (c.p <= 0.f) return 1.0f;
  unsigned k0 = c.seed_lo, k1 = c.seed_hi;
  if (c.seed_dev) {
    const unsigned long long x = *c.seed_dev;
    k0 ^= (unsigned)x;
    k1 ^= (unsigned)(x >> 32);
  }
  const unsigned long long ctr = elem >> 2;
  const Philox4 r = philox4x32_10((unsigned)ctr, (unsigned)(ctr >> 32), c.site, 0u, k0, k1);
  const unsigned lane = (unsigned)(elem & 3ull);
  const unsigned word = lane == 0 ? r.x : lane == 1 ? r.y : lane == 2 ? r.z : r.w;
  const float u = (float)(word >> 8) * 5.9604644775390625e-08f;  // 2^-24
  return u >= c.p ? c.scale : 0.0f;
}

__host__ __device__ __forceinline__ int ceil_div(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long ceil_div_ll(long long a, long long b) { return (a + b - 1) / b; }

}  // namespace carca
