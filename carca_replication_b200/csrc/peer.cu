// All-reduce (sum) of a flat fp32 buffer over the GPUs of ONE NVSwitch box, as a single kernel on peer-mapped memory:
// the gradient exchange of the data-parallel training step (the reference is single-device and has no
// counterpart).  Declared in include/carca_b200.h.
//
// Every rank owns a communication buffer (cudaMalloc + cudaIpcGetMemHandle, opened by all peers): flags + n floats.
// The kernel runs with the SAME grid on every rank; block b of every rank works on sub-chunk b of each of the W slices
// of the buffer, so that a per-block handshake between equal-numbered blocks is all the synchronisation it needs:
//
//   copy-in    block b copies sub-chunk b of every slice from the local tensor into its own communication buffer
//   barrier A  block b tells block b of every peer "my sub-chunks are in place" (flag = epoch, release.sys over NVLink)
//   reduce     rank r owns slice r: block b reads sub-chunk b of slice r from ALL ranks (peer loads, fixed rank order
//              0..W-1: every rank ends up with bit-identical sums) and writes the sum into EVERY rank's buffer (peer stores)
//   barrier B  "my sums are written everywhere"
//   copy-out   block b copies sub-chunk b of every slice back into the local tensor
// When the tensor lives IN the communication buffer (the data-parallel wrapper hands the fused training step its flat
// gradient buffer from there) both copies disappear and the exchange is barrier, reduce + push, barrier.
//
// NVLink traffic per GPU: (W-1)/W of the buffer in, the same out — the two-shot all-reduce minimum — with no staging
// copies between phases and no host involvement; the whole exchange is one graph node.  Waits are bounded in time: a
// peer that has not arrived after ~20 s sets bit 8 of the status word instead of hanging the GPU.
#include "../../include/carca_b200.h"

#include "common.cuh"

using namespace carca;

#ifdef CARCA_EMU
extern "C" {
int64_t carca_peer_buffer_bytes(int64_t) { return 0; }
int64_t carca_peer_data_offset(void) { return 0; }
int carca_peer_alloc(void**, int64_t) { return fail(-5, "peer all-reduce: not available under the CPU emulator"); }
int carca_peer_free(void*) { return 0; }
int carca_peer_export(void*, void*) { return fail(-5, "peer all-reduce: not available under the CPU emulator"); }
int carca_peer_open(const void*, void**) { return fail(-5, "peer all-reduce: not available under the CPU emulator"); }
int carca_peer_close(void*) { return 0; }
int carca_peer_allreduce(float*, int64_t, void* const*, int, int, int32_t*, void*) {
  return fail(-5, "peer all-reduce: not available under the CPU emulator");
}
}
#else
namespace {

constexpr int PEER_MAX_WORLD = 8;
constexpr int PEER_MAX_BLOCKS = 128;
constexpr int PEER_THREADS = 512;
// layout of a communication buffer
constexpr long long PEER_OFF_EPOCH = 0;      // int epoch, int ticket
constexpr long long PEER_OFF_FLAGS = 4096;   // int [2][PEER_MAX_BLOCKS][PEER_MAX_WORLD]
constexpr long long PEER_OFF_DATA = 16384;

struct PeerArgs {
  unsigned char* base[PEER_MAX_WORLD];
  float* local;
  long long n4;          // float4 elements (n padded up inside the buffer; the tail of the last one is handled apart)
  long long n;
  int rank, world;
  int* status;
};

__device__ __forceinline__ void st_release_sys(int* p, int v) {
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ int ld_acquire_sys(const int* p) {
  int v;
  asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// reads that must see what a PEER wrote into this memory during the kernel: no L1
__device__ __forceinline__ float4 ld_cv(const float4* p) {
  float4 v;
  asm volatile("ld.volatile.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}

// block-level handshake with the same block of every peer (phase 0 / 1)
__device__ __forceinline__ void peer_barrier(const PeerArgs& a, int phase, int epoch) {
  // the CTA barrier orders every thread's writes before the releasing store below, and the release is cumulative at
  // system scope: no per-thread __threadfence_system() (512 full fences per block cost more than the exchange itself)
  __syncthreads();
  const int t = threadIdx.x;
  if (t < a.world && t != a.rank) {
    int* theirs = reinterpret_cast<int*>(a.base[t] + PEER_OFF_FLAGS) + ((long long)phase * PEER_MAX_BLOCKS + blockIdx.x) * PEER_MAX_WORLD + a.rank;
    st_release_sys(theirs, epoch);
    const int* mine = reinterpret_cast<const int*>(a.base[a.rank] + PEER_OFF_FLAGS) + ((long long)phase * PEER_MAX_BLOCKS + blockIdx.x) * PEER_MAX_WORLD + t;
    const long long t0 = clock64();
    while (ld_acquire_sys(mine) - epoch < 0) {
      if (clock64() - t0 > 40000000000ll) {   // ~20 s: a rank may be far behind (first step), a dead one must not hang us
        atomicOr(a.status, 8);
        break;
      }
    }
  }
  __syncthreads();
}

template <int W>
__global__ void __launch_bounds__(PEER_THREADS) peer_allreduce_kernel(const PeerArgs a) {
  __shared__ int s_epoch;
  int* ctl = reinterpret_cast<int*>(a.base[a.rank] + PEER_OFF_EPOCH);
  if (threadIdx.x == 0) s_epoch = ctl[0] + 1;
  __syncthreads();
  const int epoch = s_epoch;
  const int world = W > 0 ? W : a.world;
  const long long S4 = (a.n4 + world - 1) / world;                 // float4 per slice
  const long long C4 = (S4 + gridDim.x - 1) / gridDim.x;           // float4 per sub-chunk
  float4* mine = reinterpret_cast<float4*>(a.base[a.rank] + PEER_OFF_DATA);
  const float4* loc = reinterpret_cast<const float4*>(a.local);
  const long long full4 = a.n >> 2;                                // whole float4 of the local tensor

  // `local` IS this rank's communication buffer (PeerAllReduce.buffer()): nothing to copy in or out
  const bool in_place = reinterpret_cast<const unsigned char*>(a.local) == a.base[a.rank] + PEER_OFF_DATA;

  // ---- copy-in
  for (int s = 0; s < world && !in_place; ++s) {
    const long long lo = s * S4 + blockIdx.x * C4, hi = min(min(lo + C4, (s + 1) * S4), a.n4);
    for (long long i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
      float4 v;
      if (i < full4) {
        v = loc[i];
      } else {                                                     // the tensor's last, partial float4
        const float* l = a.local + 4 * i;
        const long long r = a.n - 4 * i;
        v = make_float4(r > 0 ? l[0] : 0.f, r > 1 ? l[1] : 0.f, r > 2 ? l[2] : 0.f, 0.f);
      }
      mine[i] = v;
    }
  }
  peer_barrier(a, 0, epoch);

  // ---- reduce slice `rank` over all ranks, write the sums to all ranks.  U x W peer loads of 16 bytes are in flight
  // per thread before the first add (an NVLink round trip is microseconds: the loop is latency-, not issue-bound)
  {
    constexpr int U = W == 2 ? 8 : (W == 4 ? 4 : 2);
    const long long lo = a.rank * S4 + blockIdx.x * C4, hi = min(min(lo + C4, (a.rank + 1) * S4), a.n4);
    for (long long i0 = lo; i0 < hi; i0 += U * PEER_THREADS) {
      float4 v[U][PEER_MAX_WORLD];
#pragma unroll
      for (int p = 0; p < PEER_MAX_WORLD; ++p) {
        if (p < world) {
          const float4* src = reinterpret_cast<const float4*>(a.base[p] + PEER_OFF_DATA);
#pragma unroll
          for (int q = 0; q < U; ++q) {
            const long long k = i0 + threadIdx.x + (long long)q * PEER_THREADS;
            if (k < hi) v[q][p] = ld_cv(src + k);
          }
        }
      }
#pragma unroll
      for (int q = 0; q < U; ++q) {
        const long long k = i0 + threadIdx.x + (long long)q * PEER_THREADS;
        if (k < hi) {
          float4 acc = v[q][0];
#pragma unroll
          for (int p = 1; p < PEER_MAX_WORLD; ++p)
            if (p < world) { acc.x += v[q][p].x; acc.y += v[q][p].y; acc.z += v[q][p].z; acc.w += v[q][p].w; }
#pragma unroll
          for (int p = 0; p < PEER_MAX_WORLD; ++p)
            if (p < world) reinterpret_cast<float4*>(a.base[p] + PEER_OFF_DATA)[k] = acc;
        }
      }
    }
  }
  peer_barrier(a, 1, epoch);

  // ---- copy-out
  float4* out = reinterpret_cast<float4*>(a.local);
  for (int s = 0; s < world && !in_place; ++s) {
    const long long lo = s * S4 + blockIdx.x * C4, hi = min(min(lo + C4, (s + 1) * S4), a.n4);
    for (long long i = lo + threadIdx.x; i < hi; i += PEER_THREADS) {
      const float4 v = ld_cv(mine + i);
      if (i < full4) {
        out[i] = v;
      } else {
        float* l = a.local + 4 * i;
        const long long r = a.n - 4 * i;
        if (r > 0) l[0] = v.x;
        if (r > 1) l[1] = v.y;
        if (r > 2) l[2] = v.z;
      }
    }
  }
  // the last block to finish publishes the epoch for the next call (a graph replay reads it from the device)
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(&ctl[1], 1) == (int)gridDim.x - 1) {
      ctl[1] = 0;
      ctl[0] = epoch;
      __threadfence();
    }
  }
}

}  // namespace

extern "C" {

int64_t carca_peer_data_offset(void) { return PEER_OFF_DATA; }

int64_t carca_peer_buffer_bytes(int64_t n_floats) { return PEER_OFF_DATA + ((n_floats + 3) / 4) * 16 + 256; }

int carca_peer_alloc(void** base, int64_t bytes) {
  CARCA_REQUIRE(base != nullptr && bytes >= PEER_OFF_DATA, "peer_alloc: bad arguments");
  cudaError_t e = cudaMalloc(base, (size_t)bytes);
  if (e != cudaSuccess) return fail(-3, "peer_alloc: cudaMalloc(%lld): %s", (long long)bytes, cudaGetErrorString(e));
  e = cudaMemset(*base, 0, (size_t)bytes);
  if (e != cudaSuccess) return fail(-3, "peer_alloc: cudaMemset: %s", cudaGetErrorString(e));
  return 0;
}

int carca_peer_free(void* base) {
  if (base) cudaFree(base);
  return 0;
}

int carca_peer_export(void* base, void* handle64) {
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
  cudaError_t e = cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), base);
  if (e != cudaSuccess) return fail(-3, "peer_export: cudaIpcGetMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

int carca_peer_open(const void* handle64, void** base) {
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  cudaError_t e = cudaIpcOpenMemHandle(base, h, cudaIpcMemLazyEnablePeerAccess);
  if (e != cudaSuccess) return fail(-3, "peer_open: cudaIpcOpenMemHandle: %s", cudaGetErrorString(e));
  return 0;
}

int carca_peer_close(void* base) {
  if (base) cudaIpcCloseMemHandle(base);
  return 0;
}

int carca_peer_allreduce(float* local, int64_t n, void* const* bases, int rank, int world, int32_t* status, void* stream) {
  CARCA_REQUIRE(world >= 1 && world <= PEER_MAX_WORLD && rank >= 0 && rank < world, "peer_allreduce: rank %d of %d", rank, world);
  CARCA_REQUIRE(local != nullptr && bases != nullptr && status != nullptr && n >= 0, "peer_allreduce: null argument");
  CARCA_REQUIRE((reinterpret_cast<uintptr_t>(local) & 15) == 0, "peer_allreduce: the tensor must be 16-byte aligned");
  if (n == 0 || world == 1) return 0;
  PeerArgs a;
  memset(&a, 0, sizeof(a));
  for (int p = 0; p < world; ++p) {
    CARCA_REQUIRE(bases[p] != nullptr, "peer_allreduce: buffer of rank %d is not mapped", p);
    a.base[p] = reinterpret_cast<unsigned char*>(bases[p]);
  }
  a.local = local; a.n = n; a.n4 = (n + 3) / 4; a.rank = rank; a.world = world; a.status = status;
  const long long S4 = (a.n4 + world - 1) / world;
  const int grid = (int)max(1ll, min((long long)PEER_MAX_BLOCKS, ceil_div_ll(S4, 2 * PEER_THREADS)));   // same on every rank
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (world == 2) CARCA_LAUNCH(peer_allreduce_kernel<2>, dim3(grid), dim3(PEER_THREADS), 0, st, a);
  else if (world == 4) CARCA_LAUNCH(peer_allreduce_kernel<4>, dim3(grid), dim3(PEER_THREADS), 0, st, a);
  else if (world == 8) CARCA_LAUNCH(peer_allreduce_kernel<8>, dim3(grid), dim3(PEER_THREADS), 0, st, a);
  else CARCA_LAUNCH(peer_allreduce_kernel<0>, dim3(grid), dim3(PEER_THREADS), 0, st, a);
  return check_launch("peer_allreduce");
}

}  // extern "C"
#endif
