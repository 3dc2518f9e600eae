// Known-answer self test of the tcgen05 primitives in umma.cuh:
//   C[128, N] = A[128, K] B[N, K]^T   on the tf32 tensor cores, accumulator in TMEM.
// mode bits: 1 = 3xTF32 split (fp32 grade), 2 = A operand read from TMEM (written with tcgen05.st)
// instead of shared memory, 4 = B operand stored MN-major instead of K-major.
// Every wait is bounded: on a timeout the kernel reports status 1 instead of hanging.
#include "../../include/carca_b200.h"
#include "common.cuh"
#include "umma.cuh"

namespace carca {

#ifndef CARCA_EMU
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ C, int N, int K, int mode,
                                                               int* __restrict__ status) {
  CARCA_DYN_SMEM(float, sm);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const bool split = mode & 1, a_tmem = mode & 2, b_mn = mode & 4;
  float* a_hi = sm;                       // [K/4][128][4]
  float* a_lo = a_hi + 128 * K;
  float* b_hi = a_lo + 128 * K;           // K-major [K/4][N][4]  or  MN-major [K/8][N/4][8][4]
  float* b_lo = b_hi + N * K;
  const int tid = threadIdx.x, warp = tid / 32;
  const int cols = 512;                   // accumulator at column 0, TMEM A operand at 256 (hi) / 384 (lo)

  if (warp == 0) umma::tmem_alloc(&tmem_slot, cols);
  if (tid == 0) umma::mbar_init(&bar, 1);
  for (int e = tid; e < 128 * K; e += 128) {
    const int m = e / K, k = e % K;
    const float x = A[e];
    const int idx = ((k / 4) * 128 + m) * 4 + (k % 4);
    a_hi[idx] = x;
    a_lo[idx] = umma::tf32_lo(x);
  }
  for (int e = tid; e < N * K; e += 128) {
    const int n = e / K, k = e % K;
    const float x = B[e];
    const int idx = b_mn ? ((((k / 8) * (N / 4) + n / 4) * 8 + (k % 8)) * 4 + (n % 4))
                         : (((k / 4) * N + n) * 4 + (k % 4));
    b_hi[idx] = x;
    b_lo[idx] = umma::tf32_lo(x);
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  const uint32_t lane_base = (uint32_t)(warp * 32) << 16;

  if (a_tmem) {   // each thread owns row `tid`: write it (and its tf32 remainder) into TMEM columns
    for (int k0 = 0; k0 < K; k0 += 8) {
      float hi[8], lo[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        hi[j] = A[tid * K + k0 + j];
        lo[j] = umma::tf32_lo(hi[j]);
      }
      umma::tmem_st8(tmem + lane_base + 256 + k0, hi);
      umma::tmem_st8(tmem + lane_base + 384 + k0, lo);
    }
    umma::tmem_st_wait();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
  }

  if (tid == 0) {
    const uint32_t idesc = b_mn ? umma::idesc_tf32_bmn(N) : umma::idesc_tf32(N);
    const uint32_t a_lbo = 128 * 16;
    const uint32_t b_lbo = b_mn ? (uint32_t)N * 32 : (uint32_t)N * 16;
    const uint32_t b_step = b_mn ? b_lbo : 2 * b_lbo;      // bytes per K = 8 step
    const int passes = split ? 3 : 1;
    for (int p = 0; p < passes; ++p) {
      const bool a_is_lo = (p == 1), b_is_lo = (p == 2);
      const uint32_t a_s = umma::smem_u32(a_is_lo ? a_lo : a_hi), b_s = umma::smem_u32(b_is_lo ? b_lo : b_hi);
      for (int ks = 0; ks < K / 8; ++ks) {
        const uint64_t db = umma::smem_desc(b_s + ks * b_step, b_lbo, 128);
        const bool acc = !(p == 0 && ks == 0);
        if (a_tmem) {
          umma::mma_tf32_ts(tmem, tmem + (a_is_lo ? 384 : 256) + ks * 8, db, idesc, acc);
        } else {
          const uint64_t da = umma::smem_desc(a_s + ks * 2 * a_lbo, a_lbo, 128);
          umma::mma_tf32(tmem, da, db, idesc, acc);
        }
      }
    }
    umma::commit(&bar);
  }
  const bool ok = umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    const int row = warp * 32 + (tid % 32);
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tmem + lane_base + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) C[row * N + c0 + j] = v[j];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, cols);
}
#endif

}  // namespace carca

extern "C" int carca_umma_selftest(float* C, const float* A, const float* B, int N, int K, int mode, int32_t* status,
                                   void* stream) {
#ifndef CARCA_EMU
  using namespace carca;
  if (N % 16 != 0 || N < 16 || N > 256 || K % 8 != 0 || K < 8) return fail(-2, "umma_selftest: N%%16, K%%8 required");
  if ((mode & 2) && K > 128) return fail(-2, "umma_selftest: the TMEM A operand test takes K <= 128");
  const size_t smem = sizeof(float) * (size_t)(2 * 128 * K + 2 * N * K);
  if (smem > 200 * 1024) return fail(-2, "umma_selftest: operands do not fit in shared memory");
  auto k = umma_selftest_kernel;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(status, 0, sizeof(int), reinterpret_cast<cudaStream_t>(stream));
  k<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, C, N, K, mode, status);
  return check_launch("umma_selftest");
#else
  (void)C; (void)A; (void)B; (void)N; (void)K; (void)mode; (void)status; (void)stream;
  return carca::fail(-5, "umma_selftest: tcgen05 is not available under the CPU emulator");
#endif
}

// ---------------------------------------------------------------------------------------------
// Layout probe: the host supplies raw shared-memory images of both operands plus every descriptor
// field, so operand layouts can be explored from Python without recompiling.
namespace carca {
#ifndef CARCA_EMU
__global__ void __launch_bounds__(128, 1) umma_probe_kernel(float* __restrict__ C, const float* __restrict__ a_img,
                                                            int a_floats, const float* __restrict__ b_img, int b_floats,
                                                            int N, int ksteps, uint32_t a_lbo, uint32_t a_sbo,
                                                            uint32_t a_step, uint32_t b_lbo, uint32_t b_sbo,
                                                            uint32_t b_step, uint32_t idesc, int* __restrict__ status) {
  CARCA_DYN_SMEM(float, sm);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* a_s = sm;
  float* b_s = sm + ((a_floats + 255) / 256) * 256;
  const int tid = threadIdx.x, warp = tid / 32;
  if (warp == 0) umma::tmem_alloc(&tmem_slot, 256);
  if (tid == 0) umma::mbar_init(&bar, 1);
  for (int e = tid; e < a_floats; e += 128) a_s[e] = a_img[e];
  for (int e = tid; e < b_floats; e += 128) b_s[e] = b_img[e];
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;
  if (tid == 0) {
    for (int ks = 0; ks < ksteps; ++ks)
      umma::mma_tf32(tmem, umma::smem_desc(umma::smem_u32(a_s) + ks * a_step, a_lbo, a_sbo),
                     umma::smem_desc(umma::smem_u32(b_s) + ks * b_step, b_lbo, b_sbo), idesc, ks > 0);
    umma::commit(&bar);
  }
  const bool ok = umma::mbar_wait(&bar, 0);
  umma::fence_after_sync();
  if (!ok) {
    if (tid == 0) status[0] = 1;
  } else {
    for (int c0 = 0; c0 < N; c0 += 8) {
      float v[8];
      umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) C[tid * N + c0 + j] = v[j];
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 0) umma::tmem_free(tmem, 256);
}
#endif
}  // namespace carca

extern "C" int carca_umma_probe(float* C, const float* a_img, int a_floats, const float* b_img, int b_floats, int N,
                                int ksteps, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_lbo,
                                uint32_t b_sbo, uint32_t b_step, uint32_t idesc, int32_t* status, void* stream) {
#ifndef CARCA_EMU
  using namespace carca;
  const size_t smem = sizeof(float) * (size_t)(((a_floats + 255) / 256) * 256 + b_floats);
  if (smem > 200 * 1024) return fail(-2, "umma_probe: images do not fit in shared memory");
  auto k = umma_probe_kernel;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(status, 0, sizeof(int), reinterpret_cast<cudaStream_t>(stream));
  k<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(C, a_img, a_floats, b_img, b_floats, N, ksteps, a_lbo, a_sbo,
                                                              a_step, b_lbo, b_sbo, b_step, idesc, status);
  return check_launch("umma_probe");
#else
  return carca::fail(-5, "umma_probe: tcgen05 is not available under the CPU emulator");
#endif
}
