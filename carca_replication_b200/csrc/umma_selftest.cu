// Known-answer self test of the tcgen05 primitives in umma.cuh:
//   C[128, N] = A[128, K] B[N, K]^T   with tf32 tensor-core MMAs (optionally the 3xTF32 split),
// operands staged into shared memory in the K-major no-swizzle layout, accumulator in TMEM.
// Every wait is bounded: on a timeout the kernel reports status 1 instead of hanging.
#include "../../include/carca_b200.h"
#include "common.cuh"
#include "umma.cuh"

namespace carca {

#ifndef CARCA_EMU
__global__ void __launch_bounds__(128, 1) umma_selftest_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                               float* __restrict__ C, int N, int K, int split,
                                                               int* __restrict__ status) {
  CARCA_DYN_SMEM(float, sm);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  float* a_hi = sm;                       // [K/4][128][4]
  float* a_lo = a_hi + 128 * K;
  float* b_hi = a_lo + 128 * K;           // [K/4][N][4]
  float* b_lo = b_hi + N * K;
  const int tid = threadIdx.x, warp = tid / 32;
  int cols = 32;
  while (cols < N) cols *= 2;

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
    const int idx = ((k / 4) * N + n) * 4 + (k % 4);
    b_hi[idx] = x;
    b_lo[idx] = umma::tf32_lo(x);
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem = tmem_slot;

  if (tid == 0) {
    umma::mma_tf32_k(tmem, umma::smem_u32(a_hi), umma::smem_u32(b_hi), N, K, true);
    if (split) {
      umma::mma_tf32_k(tmem, umma::smem_u32(a_lo), umma::smem_u32(b_hi), N, K, false);
      umma::mma_tf32_k(tmem, umma::smem_u32(a_hi), umma::smem_u32(b_lo), N, K, false);
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
      umma::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
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

extern "C" int carca_umma_selftest(float* C, const float* A, const float* B, int N, int K, int split, int32_t* status,
                                   void* stream) {
#ifndef CARCA_EMU
  using namespace carca;
  if (N % 16 != 0 || N < 16 || N > 256 || K % 8 != 0 || K < 8) return fail(-2, "umma_selftest: N%%16, K%%8 required");
  const size_t smem = sizeof(float) * (size_t)(2 * 128 * K + 2 * N * K);
  if (smem > 200 * 1024) return fail(-2, "umma_selftest: operands do not fit in shared memory");
  auto k = umma_selftest_kernel;
  if (smem > 48 * 1024) cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaMemsetAsync(status, 0, sizeof(int), reinterpret_cast<cudaStream_t>(stream));
  k<<<1, 128, smem, reinterpret_cast<cudaStream_t>(stream)>>>(A, B, C, N, K, split, status);
  return check_launch("umma_selftest");
#else
  (void)C; (void)A; (void)B; (void)N; (void)K; (void)split; (void)status; (void)stream;
  return carca::fail(-5, "umma_selftest: tcgen05 is not available under the CPU emulator");
#endif
}
