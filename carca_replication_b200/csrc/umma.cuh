// tcgen05 (5th-generation tensor core) primitives for sm_100a, written as inline PTX:
// TMEM allocation, shared-memory matrix descriptors, the tf32 MMA issue, commit -> mbarrier,
// bounded mbarrier waits and TMEM -> register loads.  Not available under the CPU emulator.
//
// Operand layout used throughout (K-major, no swizzle, 32-bit elements): a [rows x K] operand is
// stored as [K/4][rows][4] floats, i.e. 16-byte chunks of 4 consecutive k for one row, rows
// contiguous.  In the descriptor's terms a core matrix is 8 rows x 16 B = 128 contiguous bytes,
// SBO (next 8 rows) = 128 B, LBO (next 16-byte k chunk) = rows * 16 B.  One tf32 MMA consumes
// K = 8 (two chunks), so successive instructions advance the start address by 2 * LBO.
//
// fp32-grade accuracy from the tf32 tensor cores ("3xTF32"): x = hi + lo with hi = x truncated to
// tf32 (what the hardware reads from the fp32 word) and lo = x - hi (exact), and
//   A B^T ~= hi(A) hi(B)^T + lo(A) hi(B)^T + hi(A) lo(B)^T     (error ~2^-21 per product).
#pragma once
#ifndef CARCA_EMU
#include <cuda_runtime.h>
#include <cstdint>

namespace carca {
namespace umma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ float tf32_hi(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
__device__ __forceinline__ float tf32_lo(float x) { return x - tf32_hi(x); }

// one lane of a converged warp (the branch taken around it must be warp-uniform)
__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xFFFFFFFFu));
  return pred != 0;
}

// ---- TMEM allocation (one full warp), address written to a shared slot
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, int columns) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)),
               "r"(columns)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_free(uint32_t addr, int columns) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(columns) : "memory");
}

__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// generic-proxy shared-memory writes -> visible to the async proxy (tensor core operand fetch)
__device__ __forceinline__ void fence_smem_to_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%1], %0;" ::"r"(count), "r"(smem_u32(bar)) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ bool mbar_try(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P1;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
      "selp.b32 %0, 1, 0, P1;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: returns false (never hangs) if the phase does not complete.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int max_spins = 4000000) {
#pragma unroll 1
  for (int i = 0; i < max_spins; ++i)
    if (mbar_try(bar, parity)) return true;
  return false;
}
// all previously issued MMAs of this thread arrive on the mbarrier when they complete
__device__ __forceinline__ void commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// ---- descriptors
// K-major, no swizzle: lbo/sbo in bytes (see header comment)
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;
  d |= 1ull << 46;   // descriptor version for sm_100
  return d;          // base_offset 0, lbo_mode 0, layout_type 0 (no swizzle)
}
// tf32 x tf32 -> f32, both operands K-major, M = 128
__host__ __device__ constexpr uint32_t idesc_tf32(int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T, one K = 8 step
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                         bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// Full product over K (multiple of 8) of operands stored [K/4][rows][4]; issued by ONE thread.
//   a_rows = 128, b_rows = N.  `first` clears the accumulator on the first step.
__device__ __forceinline__ void mma_tf32_k(uint32_t d_tmem, uint32_t a_saddr, uint32_t b_saddr, int N, int K,
                                           bool first) {
  const uint32_t idesc = idesc_tf32(N);
  const uint32_t a_lbo = 128 * 16, b_lbo = (uint32_t)N * 16;
  for (int ks = 0; ks < K / 8; ++ks) {
    const uint64_t da = smem_desc(a_saddr + ks * 2 * a_lbo, a_lbo, 128);
    const uint64_t db = smem_desc(b_saddr + ks * 2 * b_lbo, b_lbo, 128);
    mma_tf32(d_tmem, da, db, idesc, !(first && ks == 0));
  }
}

// Same with the A operand read from TMEM (lane = row, one 32-bit column per k element)
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                            bool accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"((uint32_t)accumulate)
      : "memory");
}

// B operand stored MN-major, no swizzle: [K/8][N/4][8 k][4 n] floats; one K = 8 step is one k group.
//   sbo = next 16-byte n chunk (128 B), lbo = next k group (N_total * 32 B)
__host__ __device__ __forceinline__ uint32_t idesc_tf32_bmn(int n) { return idesc_tf32(n) | (1u << 16); }

// ---- registers -> TMEM: this warp's 32 lanes, 8 consecutive columns
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const float (&v)[8]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
               :
               : "r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])),
                 "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])),
                 "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---- TMEM -> registers: this warp's 32 lanes (rows 32*(warp%4) ..), 8 consecutive columns
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
  // the registers are only valid after the wait: tie them to it so nothing is hoisted above
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7])
               :
               : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

}  // namespace umma
}  // namespace carca
#endif  // CARCA_EMU
