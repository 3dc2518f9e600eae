// TMEM <-> register moves wider than umma.cuh's 8-column primitives (32x32b shape: lane = row,
// one 32-bit register per column).  Each load helper is ONE asm statement that ends in
// tcgen05.wait::ld, so the destination registers are valid when it returns and several chunks
// share a single wait.  Stores do not wait: call umma::tmem_st_wait() (publish) before the data is
// consumed by an MMA or read back.
#pragma once
#ifndef CARCA_EMU
#include <cstdint>

namespace carca {
namespace umma {

// 1 chunk(s) of 8 consecutive columns -> v[0..8)
__device__ __forceinline__ void tmem_ld_1x8(uint32_t a0, float (&v)[8]) {
  uint32_t r[8];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7])
      : "r"(a0)
      : "memory");
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}

// 1 chunk(s) of 16 consecutive columns -> v[0..16)
__device__ __forceinline__ void tmem_ld_1x16(uint32_t a0, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
      : "r"(a0)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 1 chunk(s) of 32 consecutive columns -> v[0..32)
__device__ __forceinline__ void tmem_ld_1x32(uint32_t a0, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15]), "=&r"(r[16]), "=&r"(r[17]), "=&r"(r[18]), "=&r"(r[19]), "=&r"(r[20]), "=&r"(r[21]), "=&r"(r[22]), "=&r"(r[23]), "=&r"(r[24]), "=&r"(r[25]), "=&r"(r[26]), "=&r"(r[27]), "=&r"(r[28]), "=&r"(r[29]), "=&r"(r[30]), "=&r"(r[31])
      : "r"(a0)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 2 chunk(s) of 16 consecutive columns -> v[0..32)
__device__ __forceinline__ void tmem_ld_2x16(uint32_t a0, uint32_t a1, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%32];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%33];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15]), "=&r"(r[16]), "=&r"(r[17]), "=&r"(r[18]), "=&r"(r[19]), "=&r"(r[20]), "=&r"(r[21]), "=&r"(r[22]), "=&r"(r[23]), "=&r"(r[24]), "=&r"(r[25]), "=&r"(r[26]), "=&r"(r[27]), "=&r"(r[28]), "=&r"(r[29]), "=&r"(r[30]), "=&r"(r[31])
      : "r"(a0), "r"(a1)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 4 chunk(s) of 8 consecutive columns -> v[0..32)
__device__ __forceinline__ void tmem_ld_4x8(uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%32];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%33];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%16, %17, %18, %19, %20, %21, %22, %23}, [%34];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%24, %25, %26, %27, %28, %29, %30, %31}, [%35];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15]), "=&r"(r[16]), "=&r"(r[17]), "=&r"(r[18]), "=&r"(r[19]), "=&r"(r[20]), "=&r"(r[21]), "=&r"(r[22]), "=&r"(r[23]), "=&r"(r[24]), "=&r"(r[25]), "=&r"(r[26]), "=&r"(r[27]), "=&r"(r[28]), "=&r"(r[29]), "=&r"(r[30]), "=&r"(r[31])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3)
      : "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 2 chunk(s) of 8 consecutive columns -> v[0..16)
__device__ __forceinline__ void tmem_ld_2x8(uint32_t a0, uint32_t a1, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%16];\n\t"
      "tcgen05.ld.sync.aligned.32x32b.x8.b32 {%8, %9, %10, %11, %12, %13, %14, %15}, [%17];\n\t"
      "tcgen05.wait::ld.sync.aligned;"
      : "=&r"(r[0]), "=&r"(r[1]), "=&r"(r[2]), "=&r"(r[3]), "=&r"(r[4]), "=&r"(r[5]), "=&r"(r[6]), "=&r"(r[7]), "=&r"(r[8]), "=&r"(r[9]), "=&r"(r[10]), "=&r"(r[11]), "=&r"(r[12]), "=&r"(r[13]), "=&r"(r[14]), "=&r"(r[15])
      : "r"(a0), "r"(a1)
      : "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 8 consecutive columns <- v[off .. off+8)
template <int NV>
__device__ __forceinline__ void tmem_st_x8(uint32_t a, const float (&vv)[NV], int off) {
  const float* v = vv + off;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%8], {%0, %1, %2, %3, %4, %5, %6, %7};"
               :
               : "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(a)
               : "memory");
}

// 16 consecutive columns <- v[off .. off+16)
template <int NV>
__device__ __forceinline__ void tmem_st_x16(uint32_t a, const float (&vv)[NV], int off) {
  const float* v = vv + off;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%16], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15};"
               :
               : "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(a)
               : "memory");
}

// 32 consecutive columns <- v[off .. off+32)
template <int NV>
__device__ __forceinline__ void tmem_st_x32(uint32_t a, const float (&vv)[NV], int off) {
  const float* v = vv + off;
  asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%32], {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31};"
               :
               : "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])), "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])), "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])), "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])), "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])), "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])), "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])), "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31])), "r"(a)
               : "memory");
}

}  // namespace umma
}  // namespace carca
#endif  // CARCA_EMU
