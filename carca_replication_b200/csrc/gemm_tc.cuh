// tcgen05 (tensor-core) version of the general GEMM of gemm.cuh — same GemmArgs, same epilogue, same
// dropout element indexing, so it is a drop-in for every projection / weight-gradient product of the
// training path (nn.Linear / Conv1d(k=1) forward and autograd, src/carca.py:86,89,238-240,307,311):
//
//   C[m,n] (=|+=)  rowmask[m] * ( drop( act( alpha * sum_k A'(m,k) B'(k,n) + bias[n] (+ C_old) ) ) + R[m % r_mod, n] )
//
// One CTA computes a 128 x BN tile (BN = 64 / 128 / 256, so that N <= 256 reads A once).  Per 32-wide k
// step both operands go global -> registers -> shared memory in the K-major no-swizzle layout of
// umma.cuh ([k/4][row][4] floats, chunk stride padded by 16 bytes so coalesced loads store without
// bank conflicts) as a tf32 pair (hi = the fp32 value, lo = its remainder); one elected lane issues
// hi*hi + lo*hi + hi*lo (3xTF32, fp32-grade) tcgen05.mma with the accumulator in tensor memory; the
// global loads of step t+1 are in flight while the MMAs of step t run, and 2-3 CTAs per SM overlap
// each other's staging and epilogue.  All four storage cases are handled at staging time:
//   rows contiguous along k  (A: !transA with optional row gather, B: transB)      -> coalesced 128-bit loads
//   rows contiguous along m/n (A: transA, B: !transB with optional k-row gather)   -> transposing scalar stores
// Epilogue: TMEM -> registers -> XOR-swizzled shared tile -> coalesced 128-bit global accesses.
// Split-K (weight gradients: tiny output, reduction over every position) accumulates with atomics.
#pragma once
#include "gemm.cuh"
#include "tmem_io.cuh"
#include "umma.cuh"

#ifndef CARCA_EMU
namespace carca {

constexpr int GT_BM = 128, GT_BK = 32, GT_THREADS = 256;
constexpr int gt_lbo(int rows) { return rows * 16 + 16; }   // bytes between 4-k chunks (padded)

template <int BN>
struct GemmTcSmem {
  // operands of one k step: [hi | lo][8 chunks]; the same bytes hold the C tile during the epilogue
  float a[2][8 * gt_lbo(GT_BM) / 4];
  float b[2][8 * gt_lbo(BN) / 4];
  uint64_t bar;
  uint32_t tmem_slot;
};

// This thread's quads of a [ROWS x 32] operand tile.  kmode: element (row, k) at P[G(row) * ld + k]
// (contiguous along k, G = optional gather over rows); else at P[G(k) * ld + row] (contiguous along rows,
// G = optional gather over k).
template <int ROWS>
__device__ __forceinline__ void fetch_tile(float4 (&r)[ROWS / 32], const float* __restrict__ P, long long ld,
                                           bool kmode, const int* __restrict__ gather, int row0, int n_rows, int k0,
                                           int k_end, bool vec_ok) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    const int q = tid + i * GT_THREADS;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (kmode) {   // 8 consecutive lanes read the 128 contiguous bytes of one row
      const int row = row0 + q / 8, k = k0 + 4 * (q % 8);
      if (row < n_rows && k < k_end) {
        const long long gr = gather ? (long long)gather[row] : (long long)row;
        const float* src = P + gr * ld + k;
        if (vec_ok && k + 3 < k_end) {
          const float4 x = *reinterpret_cast<const float4*>(src);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (k + j < k_end) v[j] = src[j];
        }
      }
    } else {
      const int k = k0 + q / (ROWS / 4), row = row0 + 4 * (q % (ROWS / 4));
      if (k < k_end && row < n_rows) {
        const long long gk = gather ? (long long)gather[k] : (long long)k;
        const float* src = P + gk * ld + row;
        if (vec_ok && row + 3 < n_rows) {
          const float4 x = *reinterpret_cast<const float4*>(src);
          v[0] = x.x; v[1] = x.y; v[2] = x.z; v[3] = x.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (row + j < n_rows) v[j] = src[j];
        }
      }
    }
    r[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
}

template <int ROWS>
__device__ __forceinline__ void stash_tile(float* __restrict__ hi, float* __restrict__ lo, const float4 (&r)[ROWS / 32],
                                           bool kmode) {
  constexpr int LBO4 = gt_lbo(ROWS) / 16;   // chunk stride in float4 units
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < ROWS / 32; ++i) {
    const int q = tid + i * GT_THREADS;
    const float4 x = r[i];
    const float4 l = make_float4(umma::tf32_lo(x.x), umma::tf32_lo(x.y), umma::tf32_lo(x.z), umma::tf32_lo(x.w));
    if (kmode) {   // 4 consecutive k of one row = one 16-byte chunk entry
      const int row = q / 8, kq = q % 8;
      reinterpret_cast<float4*>(hi)[kq * LBO4 + row] = x;
      reinterpret_cast<float4*>(lo)[kq * LBO4 + row] = l;
    } else {       // 4 consecutive rows at one k: transpose into the chunk entries
      const int k = q / (ROWS / 4), row = 4 * (q % (ROWS / 4));
      const int base = ((k / 4) * LBO4 + row) * 4 + (k % 4);
      hi[base] = x.x; hi[base + 4] = x.y; hi[base + 8] = x.z; hi[base + 12] = x.w;
      lo[base] = l.x; lo[base + 4] = l.y; lo[base + 8] = l.z; lo[base + 12] = l.w;
    }
  }
}

template <int BN>
__global__ void __launch_bounds__(GT_THREADS, (BN > 128 ? 2 : 3)) gemm_tc_kernel(const GemmArgs g, int vec_a, int vec_b, int vec_c) {
  constexpr int BNC = BN > 128 ? 128 : BN;   // columns of one epilogue pass
  CARCA_DYN_SMEM(unsigned char, raw);
  GemmTcSmem<BN>& s = *reinterpret_cast<GemmTcSmem<BN>*>(raw);
  const int tid = threadIdx.x, w = tid / 32, lane = tid % 32;
  const int n0 = blockIdx.x * BN;
  const int gM = g.m_dev ? min(g.M, *g.m_dev) : g.M;
  if ((int)blockIdx.y * GT_BM >= gM) return;   // (whole CTA, before any barrier / TMEM allocation)
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int nk = k_end > k_begin ? (k_end - k_begin + GT_BK - 1) / GT_BK : 0;
  const bool issuer_warp = __shfl_sync(kFull, w, 0) == 0;

  if (w == 0) umma::tmem_alloc(&s.tmem_slot, BN);
  if (tid == 0) umma::mbar_init(&s.bar, 1);
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = s.tmem_slot;
  uint32_t n_commit = 0;   // commits so far on s.bar (running over the row tiles of this CTA)
  bool ok = true;

  // row tiles blockIdx.y, + gridDim.y, ...: the grid's y extent may be capped below the tile count (rows known on
  // the device only: a worst-case grid would launch thousands of CTAs that exit at once)
#pragma unroll 1
  for (int m0 = blockIdx.y * GT_BM; m0 < gM; m0 += gridDim.y * GT_BM) {
  __syncthreads();   // the previous tile's epilogue no longer reads the shared C tile (same bytes as the operands)

  const bool a_k = !g.transA, b_k = g.transB != 0;
  float4 ra[GT_BM / 32], rb[BN / 32];
  if (nk > 0) {
    fetch_tile<GT_BM>(ra, g.A, g.lda, a_k, a_k ? g.a_rows : nullptr, m0, gM, k_begin, k_end, vec_a);
    fetch_tile<BN>(rb, g.B, g.ldb, b_k, b_k ? nullptr : g.b_rows, n0, g.N, k_begin, k_end, vec_b);
  }
#pragma unroll 1
  for (int kt = 0; kt < nk; ++kt) {
    if (kt >= 1) ok = umma::mbar_wait(&s.bar, (n_commit - 1) & 1u) && ok;   // MMAs of step kt-1 read the buffers
    umma::fence_after_sync();
    stash_tile<GT_BM>(s.a[0], s.a[1], ra, a_k);
    stash_tile<BN>(s.b[0], s.b[1], rb, b_k);
    umma::fence_smem_to_async();
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    if (kt + 1 < nk) {   // next step's global loads fly while this step's MMAs run
      const int k0 = k_begin + (kt + 1) * GT_BK;
      fetch_tile<GT_BM>(ra, g.A, g.lda, a_k, a_k ? g.a_rows : nullptr, m0, gM, k0, k_end, vec_a);
      fetch_tile<BN>(rb, g.B, g.ldb, b_k, b_k ? nullptr : g.b_rows, n0, g.N, k0, k_end, vec_b);
    }
    if (issuer_warp) {
      if (umma::elect_one()) {
        constexpr uint32_t idesc = umma::idesc_tf32(BN);
        constexpr uint32_t a_lbo = gt_lbo(GT_BM), b_lbo = gt_lbo(BN);
#pragma unroll
        for (int p = 0; p < 3; ++p) {
          const uint64_t da = umma::smem_desc(umma::smem_u32(s.a[p == 1 ? 1 : 0]), a_lbo, 128);
          const uint64_t db = umma::smem_desc(umma::smem_u32(s.b[p == 2 ? 1 : 0]), b_lbo, 128);
#pragma unroll
          for (int ks = 0; ks < GT_BK / 8; ++ks)
            umma::mma_tf32(tmem0, da + (uint64_t)(ks * ((2 * a_lbo) >> 4)), db + (uint64_t)(ks * ((2 * b_lbo) >> 4)), idesc,
                           !(kt == 0 && p == 0 && ks == 0));
        }
        umma::commit(&s.bar);
      }
    }
    ++n_commit;
  }
  if (nk > 0) {
    ok = umma::mbar_wait(&s.bar, (n_commit - 1) & 1u) && ok;
    umma::fence_after_sync();
  }

  // ---- epilogue.  Pass over BNC columns: TMEM -> registers -> swizzled smem tile [128][BNC] (float4 j of row
  // r at j ^ (r % 32)) -> rows written by warps with coalesced 128-bit accesses
  float4* const ct = reinterpret_cast<float4*>(raw);
  const int trow_id = 32 * (w % 4) + lane, half = w / 4;
  const uint32_t trow = tmem0 + ((uint32_t)(32 * (w % 4)) << 16);
  const bool split = gridDim.z > 1;
  constexpr int Q = BNC / 4;            // float4 per tile row
  constexpr int RPW = 32 / Q > 0 ? 32 / Q : 1;   // rows one warp covers per iteration (BNC = 64: 2, 128: 1)
#pragma unroll 1
  for (int pass = 0; pass < BN / BNC; ++pass) {
    __syncthreads();   // operand buffers (pass 0) / the previous pass's tile are no longer read
#pragma unroll
    for (int c0 = 0; c0 < BNC / 2; c0 += 32) {
      const int col = half * (BNC / 2) + c0;   // within the pass
      float v[32];
      if (nk > 0) {
        umma::tmem_ld_1x32(trow + pass * BNC + col, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 8; ++j)
        ct[trow_id * Q + (((col / 4) + j) ^ (trow_id % 32 % Q))] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
    }
    __syncthreads();
#pragma unroll 1
    for (int r0 = w * RPW; r0 < GT_BM; r0 += 8 * RPW) {
      const int r = r0 + lane / Q, jq = lane % Q;
      const int gm = m0 + r, gn = n0 + pass * BNC + 4 * jq;
      if (gm >= gM || gn >= g.N) continue;
      const float4 t = ct[r * Q + (jq ^ (r % 32 % Q))];
      float x[4] = {t.x, t.y, t.z, t.w};
      float* dst = g.C + (long long)gm * g.ldc + gn;
      const bool full = vec_c && gn + 3 < g.N;
      if (split) {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < g.N) atomicAdd(dst + j, ok ? g.alpha * x[j] : NAN);
        continue;
      }
      const float rm = g.row_mask ? g.row_mask[gm] : 1.0f;
      float old[4] = {0.f, 0.f, 0.f, 0.f}, res[4] = {0.f, 0.f, 0.f, 0.f};
      if (g.accumulate) {
        if (full) {
          const float4 o = *reinterpret_cast<const float4*>(dst);
          old[0] = o.x; old[1] = o.y; old[2] = o.z; old[3] = o.w;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (gn + j < g.N) old[j] = dst[j];
        }
      }
      if (g.R) {
        const float* rp = g.R + (long long)(g.r_mod ? gm % g.r_mod : gm) * g.ldr + gn;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < g.N) res[j] = rp[j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if (gn + j >= g.N) continue;
        float y = ok ? g.alpha * x[j] : NAN;   // an MMA completion wait timed out: poison the result, never hang
        if (g.bias) y += g.bias[gn + j];
        y += old[j];
        if (g.act == 1) y = y > 0.f ? y : kLeakySlope * y;
        if (g.drop.p > 0.f) y *= drop_factor(g.drop, (unsigned long long)gm * (unsigned long long)g.N + gn + j);
        x[j] = (y + res[j]) * rm;
      }
      if (full) {
        *reinterpret_cast<float4*>(dst) = make_float4(x[0], x[1], x[2], x[3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (gn + j < g.N) dst[j] = x[j];
      }
    }
  }
  umma::fence_before_sync();   // this tile's TMEM reads are ordered before the next tile's MMAs
  }   // row tiles
  umma::fence_before_sync();
  __syncthreads();
  if (w == 0) umma::tmem_free(tmem0, BN);
}

}  // namespace carca
#endif  // CARCA_EMU
