// Causal self-attention of the bf16 packed-rows pipeline on the tensor cores (src/carca.py:246-256 inside :297-318),
// fused with the LN1 residual and LayerNorm 2 — the tcgen05 replacement of rows_attn_ln_kernel for L <= 129.
//
// Rows are packed user after user, so a 128-row query tile only needs the keys from the segment start of its FIRST
// row to its own last row: a WINDOW of win = 128 + (< L) keys that may reach back into the previous tile(s).  Per
// (tile, head) the kernel runs, FlashAttention style but with the whole window in one pass:
//
//   S[128 x win] = Q_h K_h^T          tcgen05.mma kind::f16, M = 128, N = win (16-aligned, run-time), K = d / H
//   P = exp2((S - rowmax) * scale)    thread = query row; key j is visible to row r iff seg(r) <= j <= r (same user,
//                                     causal); everything else is an exact 0.  A warp only touches the columns its 32
//                                     rows can see (<= 32 + L - 1 of the window), the rest of P is zero-filled once
//   O_h[128 x d/H] = P V_h            A = P (bf16, written K-major into shared memory), B = V_h MN-major (the natural
//                                     row layout of V: no transpose anywhere)
//   after the last head:              x = O / rowsum + LN1 residual, LayerNorm 2 -> fp32 rows + operand tiles for FFN-1
//
// The producing GEMM writes the operands in the layouts the copies want (rows_gemm_kernel, EPI_BIAS_TILE / EPI_KMAJ /
// EPI_VMN): Q as operand tiles (a head's K-slice of a tile is one contiguous 128 x d/H block), K as ONE K-major operand
// over all rows [d/8][Rp][8] (a window of a head's 8-feature group is one contiguous copy), V as MN-major 8 x 8 blocks
// per head [H][Rp/8][d/H/8][8 keys][8 features] (a head's window is one contiguous copy).  All loads are cp.async.bulk
// on the TMA engine completing on mbarriers, two stages deep.
//
// Roles: warp 0 producer, warp 1 MMA issuer, warps 2..5 softmax + LayerNorm (thread = row, TMEM lane quarter = warp % 4).
// TMEM: S at column 0, O (all heads side by side = the attention output row) behind it.  At d = 64 with L <= 57 a CTA
// needs 256 TMEM columns and 113 KB of shared memory, so TWO CTAs share an SM and one CTA's softmax runs under the
// other's MMAs.
#pragma once
#include "rows_bf16.cuh"
#ifndef CARCA_EMU

namespace carca {
namespace rows {

constexpr int AT_THREADS = 192;
constexpr int AT_NST = 2;

struct AttnTcArgs {
  const bf16* Qt;           // Q operand tiles [tile][D/8][128][8]
  const bf16* Kk;           // K-major over all rows [D/8][Rp][8]
  const bf16* Vm;           // MN-major blocks [H][Rp/8][DH/8][8][8]
  long long Rp;
  const float* QN;          // LN1(x) rows (residual, src/carca.py:302)
  const int *row_src, *row_seg, *n_rows;
  const float *ln_g, *ln_b;
  float* S2;                // LN2 rows, fp32
  bf16* S2A;                // LN2 operand tiles
  int residual;
  int win_max;              // largest window (multiple of 16, <= 256): shared-memory strides
  int s_cols;               // TMEM columns reserved for S (O starts here)
  int tmem_cols;            // allocation (power of two)
  int vswap;                // development: swap the LBO / SBO roles of the MN-major V descriptor
  int* status;
};

template <int D, int H>
struct AttnTcCfg {
  static constexpr int DH = D / H;
  static constexpr int Q_BYTES = 128 * DH * 2;
  static size_t stage_bytes(int win_max) { return (size_t)Q_BYTES + 2 * (size_t)win_max * DH * 2; }
  static size_t smem_bytes(int win_max) {
    return AT_NST * stage_bytes(win_max) + (size_t)128 * win_max * 2 + 2 * D * sizeof(float) + 128;
  }
};

__device__ __forceinline__ void attn_window(const int* __restrict__ row_seg, int tile, int win_max, int& w0, int& win) {
  const int end = (tile + 1) * TILE;
  win = min((end - __ldg(row_seg + (long long)tile * TILE) + 15) & ~15, win_max);
  w0 = end - win;
}

template <int D, int H>
__global__ void __launch_bounds__(AT_THREADS, (D == 64 ? 2 : 1)) rows_attn_tc_kernel(const AttnTcArgs a) {
  using Cfg = AttnTcCfg<D, H>;
  constexpr int DH = Cfg::DH, FG = DH / 8;       // 8-feature groups per head
  extern __shared__ __align__(128) unsigned char at_raw[];
  const int WM = a.win_max;
  const size_t stage_bytes = (size_t)Cfg::Q_BYTES + 2 * (size_t)WM * DH * 2;
  unsigned char* p_smem = at_raw + AT_NST * stage_bytes;                 // P [win/8][128][8] bf16
  float* prm = reinterpret_cast<float*>(p_smem + (size_t)128 * WM * 2);  // ln_g | ln_b
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + 2 * D);
  uint64_t* full = bars;                 // [NST] loads landed
  uint64_t* empty = bars + AT_NST;       // [NST] MMAs have read the stage
  uint64_t* s_full = empty + AT_NST;     // S complete (and the previous P consumed)
  uint64_t* p_full = s_full + 1;         // P written, S read (4 warp arrivals)
  uint64_t* o_full = p_full + 1;         // O of the tile complete
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int s = 0; s < AT_NST; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    umma::mbar_init(s_full, 1);
    umma::mbar_init(p_full, 4);
    umma::mbar_init(o_full, 1);
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, a.tmem_cols);
  for (int i = threadIdx.x; i < D; i += AT_THREADS) {
    prm[i] = a.ln_g[i];
    prm[D + i] = a.ln_b[i];
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = *tmem_slot;
  pdl_wait();
  const int R = *a.n_rows;
  const int n_tiles = (R + TILE - 1) / TILE;

  if (warp == 0) {
    // ===== producer: per (tile, head) the head's Q block, K window (one copy per 8-feature group) and V window
    if (lane == 0) {
      uint32_t it = 0;
      bool ok = true;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
        int w0, win;
        attn_window(a.row_seg, tile, WM, w0, win);
        for (int h = 0; h < H && ok; ++h, ++it) {
          const uint32_t s = it % AT_NST, ph = (it / AT_NST) & 1u;
          ok = wait_or_flag(&empty[s], ph ^ 1u, a.status, 64);
          if (!ok) break;
          unsigned char* st = at_raw + s * stage_bytes;
          mbar_expect_tx(&full[s], (uint32_t)(Cfg::Q_BYTES + 2 * win * DH * 2));
          bulk_g2s(st, a.Qt + ((long long)tile * (D / 8) + h * FG) * (TILE * 8), Cfg::Q_BYTES, &full[s]);
          unsigned char* ks = st + Cfg::Q_BYTES;
#pragma unroll
          for (int f = 0; f < FG; ++f)
            bulk_g2s(ks + (size_t)f * win * 16, a.Kk + ((long long)(h * FG + f) * a.Rp + w0) * 8, (uint32_t)win * 16, &full[s]);
          bulk_g2s(ks + (size_t)WM * DH * 2, a.Vm + ((long long)h * (a.Rp / 8) + w0 / 8) * (FG * 64), (uint32_t)win * DH * 2,
                   &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer.  Item i = (tile, head); order: PV(i-1), S(i) — S(i) may only overwrite the S columns once the
    // softmax of item i-1 has read them (p_full), and PV(i-1) needs that same P.
    const uint32_t q_off = 0, k_off = Cfg::Q_BYTES, v_off = Cfg::Q_BYTES + (uint32_t)WM * DH * 2;
    const uint32_t smem0 = umma::smem_u32(at_raw), p_addr = umma::smem_u32(p_smem);
    const uint32_t o_tmem = tmem0 + a.s_cols;
    const uint32_t v_lbo = a.vswap ? 128u : (uint32_t)DH * 16, v_sbo = a.vswap ? (uint32_t)DH * 16 : 128u;
    uint32_t it = 0;
    bool ok = true;
    int prev_win = 0, prev_h = 0;
    uint32_t prev_s = 0;
    auto issue_pv = [&]() {   // P V of the previous item, then release its stage (and publish O after the last head)
      if (umma::elect_one()) {
        const uint32_t idesc = idesc_bf16(DH) | (1u << 16);   // B operand MN-major
        const uint32_t vb = smem0 + prev_s * (uint32_t)stage_bytes + v_off;
        for (int ks = 0; ks < prev_win / 16; ++ks) {
          const uint64_t da = umma::smem_desc(p_addr + ks * 2 * (TILE * 16), TILE * 16, 128);
          const uint64_t db = umma::smem_desc(vb + ks * 2 * (DH * 16), v_lbo, v_sbo);
          mma_bf16_ss(o_tmem + prev_h * DH, da, db, idesc, ks != 0);
        }
        umma::commit(&empty[prev_s]);
        if (prev_h == H - 1) umma::commit(o_full);
      }
      __syncwarp();
    };
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x) {
      int w0, win;
      attn_window(a.row_seg, tile, WM, w0, win);
      for (int h = 0; h < H && ok; ++h, ++it) {
        const uint32_t s = it % AT_NST, ph = (it / AT_NST) & 1u;
        if (it > 0) {
          ok = wait_or_flag(p_full, (it - 1) & 1u, a.status, 128);
          if (!ok) break;
          umma::fence_after_sync();
          issue_pv();
        }
        ok = wait_or_flag(&full[s], ph, a.status, 256);
        if (!ok) break;
        umma::fence_after_sync();
        if (umma::elect_one()) {
          const uint32_t idesc = idesc_bf16(win);
          const uint32_t qb = smem0 + s * (uint32_t)stage_bytes + q_off, kb = smem0 + s * (uint32_t)stage_bytes + k_off;
          const uint32_t k_lbo = (uint32_t)win * 16;
#pragma unroll
          for (int ks = 0; ks < DH / 16; ++ks) {
            const uint64_t da = umma::smem_desc(qb + ks * 2 * (TILE * 16), TILE * 16, 128);
            const uint64_t db = umma::smem_desc(kb + ks * 2 * k_lbo, k_lbo, 128);
            mma_bf16_ss(tmem0, da, db, idesc, ks != 0);
          }
          umma::commit(s_full);
        }
        __syncwarp();
        prev_win = win; prev_h = h; prev_s = s;
      }
    }
    if (ok && it > 0) {
      ok = wait_or_flag(p_full, (it - 1) & 1u, a.status, 128);
      if (ok) {
        umma::fence_after_sync();
        issue_pv();
      }
    }
  } else {
    // ===== softmax + LayerNorm: thread = query row of the tile
    const int quarter = warp & 3, row = 32 * quarter + lane;
    const uint32_t tb = tmem0 + ((uint32_t)(32 * quarter) << 16);
    const uint32_t o_tb = tb + a.s_cols;
    const float sc = 1.4426950408889634f * rsqrtf((float)DH);
    uint4* p_row = reinterpret_cast<uint4*>(p_smem) + row;     // chunk c (8 keys) of this row: p_row[c * 128]
    uint32_t it = 0, tcount = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++tcount) {
      int w0, win;
      attn_window(a.row_seg, tile, WM, w0, win);
      const long long r = (long long)tile * TILE + row;
      const bool live = r < R;
      const int src = live ? __ldg(a.row_src + r) : (int)0x80000000u;
      const bool valid = src >= 0;                              // (a padding query row attends to nothing)
      const int lo = valid ? max(__ldg(a.row_seg + r) - w0, 0) : 0x7fffffff;
      const int hi = valid ? (int)(r - w0) : -1;
      int wlo = lo, whi = hi;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        wlo = min(wlo, __shfl_xor_sync(kFull, wlo, o));
        whi = max(whi, __shfl_xor_sync(kFull, whi, o));
      }
      const int c_lo = wlo >> 4, c_hi = min(whi, win - 1) >> 4;   // 16-column chunks this warp computes (none: c_lo > c_hi)
      if (a.residual && live) {      // the LayerNorm tail's residual row: into L2 now, it is read after the last head
#pragma unroll
        for (int c = 0; c < D / 32; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(a.QN + r * D + 32 * c));
      }
      float zinv[H];
#pragma unroll
      for (int hh = 0; hh < H; ++hh) zinv[hh] = 0.f;
      for (int h = 0; h < H && ok; ++h, ++it) {
        ok = wait_or_flag(s_full, it & 1u, a.status, 512);
        if (!ok) break;
        umma::fence_after_sync();
        float m = -INFINITY;
#pragma unroll 1
        for (int c = c_lo; c <= c_hi; ++c) {
          float v[16];
          umma::tmem_ld_1x16(tb + 16 * c, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = 16 * c + e;
            m = fmaxf(m, (j >= lo && j <= hi) ? v[e] : -INFINITY);
          }
        }
        const float msc = valid ? m * sc : 0.f;
        float z = 0.f;
        if (h == 0) {   // columns no row of this warp can see: exact zeros, written once per tile (the range is per tile)
          const uint4 zero = make_uint4(0, 0, 0, 0);
          for (int c = 0; c < (win >> 3); ++c)
            if (c < 2 * c_lo || c > 2 * c_hi + 1) p_row[c * 128] = zero;
        }
#pragma unroll 1
        for (int c = c_lo; c <= c_hi; ++c) {
          float v[16];
          umma::tmem_ld_1x16(tb + 16 * c, v);
#pragma unroll
          for (int e = 0; e < 16; ++e) {
            const int j = 16 * c + e;
            v[e] = (j >= lo && j <= hi) ? ex2f(fmaf(v[e], sc, -msc)) : 0.f;
            z += v[e];
          }
          p_row[(2 * c) * 128] = pack8(&v[0]);
          p_row[(2 * c + 1) * 128] = pack8(&v[8]);
        }
        {
          const float zi = z > 0.f ? 1.0f / z : 0.f;
#pragma unroll
          for (int hh = 0; hh < H; ++hh) zinv[hh] = hh == h ? zi : zinv[hh];
        }
        umma::fence_smem_to_async();
        umma::fence_before_sync();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
      }
      if (!ok) break;
      // ----- O / rowsum + residual -> LayerNorm 2 (src/carca.py:302-303), as the EPI_LN epilogue of the GEMM does it
      // The residual row (fp32, HBM) is the only global read of this tail and a dependent one: its first chunk is
      // requested before the wait for O (the rest of the row was sent to L2 when the tile started), chunk c + 1 while
      // chunk c is processed (ncu: a quarter of the kernel's samples sat on this load when it was issued in place).
      const bool use_res = a.residual && live;
      float rs[32], rn[32];
#pragma unroll
      for (int e = 0; e < 32; ++e) rs[e] = 0.f;
      if (use_res) {
#pragma unroll
        for (int q = 0; q < 4; ++q) ldg256(a.QN + r * D + 8 * q, *reinterpret_cast<float(*)[8]>(&rs[8 * q]));
      }
      ok = wait_or_flag(o_full, tcount & 1u, a.status, 1024);
      if (!ok) break;
      umma::fence_after_sync();
      float sum = 0.f, sq = 0.f;
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        if (use_res && c + 1 < D / 32) {
#pragma unroll
          for (int q = 0; q < 4; ++q) ldg256(a.QN + r * D + 32 * (c + 1) + 8 * q, *reinterpret_cast<float(*)[8]>(&rn[8 * q]));
        }
        float v[32];
        umma::tmem_ld_1x32(o_tb + 32 * c, v);
        float zi = zinv[0];
#pragma unroll
        for (int h = 1; h < H; ++h) zi = (32 * c) / DH == h ? zinv[h] : zi;
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          v[e] = fmaf(v[e], zi, rs[e]);
          sum += v[e];
          sq = fmaf(v[e], v[e], sq);
        }
        umma::tmem_st_x32(o_tb + 32 * c, v, 0);
        if (use_res) {
#pragma unroll
          for (int e = 0; e < 32; ++e) rs[e] = rn[e];
        }
      }
      umma::tmem_st_wait();
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + kLnEps);
#pragma unroll 1
      for (int c = 0; c < D / 32; ++c) {
        float v[32];
        umma::tmem_ld_1x32(o_tb + 32 * c, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = (v[e] - mean) * rstd * prm[32 * c + e] + prm[D + 32 * c + e];
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) stg256(a.S2 + r * D + 32 * c + 8 * q, &v[8 * q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(a.S2A + tile_off<D>(r, 4 * c + q)) = pack8(&v[8 * q]);
      }
      umma::fence_before_sync();   // the O columns are rewritten by the next tile's first P V
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_free(tmem0, a.tmem_cols);
}

}  // namespace rows
}  // namespace carca
#endif  // CARCA_EMU
