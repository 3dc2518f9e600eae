// Chained GEMMs of one encoder block tail in ONE kernel (bf16 rows pipeline, d = 64; src/carca.py:305-318 followed by
// :238-240 of the NEXT block, or by the decoder's key / value projections after the last block):
//
//   F  = LeakyReLU(s2 W1^T + b1)                               G1, epilogue 1 -> shared memory (operand layout)
//   x' = F W2^T + b2 + s2,  n = LayerNorm_next(x')             G2, epilogue 2 -> fp32 rows of n (residual stream) to HBM,
//                                                                  x' and n as operands to shared memory
//   Q = n Wq^T + bq, K = x' Wk^T + bk, V = x' Wv^T + bv        G3 (three accumulators), epilogue 3 -> the attention's
//                                                                  operand layouts (rows_attn_tc.cuh)          [tail 1]
//   Kd = n dWk^T + bk (+ context terms), u = <n dWv^T + bv, wf> G3, epilogue 3 -> the decoder's inputs            [tail 2]
//
// With the separate kernels every link is a launch whose ~450 tiles (8192 Beauty users) make a single short wave —
// prologue, one tile, drain — and F, x', n travel through HBM as bf16 operand tiles in between.  Here a 128-row tile
// stays on its SM for the whole chain: all five 64 x 64 weight matrices are resident in shared memory (40 KB), the
// intermediate operands never leave it, and per tile only the s2 tile comes in and Q / K / V (+ the fp32 rows) go out.
// A CTA works on one tile at a time (the links are dependent); two CTAs share an SM (105 KB of shared memory and 256
// TMEM columns each), so one CTA's epilogue runs under the other's MMAs.
// Roles: warp 0 producer (the next tile's operand is fetched during the current chain), warp 1 MMA issuer, warps 2..5
// epilogues (thread = row).
#pragma once
#include "rows_bf16.cuh"
#ifndef CARCA_EMU

namespace carca {
namespace rows {

constexpr int FC_THREADS = 192;
constexpr int FC_D = 64;
constexpr int FC_W_BYTES = 8 * FC_D * 16;          // one packed 64 x 64 bf16 weight matrix: 8 KB
constexpr int FC_TILE_BYTES = 8 * TILE * 16;       // one operand tile: 16 KB

struct FfnChainArgs {
  const bf16* A;            // s2 operand tiles (LayerNorm 2 output) [n_tiles][8][128][8]
  const bf16* W12;          // packed W1 | W2 (contiguous)
  const bf16* W3;           // tail 1: packed Wq | Wk | Wv of the next block; tail 2: decoder Wk | Wv
  const float *b1, *b2;
  const float* resid;       // s2 fp32 rows (null: no residual)
  const float *ln_g, *ln_b; // the NEXT LayerNorm (LN1 of block b + 1, or the final norm)
  float* out_f32;           // n rows, fp32
  int tail;                 // 0: none; 1: Q / K / V for the tensor-core attention; 2: decoder keys / value folds
  const float *bq, *bk, *bv;
  bf16 *Qt, *Kk, *Vm;       // tail 1 outputs (EPI_BIAS_TILE / EPI_KMAJ / EPI_VMN layouts)
  long long ld_rows;
  float* Kd;                // tail 2: decoder keys, fp32 rows [R, 64]
  const float* McQ;         //         [64][8]
  float* KM;                //         km[r][h][k]
  const float* wf;          //         scorer weight [64]
  float* U;                 //         u[r][h]
  const int* n_rows;
  int H;
  int* status;
};

struct FfnChainSmem {
  unsigned char w[5][FC_W_BYTES];
  unsigned char slot[2][FC_TILE_BYTES];   // slot t & 1: the tile's s2 operand, then F, then n
  unsigned char xs[FC_TILE_BYTES];        // x' operand
  float prm[16 * FC_D];                   // b1 | b2 | ln_g | ln_b | bq | bk | bv | wf | McQ^T [8][64]
  uint64_t a_full[2], a_empty[2], w_full, g1_full, f_ready, g2_full, x_ready, g3_full;
  uint32_t tmem_slot;
};

__global__ void __launch_bounds__(FC_THREADS, 2) rows_ffn_chain_kernel(const FfnChainArgs a) {
  constexpr int D = FC_D;
  extern __shared__ __align__(128) unsigned char fc_raw[];
  FfnChainSmem& s = *reinterpret_cast<FfnChainSmem*>(fc_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_w3 = a.tail == 1 ? 3 : (a.tail == 2 ? 2 : 0);
  pdl_trigger();

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) { umma::mbar_init(&s.a_full[i], 1); umma::mbar_init(&s.a_empty[i], 1); }
    umma::mbar_init(&s.w_full, 1);
    umma::mbar_init(&s.g1_full, 1);
    umma::mbar_init(&s.f_ready, 4);
    umma::mbar_init(&s.g2_full, 1);
    umma::mbar_init(&s.x_ready, 4);
    umma::mbar_init(&s.g3_full, 1);
  }
  if (warp == 1) umma::tmem_alloc(&s.tmem_slot, 256);
  for (int i = threadIdx.x; i < D; i += FC_THREADS) {
    s.prm[i] = a.b1[i];
    s.prm[D + i] = a.b2[i];
    s.prm[2 * D + i] = a.ln_g[i];
    s.prm[3 * D + i] = a.ln_b[i];
    s.prm[4 * D + i] = a.bq ? a.bq[i] : 0.f;
    s.prm[5 * D + i] = a.bk ? a.bk[i] : 0.f;
    s.prm[6 * D + i] = a.bv ? a.bv[i] : 0.f;
    s.prm[7 * D + i] = a.wf ? a.wf[i] : 0.f;
  }
  if (a.tail == 2)
    for (int i = threadIdx.x; i < 8 * D; i += FC_THREADS) s.prm[8 * D + (i % 8) * D + i / 8] = a.McQ[i];
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = s.tmem_slot;
  pdl_wait();
  const int R = *a.n_rows;
  const int n_tiles = (R + TILE - 1) / TILE;

  if (warp == 0) {
    // ===== producer
    if (lane == 0 && (int)blockIdx.x < n_tiles) {
      mbar_expect_tx(&s.w_full, (uint32_t)((2 + n_w3) * FC_W_BYTES));
      bulk_g2s(s.w[0], a.W12, 2 * FC_W_BYTES, &s.w_full);
      for (int i = 0; i < n_w3; ++i) bulk_g2s(s.w[2 + i], a.W3 + (size_t)i * D * D, FC_W_BYTES, &s.w_full);
      uint32_t t = 0;
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++t) {
        const uint32_t sl = t & 1u, ph = (t >> 1) & 1u;
        if (!wait_or_flag(&s.a_empty[sl], ph ^ 1u, a.status, 2048)) break;
        mbar_expect_tx(&s.a_full[sl], FC_TILE_BYTES);
        bulk_g2s(s.slot[sl], reinterpret_cast<const unsigned char*>(a.A) + (size_t)tile * FC_TILE_BYTES, FC_TILE_BYTES, &s.a_full[sl]);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer
    if ((int)blockIdx.x < n_tiles) {
      constexpr uint32_t idesc = idesc_bf16(D);
      constexpr uint32_t a_lbo = TILE * 16, b_lbo = D * 16;
      const uint32_t w_addr = umma::smem_u32(s.w[0]), x_addr = umma::smem_u32(s.xs);
      auto gemm = [&](uint32_t d_tmem, uint32_t a_addr, int wi) {
#pragma unroll
        for (int ks = 0; ks < 4; ++ks)
          mma_bf16_ss(d_tmem, umma::smem_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128),
                      umma::smem_desc(w_addr + wi * FC_W_BYTES + ks * 2 * b_lbo, b_lbo, 128), idesc, ks != 0);
      };
      bool ok = wait_or_flag(&s.w_full, 0, a.status, 2048);
      uint32_t t = 0;
      for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++t) {
        const uint32_t sl = t & 1u, ph = (t >> 1) & 1u, tp = t & 1u;
        const uint32_t sa = umma::smem_u32(s.slot[sl]);
        ok = wait_or_flag(&s.a_full[sl], ph, a.status, 2048);
        if (!ok) break;
        umma::fence_after_sync();
        if (umma::elect_one()) {
          gemm(tmem0, sa, 0);
          umma::commit(&s.g1_full);
        }
        __syncwarp();
        ok = wait_or_flag(&s.f_ready, tp, a.status, 2048);
        if (!ok) break;
        umma::fence_after_sync();
        if (umma::elect_one()) {
          gemm(tmem0, sa, 1);
          umma::commit(&s.g2_full);
        }
        __syncwarp();
        ok = wait_or_flag(&s.x_ready, tp, a.status, 2048);
        if (!ok) break;
        umma::fence_after_sync();
        if (umma::elect_one()) {
          if (a.tail == 1) {
            gemm(tmem0 + 64, sa, 2);        // Q from n
            gemm(tmem0 + 128, x_addr, 3);   // K from x'
            gemm(tmem0 + 192, x_addr, 4);   // V from x'
          } else if (a.tail == 2) {
            gemm(tmem0 + 64, sa, 2);        // decoder keys and values from the encoded profile n
            gemm(tmem0 + 128, sa, 3);
          }
          if (a.tail != 0) umma::commit(&s.g3_full);
          umma::commit(&s.a_empty[sl]);
        }
        __syncwarp();
      }
    }
  } else {
    // ===== epilogues: thread = row
    const int quarter = warp & 3, row = 32 * quarter + lane;
    const uint32_t tb = tmem0 + ((uint32_t)(32 * quarter) << 16);
    const int H = a.H, DH = D / H;
    uint32_t t = 0;
    bool ok = true;
    for (int tile = blockIdx.x; tile < n_tiles && ok; tile += gridDim.x, ++t) {
      const uint32_t sl = t & 1u, tp = t & 1u;
      uint4* slot_row = reinterpret_cast<uint4*>(s.slot[sl]) + row;      // k-group kg of this row: slot_row[kg * 128]
      uint4* xs_row = reinterpret_cast<uint4*>(s.xs) + row;
      const long long r = (long long)tile * TILE + row;
      const bool live = r < R;
      const bool use_res = a.resid != nullptr && live;
      float rs[2][32];
      if (use_res) {      // the residual row of epilogue 2: requested now, its latency hides behind G1 / epilogue 1 / G2
#pragma unroll
        for (int c = 0; c < 2; ++c)
#pragma unroll
          for (int q = 0; q < 4; ++q) ldg256(a.resid + r * D + 32 * c + 8 * q, *reinterpret_cast<float(*)[8]>(&rs[c][8 * q]));
      }
      // ----- epilogue 1: F = LeakyReLU(acc + b1) -> slot (over the s2 operand G1 has consumed)
      ok = wait_or_flag(&s.g1_full, tp, a.status, 4096);
      if (!ok) break;
      umma::fence_after_sync();
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        umma::tmem_ld_1x32(tb + 32 * c, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          v[e] += s.prm[32 * c + e];
          v[e] = v[e] > 0.f ? v[e] : kLeakySlope * v[e];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) slot_row[(4 * c + q) * 128] = pack8(&v[8 * q]);
      }
      umma::fence_smem_to_async();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.f_ready);
      // ----- epilogue 2: x' = acc + b2 (+ s2), n = LayerNorm(x')
      ok = wait_or_flag(&s.g2_full, tp, a.status, 4096);
      if (!ok) break;
      umma::fence_after_sync();
      float sum = 0.f, sq = 0.f;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
        umma::tmem_ld_1x32(tb + 32 * c, v);
#pragma unroll
        for (int e = 0; e < 32; ++e) {
          v[e] += s.prm[D + 32 * c + e] + (use_res ? rs[c][e] : 0.f);
          sum += v[e];
          sq = fmaf(v[e], v[e], sq);
          rs[c][e] = v[e];
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) xs_row[(4 * c + q) * 128] = pack8(&v[8 * q]);
      }
      const float mean = sum * (1.0f / D);
      const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + kLnEps);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        float v[32];
#pragma unroll
        for (int e = 0; e < 32; ++e) v[e] = (rs[c][e] - mean) * rstd * s.prm[2 * D + 32 * c + e] + s.prm[3 * D + 32 * c + e];
        if (live) {
#pragma unroll
          for (int q = 0; q < 4; ++q) stg256(a.out_f32 + r * D + 32 * c + 8 * q, &v[8 * q]);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) slot_row[(4 * c + q) * 128] = pack8(&v[8 * q]);
      }
      umma::fence_smem_to_async();
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.x_ready);
      // ----- epilogue 3
      if (a.tail == 0) continue;             // (dot-product decoder after the last block: it reads the fp32 rows of n)
      ok = wait_or_flag(&s.g3_full, tp, a.status, 4096);
      if (!ok) break;
      umma::fence_after_sync();
      if (a.tail == 1) {
#pragma unroll 1
        for (int j = 0; j < 3; ++j) {          // Q, K, V
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            float v[32];
            umma::tmem_ld_1x32(tb + 64 + 64 * j + 32 * c, v);
#pragma unroll
            for (int e = 0; e < 32; ++e) {
              v[e] += s.prm[(4 + j) * D + 32 * c + e];
              if (j > 0) v[e] = live ? v[e] : 0.f;      // rows past the batch: exact zeros (an MMA reads them)
            }
            if (j == 0) {
#pragma unroll
              for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(a.Qt + tile_off<D>(r, 4 * c + q)) = pack8(&v[8 * q]);
            } else if (j == 1) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(a.Kk + ((long long)(4 * c + q) * a.ld_rows + r) * 8) = pack8(&v[8 * q]);
            } else {
              const int FG = DH / 8;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                const int g = 4 * c + q, h = g / FG, jj = g % FG;
                *reinterpret_cast<uint4*>(a.Vm + (((long long)h * (a.ld_rows >> 3) + (r >> 3)) * FG + jj) * 64 + (r & 7) * 8) =
                    pack8(&v[8 * q]);
              }
            }
          }
        }
      } else {
        // decoder keys (fp32 rows + their context terms) and value folds, as EPI_KDEC / EPI_VDOT of rows_gemm_kernel
        float km[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) km[k] = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          umma::tmem_ld_1x32(tb + 64 + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] += s.prm[5 * D + 32 * c + e];
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) stg256(a.Kd + r * D + 32 * c + 8 * q, &v[8 * q]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float* mk = s.prm + 8 * D + k * D + 32 * c;
#pragma unroll
            for (int e = 0; e < 32; ++e) km[k] = fmaf(v[e], mk[e], km[k]);
          }
          if ((32 * (c + 1)) % DH == 0) {
            if (live) {
              float4* o = reinterpret_cast<float4*>(a.KM + (r * H + (32 * c) / DH) * 8);
              o[0] = make_float4(km[0], km[1], km[2], km[3]);
              o[1] = make_float4(km[4], km[5], km[6], km[7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) km[k] = 0.f;
          }
        }
        float u = 0.f;
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          float v[32];
          umma::tmem_ld_1x32(tb + 128 + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) u = fmaf(v[e] + s.prm[6 * D + 32 * c + e], s.prm[7 * D + 32 * c + e], u);
          if ((32 * (c + 1)) % DH == 0) {
            if (live) a.U[r * H + (32 * c) / DH] = u;
            u = 0.f;
          }
        }
      }
      umma::fence_before_sync();     // the accumulators are rewritten by the next tile's MMAs
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_free(tmem0, 256);
}

}  // namespace rows
}  // namespace carca
#endif  // CARCA_EMU
