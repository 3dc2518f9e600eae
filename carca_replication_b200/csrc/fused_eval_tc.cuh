// Tensor-core (tcgen05) version of the fused inference forward for d = 64, 2 or 4 heads, L <= 256 with at
// most 64 non-padding positions per user.
// Same contract as fused_eval_kernel (fused_eval.cuh); reference path replaced:
// CARCA.forward in eval mode, src/carca.py:411-431 with :85-95, :228-265, :297-318, :338-365.
//
// Design (per CTA: 2 users = 128 rows, one TMEM lane per row, 256 threads = 2 threads per row):
//   * ACTIVATIONS LIVE IN TENSOR MEMORY AND REGISTERS.  Every [128 x 64] MMA operand / accumulator
//     is a block of 64 TMEM columns; LayerNorm, softmax, bias/activation epilogues are row
//     operations on registers loaded with one wide tcgen05.ld and written back with tcgen05.st.
//     Activations never touch shared memory or HBM.
//   * every projection / QK^T / PV product is a tcgen05.mma with M = 128, the A operand read
//     straight from TMEM, the B operand (weights, K, V) from shared memory in the K-major
//     no-swizzle layout of umma.cuh, accumulating into other TMEM columns.
//   * fp32-grade accuracy on the tf32 tensor cores via the 3xTF32 split (hi*hi + lo*hi + hi*lo);
//     biases ride along as one extra K step against a constant [1,0,..] column block.
//   * feature ownership: thread (row, half) owns, for every head h, columns
//     h*DH + half*DH/2 .. + DH/2 of a 64-wide activation (DH = 64 / heads).  With that split the
//     per-head attention output lands on the thread that holds the residual, so LN1 -> attention
//     residual -> LN2 and FFN residual -> next LN1 run register to register.
//   * the two heads of a pair are processed together: both score MMAs, one softmax phase, both
//     PV MMAs — two completion waits per pair instead of four.
//   * shared memory holds only B operands: a 2-slot weight ring (cp.async prefetch), K (64 KB),
//     V (K-major over keys, rows = (head, user, feature), padded chunk stride so the thread-per-key
//     scalar stores are bank-conflict free), plus the small LayerNorm / context tables.
// TMEM column map: X_HI 0, X_LO 64, QN_HI 128, QN_LO 192, ACC_Q 256, ACC_K 320, ACC_V 384, ONES 448.
//   self-attention : scores/P of the pair's heads at 0..127 and 320..447, O at 128.. (+ (h%2)*2*DH)
//   cross-attention: scores/P_hi at 0..63 / 64..127, P_lo at 320.. / 384.., O at 128 + h*DH
#pragma once
#include "common.cuh"
#include "fused_eval.cuh"
#include "tmem_io.cuh"
#include "umma.cuh"

#ifndef CARCA_EMU
namespace carca {

constexpr int TC_THREADS = 256;
constexpr int TC_WFLOATS = 18 * 64 * 4;              // packed weight: 16 k-chunks + 2 bias chunks, hi or lo
constexpr int TC_VLBO = 128 * 16 + 16;               // bytes between key chunks of the V operand (padded)
constexpr int TC_VFLOATS = 16 * (TC_VLBO / 4);       // floats of one V image (hi or lo): 16 chunks of 4 keys
constexpr int TC_LNROWS = 4 * FMAXB + 2;
enum { C_XHI = 0, C_XLO = 64, C_QNHI = 128, C_QNLO = 192, C_ACCQ = 256, C_ACCK = 320, C_ACCV = 384, C_ONES = 448 };

struct TcBlockW {
  const float *ln1_g, *ln1_b, *ln2_g, *ln2_b;
  const float *wq, *wk, *wv, *w1, *w2;               // packed [hi 4608 | lo 4608] with the bias folded in
};

struct TcArgs {
  const float* Tfold;
  const float* Mc;
  const float* pos;
  const int* p_x;
  const float* p_c;
  const int* o_x;
  const float* o_c;
  float* y;
  long long ldy;
  int col0;
  int B, L, T, C, H, n_blocks, residual_sa, residual_ca, decoder;
  TcBlockW blk[FMAXB];
  const float *fn_g, *fn_b;
  const float *dwk, *dwv;                            // packed decoder K / V projections
  const float *dwf, *dbf;
  const float *TQ, *tw;                              // folded candidate tables: WQ T[i] + bq [n_items, 64], <T[i], wf> [n_items]
  const float *McQ, *mcw;                            //   and their context maps: WQ Mc [64, 8], wf Mc [8]
  int* status;                                       // [0] set to 1 if an MMA wait timed out
  float* dbg;                                        // optional [128, 64]: activation `dbg_stage` of tile 0,
  int dbg_stage;                                     //   or (dbg_stage == -1) phase clock ticks of tile 0
  int cat_lo;                                        // > 0: full-catalog mode — candidate t is item cat_lo + t
  long long oc_user, oc_tgt;                         // strides (floats) of o_c over users / candidates ([B,T,C]: T*C, C;
                                                     //   one row per user, e.g. an expanded view or catalog mode: C, 0)
  const int *row_src, *row_seg;                      // packed profile rows (pack_rows_kernel)
  int* n_bins;                                       // [0] bins written by the packing pass, [1] tile scheduler counter
  const int* order;                                  // tiles in descending order of their user count (tile_order_kernel):
                                                     //   the scheduler hands out the longest tiles first
  int chunk_slices;                                  // work item = (tile, one of this many slices of the candidate chunks):
                                                     //   long candidate lists (full catalog) are spread over CTAs
  // DEC == 3 (split decoder): the kernel stops after the encoder and exports, per packed row, the decoder's key
  // K[row][64] and (u_0, u_1, kc_0, kc_1), and per user its segment (first packed row | length << 24) and <ctx map, wf>
  float* Kg;
  float* Ug;
  int* useg;
  float* cwsg;
  // decoder picked on the DEVICE by profile density (the host cannot know it without a sync): the entry point launches
  // the fp32-decoder kernel with auto_dec = 1 and the tcgen05-decoder kernel with auto_dec = 2; a kernel whose decoder
  // is not the one for this batch returns at once.  Dense = more than kDenseRows packed rows per user on average:
  // the fp32 loops cost ~ the user's valid positions per candidate, the tcgen05 loop is flat (all-valid profiles of
  // 50 positions: 5.9 vs 5.0 M users/s; Beauty-shaped, ~7 positions: 19.8 vs 22.6 M).
  int auto_dec;
};
constexpr int kDenseRows = 24;
__device__ __forceinline__ bool tc_auto_skip(int auto_dec, const int* n_bins, int B) {
  if (!auto_dec) return false;
  const bool dense = (long long)n_bins[0] * 64 > (long long)B * kDenseRows;
  return (auto_dec == 2) != dense;
}

struct TcSmem {
  float w[2][2 * TC_WFLOATS];                        // weight ring: [slot][hi | lo]
  float k_hi[16 * 128 * 4];
  float k_lo[16 * 128 * 4];
  float v_hi[TC_VFLOATS];
  float v_lo[TC_VFLOATS];
  float mct[8][64];                                  // folded context map, [context k][feature]
  float mcqt[8][64];                                 // the same through the decoder's WQ
  float mcw[8];                                      //   and through its ffn weight
  alignas(16) float ln[TC_LNROWS][64];                           // per block: ln1 g, b, ln2 g, b; then final g, b
  alignas(16) float dwf[64];
  float2 xch[2][2][128];                             // pair exchange: [slot][half][row]
  float uval[4][128];                                // decoder: <V_h[key], wf_h> per (head, key row of the tile)
  float kc[2][128];                                  // pair decoder: <K_h[key], context map of the key's user>, or -inf
                                                     //   for a key row that is padding (two heads)
  int oid[128];
  int ulist[128];                                    // segments of the tile: first row | length << 8
  int uuser[128];                                    //   and their users
  uint32_t kbits[2][2];                              // valid-key bits: [bin][key half]
  uint32_t headbits[4];                              // rows that start a segment
  int next_tile;                                     // dynamic tile scheduler
  uint64_t bar[2];
  uint32_t tmem_slot;
};

static_assert(sizeof(TcSmem) <= 227 * 1024, "TcSmem exceeds the 227 KB of shared memory a CTA can have");

struct TcCtx {
  TcSmem* s;
  uint32_t tmem;       // TMEM base + this warp's lane quarter
  int row, half, tid, pair_bar;
  uint32_t nwait;      // completion waits so far (selects barrier and parity)
  uint32_t ncommit;    // commits so far (meaningful on the issuing thread)
  int xslot;
  int* status;
};

struct TcTicks { long long* out; int n; bool tiles_only; };   // tiles_only (dbg_stage <= -2): tile starts of every tile of the CTA
__device__ __forceinline__ void tick(TcTicks& t, int label) {
  if (t.out && t.n < 2000 && (!t.tiles_only || label == 0)) {
    t.out[2 * t.n] = label;
    t.out[2 * t.n + 1] = clock64();
    ++t.n;
  }
}

template <int H>
struct Own {   // feature ownership of thread (row, half): v[h * N2 + c] <-> feature h*DH + half*N2 + c
  static constexpr int DH = 64 / H, N2 = DH / 2;
  __device__ static __forceinline__ int f0(int h, int half) { return h * DH + half * N2; }
};

template <int H>
__device__ __forceinline__ void ld_feat(const TcCtx& c, int col, float (&v)[32]) {
  const uint32_t b = c.tmem + col + c.half * Own<H>::N2;
  if constexpr (H == 2) umma::tmem_ld_2x16(b, b + 32, v);
  else umma::tmem_ld_4x8(b, b + 16, b + 32, b + 48, v);
}
template <int H>
__device__ __forceinline__ void st_feat(const TcCtx& c, int col, const float (&v)[32]) {
  const uint32_t b = c.tmem + col + c.half * Own<H>::N2;
  if constexpr (H == 2) {
    umma::tmem_st_x16(b, v, 0);
    umma::tmem_st_x16(b + 32, v, 16);
  } else {
#pragma unroll
    for (int h = 0; h < 4; ++h) umma::tmem_st_x8(b + 16 * h, v, 8 * h);
  }
}
// store as a tf32 operand pair: hi = the fp32 value (the tensor core reads its top 19 bits),
// lo = the remainder
template <int H>
__device__ __forceinline__ void st_operand(const TcCtx& c, int col_hi, int col_lo, const float (&v)[32]) {
  float lo[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) lo[j] = umma::tf32_lo(v[j]);
  st_feat<H>(c, col_hi, v);
  st_feat<H>(c, col_lo, lo);
}

// all TMEM / shared-memory writes of every thread become visible to the MMA issued afterwards
__device__ __forceinline__ void publish() {
  umma::tmem_st_wait();
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
}
// Completion tracking: commits alternate between two mbarriers so that a barrier is re-armed only
// after a CTA-wide sync that every thread reaches past its previous wait on it.
__device__ __forceinline__ void commit(TcCtx& c, uint32_t ahead = 0) {   // elected lane of the issuing warp
  umma::commit(&c.s->bar[(c.ncommit + ahead) & 1u]);
}
__device__ __forceinline__ void wait_mma(TcCtx& c) {
  const uint32_t n = c.nwait++;
  if (!umma::mbar_wait(&c.s->bar[n & 1u], (n >> 1) & 1u)) c.status[0] = 1;
  umma::fence_after_sync();
}
// the two threads of a row swap a pair of partial results (64-thread named barrier per warp pair)
__device__ __forceinline__ float2 pair_exchange(TcCtx& c, float2 mine) {
  float2* x = &c.s->xch[c.xslot][0][0];
  x[c.half * 128 + c.row] = mine;
  asm volatile("bar.sync %0, 64;" ::"r"(c.pair_bar) : "memory");
  const float2 o = x[(c.half ^ 1) * 128 + c.row];
  c.xslot ^= 1;
  return o;
}

// named barriers of the pipelined decoder loop: most warps only arrive, the iteration's two issuing warps wait
__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- MMA issue helpers (one thread) ---------------------------------------------------------
// D[d .. d+N) (=) A(tmem, 8*ksteps cols at a_hi / a_lo) x B(smem hi / lo)^T, 3xTF32
template <int N, int KSTEPS>
__device__ __forceinline__ void issue_3x(uint32_t d, uint32_t a_hi, uint32_t a_lo, uint32_t b_hi, uint32_t b_lo,
                                         uint32_t b_lbo) {
  constexpr uint32_t idesc = umma::idesc_tf32(N);
  const uint64_t dhi = umma::smem_desc(b_hi, b_lbo, 128), dlo = umma::smem_desc(b_lo, b_lbo, 128);
  const uint64_t step = (uint64_t)((2u * b_lbo) >> 4);
#pragma unroll
  for (int p = 0; p < 3; ++p) {
    const uint32_t a = (p == 1) ? a_lo : a_hi;
    const uint64_t b = (p == 2) ? dlo : dhi;
#pragma unroll
    for (int ks = 0; ks < KSTEPS; ++ks)
      umma::mma_tf32_ts(d, a + 8 * ks, b + ks * step, idesc, !(p == 0 && ks == 0));
  }
}
// projection with a packed weight (bias folded in as K step 8 against the ONES block)
__device__ __forceinline__ void issue_proj(uint32_t tmem, int d_col, int a_hi, int a_lo, const float* w_slot) {
  const uint32_t b_hi = umma::smem_u32(w_slot), b_lo = umma::smem_u32(w_slot + TC_WFLOATS);
  const uint32_t lbo = 64 * 16;
  issue_3x<64, 8>(tmem + d_col, tmem + a_hi, tmem + a_lo, b_hi, b_lo, lbo);
  constexpr uint32_t idesc = umma::idesc_tf32(64);
  umma::mma_tf32_ts(tmem + d_col, tmem + C_ONES, umma::smem_desc(b_hi + 16 * lbo, lbo, 128), idesc, true);
  umma::mma_tf32_ts(tmem + d_col, tmem + C_ONES, umma::smem_desc(b_lo + 16 * lbo, lbo, 128), idesc, true);
}

// ---- weight ring: cp.async global -> shared, 36,864 bytes per packed weight ---------------------
__device__ __forceinline__ void weight_prefetch(TcCtx& c, int slot, const float* __restrict__ w) {
  const uint32_t dst = umma::smem_u32(c.s->w[slot]);
#pragma unroll 1
  for (int i = c.tid; i < 2 * TC_WFLOATS / 4; i += TC_THREADS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(w + i * 4) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
template <int PENDING>
__device__ __forceinline__ void weight_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(PENDING) : "memory"); }

// LayerNorm of the row whose 64 features are split over the two threads of the pair: local two-pass
// statistics combined with Chan's formula (one exchange); gamma/beta rows in shared memory
template <int H>
__device__ __forceinline__ void layernorm_rows(TcCtx& c, float (&v)[32], const float* g, const float* b) {
  constexpr int N2 = Own<H>::N2;
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  const float ml = s * (1.0f / 32.0f);
  float m2 = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - ml;
    m2 = fmaf(d, d, m2);
  }
  const float2 o = pair_exchange(c, make_float2(s, m2));
  const float mean = (s + o.x) * (1.0f / 64.0f);
  const float dl = (s - o.x) * (1.0f / 32.0f);
  const float var = (m2 + o.y + 16.0f * dl * dl) * (1.0f / 64.0f);
  const float rstd = 1.0f / sqrtf(var + kLnEps);
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int q = 0; q < N2 / 4; ++q) {
      const int f = Own<H>::f0(h, c.half) + 4 * q;
      const float4 gg = *reinterpret_cast<const float4*>(g + f);
      const float4 bb = *reinterpret_cast<const float4*>(b + f);
      float* x = &v[h * N2 + 4 * q];
      x[0] = (x[0] - mean) * rstd * gg.x + bb.x;
      x[1] = (x[1] - mean) * rstd * gg.y + bb.y;
      x[2] = (x[2] - mean) * rstd * gg.z + bb.z;
      x[3] = (x[3] - mean) * rstd * gg.w + bb.w;
    }
}

__device__ __forceinline__ float ldg_now(const float* p) {
  float v;
  asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ int ldg_now_i(const int* p) {
  int v;
  asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ float4 ldg_now4(const float* p) {
  float4 v;
  asm volatile("ld.global.nc.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
  return v;
}

// e = mask * (Tfold[id] + Mc ctx (+ pos)) for this thread's features, in two steps so that the
// global loads of the next chunk can stay in flight behind the current chunk's MMAs
template <int H>
__device__ __forceinline__ void embed_load(const TcArgs& a, const TcCtx& c, int id, const float* __restrict__ table,
                                           const float* __restrict__ ctx, float (&v)[32], float (&cv)[8]) {
  constexpr int N2 = Own<H>::N2;
  if (id == 0) return;
  // volatile asm loads: issued HERE (the compiler would otherwise sink plain loads to their first use)
  if (ctx) {   // null: the context term comes from a per-user vector (embed_finish_user)
#pragma unroll
    for (int k = 0; k < 8; ++k) cv[k] = k < a.C ? ldg_now(ctx + k) : 0.f;
  }
  const float* t = table + (long long)id * 64;
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int q = 0; q < N2 / 4; ++q) {
      const float4 x = ldg_now4(t + Own<H>::f0(h, c.half) + 4 * q);
      float* o = &v[h * N2 + 4 * q];
      o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
    }
}
template <int H>
__device__ __forceinline__ void embed_finish(const TcArgs& a, const TcCtx& c, int id, const float (*ctab)[64],
                                             const float* __restrict__ pos_row, float (&v)[32], const float (&cv)[8]) {
  constexpr int N2 = Own<H>::N2;
  if (id == 0) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    return;
  }
  if (pos_row) {
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int q = 0; q < N2 / 4; ++q) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(pos_row + Own<H>::f0(h, c.half) + 4 * q));
        float* o = &v[h * N2 + 4 * q];
        o[0] += x.x; o[1] += x.y; o[2] += x.z; o[3] += x.w;
      }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    if (k < a.C) {
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int q = 0; q < N2 / 4; ++q) {
          const float4 m = *reinterpret_cast<const float4*>(&ctab[k][Own<H>::f0(h, c.half) + 4 * q]);
          float* o = &v[h * N2 + 4 * q];
          o[0] = fmaf(m.x, cv[k], o[0]); o[1] = fmaf(m.y, cv[k], o[1]);
          o[2] = fmaf(m.z, cv[k], o[2]); o[3] = fmaf(m.w, cv[k], o[3]);
        }
    }
  }
}
// same when every row of the tile shares one context row: add the precomputed map of that row
template <int H>
__device__ __forceinline__ void embed_finish_user(const TcCtx& c, int id, const float* cvec, float (&v)[32]) {
  constexpr int N2 = Own<H>::N2;
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int q = 0; q < N2 / 4; ++q) {
      const float4 m = *reinterpret_cast<const float4*>(cvec + Own<H>::f0(h, c.half) + 4 * q);
      float* o = &v[h * N2 + 4 * q];
      o[0] = id ? o[0] + m.x : 0.f; o[1] = id ? o[1] + m.y : 0.f;
      o[2] = id ? o[2] + m.z : 0.f; o[3] = id ? o[3] + m.w : 0.f;
    }
}
// bits lo..hi (inclusive, clipped to 0..31)
__device__ __forceinline__ uint32_t range_mask(int lo, int hi) {
  lo = max(lo, 0);
  hi = min(hi, 31);
  return hi < lo ? 0u : ((0xffffffffu >> (31 - hi)) & (0xffffffffu << lo));
}

// stage dumps: rows of the first two users, written at [user][position][feature] of dbg [2, 64, 64]
template <int H>
__device__ __forceinline__ void dump_regs(const TcArgs& a, const TcCtx& c, bool on, int stage, int ru, int rp,
                                          const float (&v)[32]) {
  if (on && a.dbg_stage == stage) {
#pragma unroll
    for (int h = 0; h < H; ++h)
#pragma unroll
      for (int q = 0; q < Own<H>::N2; ++q)
        a.dbg[(ru * 64 + rp) * 64 + Own<H>::f0(h, c.half) + q] = v[h * Own<H>::N2 + q];
  }
}
template <int H>
__device__ __forceinline__ void dump_tmem(const TcArgs& a, const TcCtx& c, bool on, int stage, int ru, int rp, int col) {
  if (a.dbg != nullptr && a.dbg_stage == stage) {   // warp-uniform: tcgen05.ld is a collective
    float v[32];
    ld_feat<H>(c, col, v);
    dump_regs<H>(a, c, on, stage, ru, rp, v);
  }
}

// K operand: this thread's features of the accumulator at `col` -> shared K-major chunks (hi/lo)
template <int H>
__device__ __forceinline__ void store_k_operand(TcCtx& c, int col) {
  constexpr int N2 = Own<H>::N2;
  float v[32];
  ld_feat<H>(c, col, v);
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int q = 0; q < N2 / 4; ++q) {
      const int kc = Own<H>::f0(h, c.half) / 4 + q;
      const float* x = &v[h * N2 + 4 * q];
      const float4 hi = make_float4(x[0], x[1], x[2], x[3]);
      const float4 lo = make_float4(umma::tf32_lo(hi.x), umma::tf32_lo(hi.y), umma::tf32_lo(hi.z), umma::tf32_lo(hi.w));
      reinterpret_cast<float4*>(c.s->k_hi)[kc * 128 + c.row] = hi;
      reinterpret_cast<float4*>(c.s->k_lo)[kc * 128 + c.row] = lo;
    }
}
// V operand: K-major over keys, [key / 4][(head, user, feature in head)][key % 4], padded chunk stride
template <int H>
__device__ __forceinline__ void store_v_operand(TcCtx& c, int col) {
  constexpr int DH = Own<H>::DH, N2 = Own<H>::N2;
  float v[32];
  ld_feat<H>(c, col, v);
  const int u = c.row / 64, j = c.row % 64;
  const int base = (j / 4) * (TC_VLBO / 4) + (j % 4);
  float* hi = c.s->v_hi + base;
  float* lo = c.s->v_lo + base;
#pragma unroll
  for (int h = 0; h < H; ++h)
#pragma unroll
    for (int q = 0; q < N2; ++q) {
      const int r = h * 2 * DH + u * DH + c.half * N2 + q;
      const float x = v[h * N2 + q];
      hi[r * 4] = x;
      lo[r * 4] = umma::tf32_lo(x);
    }
}

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

template <int W>
__device__ __forceinline__ void tmem_ld_w(uint32_t a, float (&v)[W]) {
  if constexpr (W == 8) umma::tmem_ld_1x8(a, v);
  else if constexpr (W == 16) umma::tmem_ld_1x16(a, v);
  else umma::tmem_ld_1x32(a, v);
}
template <int W>
__device__ __forceinline__ void tmem_st_w(uint32_t a, const float (&v)[W]) {
  if constexpr (W == 8) umma::tmem_st_x8(a, v, 0);
  else if constexpr (W == 16) umma::tmem_st_x16(a, v, 0);
  else umma::tmem_st_x32(a, v, 0);
}

// Masked softmax of this thread's W key columns for the two heads of a pair, result stored as the
// tf32 operand pair P.  bits: allowed keys; sc = log2(e) / sqrt(dh) (scores are scaled AFTER the
// additive mask in the reference, src/carca.py:253-254: a masked key stays at -inf either way, and a
// fully masked row is exactly 0, :256).  The two threads of a row cover 2W consecutive keys.
template <int W>
__device__ __forceinline__ void softmax_pair(TcCtx& c, uint32_t bits, float sc, uint32_t s0, uint32_t s1, uint32_t p0hi,
                                             uint32_t p0lo, uint32_t p1hi, uint32_t p1lo) {
  float v0[W], v1[W];
  tmem_ld_w<W>(s0, v0);
  tmem_ld_w<W>(s1, v1);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    const bool ok = (bits >> t) & 1u;
    v0[t] = ok ? v0[t] * sc : -INFINITY;
    v1[t] = ok ? v1[t] * sc : -INFINITY;
    m0 = fmaxf(m0, v0[t]);
    m1 = fmaxf(m1, v1[t]);
  }
  float2 o = pair_exchange(c, make_float2(m0, m1));
  m0 = fmaxf(m0, o.x);
  m1 = fmaxf(m1, o.y);
  m0 = (m0 == -INFINITY) ? 0.f : m0;
  m1 = (m1 == -INFINITY) ? 0.f : m1;
  float z0 = 0.f, z1 = 0.f;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    v0[t] = ex2_approx(v0[t] - m0);
    v1[t] = ex2_approx(v1[t] - m1);
    z0 += v0[t];
    z1 += v1[t];
  }
  o = pair_exchange(c, make_float2(z0, z1));
  z0 += o.x;
  z1 += o.y;
  const float i0 = z0 > 0.f ? 1.0f / z0 : 0.f, i1 = z1 > 0.f ? 1.0f / z1 : 0.f;
  float lo[W];
#pragma unroll
  for (int t = 0; t < W; ++t) {
    v0[t] *= i0;
    lo[t] = umma::tf32_lo(v0[t]);
  }
  tmem_st_w<W>(p0hi, v0);
  tmem_st_w<W>(p0lo, lo);
#pragma unroll
  for (int t = 0; t < W; ++t) {
    v1[t] *= i1;
    lo[t] = umma::tf32_lo(v1[t]);
  }
  tmem_st_w<W>(p1hi, v1);
  tmem_st_w<W>(p1lo, lo);
}

// Cross-attention of the decoder never needs the attention output itself, only its product with the scorer's
// weight (src/carca.py:343-345: y = ffn(attn + o), ffn = Linear(d, 1)):  <sum_j p_j V_h[j], wf_h> =
// sum_j p_j <V_h[j], wf_h> = sum_j p_j u_h[j].  So the decoder keeps one scalar u_h[j] per (head, key) and the
// softmax below returns this thread's share of  sum_h sum_j p_hj u_hj  directly — no P operand, no P.V MMA, no
// read-back of O.  Same masking / scaling rules as softmax_pair.  u0 / u1: the W keys of this thread, heads of the pair.
template <int W>
__device__ __forceinline__ float softmax_pair_dot(TcCtx& c, uint32_t bits, float sc, uint32_t s0, uint32_t s1,
                                                  const float* __restrict__ u0, const float* __restrict__ u1) {
  float v0[W], v1[W];
  tmem_ld_w<W>(s0, v0);
  tmem_ld_w<W>(s1, v1);
  float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
  for (int t = 0; t < W; ++t) {
    const bool ok = (bits >> t) & 1u;
    v0[t] = ok ? v0[t] * sc : -INFINITY;
    v1[t] = ok ? v1[t] * sc : -INFINITY;
    m0 = fmaxf(m0, v0[t]);
    m1 = fmaxf(m1, v1[t]);
  }
  float2 o = pair_exchange(c, make_float2(m0, m1));
  m0 = fmaxf(m0, o.x);
  m1 = fmaxf(m1, o.y);
  m0 = (m0 == -INFINITY) ? 0.f : m0;
  m1 = (m1 == -INFINITY) ? 0.f : m1;
  float z0 = 0.f, z1 = 0.f, d0 = 0.f, d1 = 0.f;
#pragma unroll
  for (int t = 0; t < W; t += 4) {
    const float4 a0 = *reinterpret_cast<const float4*>(u0 + t);
    const float4 a1 = *reinterpret_cast<const float4*>(u1 + t);
    const float ua[4] = {a0.x, a0.y, a0.z, a0.w}, ub[4] = {a1.x, a1.y, a1.z, a1.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float e0 = ex2_approx(v0[t + q] - m0), e1 = ex2_approx(v1[t + q] - m1);
      z0 += e0;
      z1 += e1;
      d0 = fmaf(e0, ua[q], d0);
      d1 = fmaf(e1, ub[q], d1);
    }
  }
  o = pair_exchange(c, make_float2(z0, z1));
  z0 += o.x;
  z1 += o.y;
  return (z0 > 0.f ? d0 / z0 : 0.f) + (z1 > 0.f ? d1 / z1 : 0.f);
}


// ---- cross-attention decoder, two heads, one thread per (user, candidate) row, fp32 FFMA ---------------------
// A candidate attends only to the keys of its OWN user (mean ~8 of 50 for Beauty-shaped profiles), so an M = 128
// score MMA over the bin's key window computes mostly masked products and pays a TMEM round trip, two barriers and
// an MMA completion per 128 rows.  Here a thread owns a row: its query (the folded table row TQ[id] + the context
// map, src/carca.py:238 with :85-95 folded) sits in 64 registers, the keys K_h[j] of its user are read from the
// K operand in shared memory (k_hi holds the full fp32 value; lanes of one user read the same address: broadcast),
// softmax runs online over key pairs, and the attention output is folded into u_h[j] = <V_h[j], wf_h> (see
// softmax_pair_dot).  No TMEM, no barrier, no exchange between threads; exact fp32 products.
struct DecRow {
  int id, sg, t;
  bool valid;
};
// query features of head h (32 of the 64 columns of TQ[id]); with h == 0 also the row's context and <T[id], wf>
__device__ __forceinline__ void dec_gather(const TcArgs& a, const TcSmem& s, const DecRow& r, int h, bool uctx,
                                           float (&e)[32], float (&cv)[8], float& twv) {
  if (r.id == 0) return;
  if (h == 0) {
    if (!uctx) {
      const float* ctx = a.o_c + (long long)s.uuser[r.sg] * a.oc_user + (long long)r.t * a.oc_tgt;
#pragma unroll
      for (int k = 0; k < 8; ++k) cv[k] = k < a.C ? ldg_now(ctx + k) : 0.f;
    }
    twv = ldg_now(a.tw + r.id);
  }
  const float* t = a.TQ + (long long)r.id * 64 + 32 * h;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = ldg_now4(t + 4 * i);
    e[4 * i] = x.x; e[4 * i + 1] = x.y; e[4 * i + 2] = x.z; e[4 * i + 3] = x.w;
  }
}
// residual term <o, wf> of the row (src/carca.py:343,:345) from the folded tables
__device__ __forceinline__ float dec_residual(const TcArgs& a, const TcSmem& s, const DecRow& r, bool uctx,
                                              const float* __restrict__ cws, const float (&cv)[8], float twv) {
  if (r.id == 0 || !a.residual_ca) return 0.f;
  float acc = twv;
  if (uctx) {
    acc += cws[r.sg];
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) acc = fmaf(s.mcw[k], cv[k], acc);
  }
  return acc;
}
// head h of one row:  sum_j softmax_j(<q_h, K_h[j]> / sqrt(dh)) u_h[j]  over the keys of the row's user (:340)
__device__ __forceinline__ float dec_head(const TcArgs& a, const TcSmem& s, const DecRow& r, int h, bool uctx,
                                          const float* __restrict__ cvecs, float (&e)[32], const float (&cv)[8],
                                          float sc) {
  if (r.id == 0) return 0.f;   // padded candidate: query mask 0 -> attention row exactly 0 (:256)
  const int ul = s.ulist[r.sg];
  const int row0 = ul & 0xff, ulen = ul >> 8;
  if (uctx) {
    const float4* cp = reinterpret_cast<const float4*>(cvecs + r.sg * 64 + 32 * h);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 m = cp[i];
      e[4 * i] += m.x; e[4 * i + 1] += m.y; e[4 * i + 2] += m.z; e[4 * i + 3] += m.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (k < a.C) {
        const float4* mp = reinterpret_cast<const float4*>(&s.mcqt[k][32 * h]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 m = mp[i];
          e[4 * i] = fmaf(m.x, cv[k], e[4 * i]);
          e[4 * i + 1] = fmaf(m.y, cv[k], e[4 * i + 1]);
          e[4 * i + 2] = fmaf(m.z, cv[k], e[4 * i + 2]);
          e[4 * i + 3] = fmaf(m.w, cv[k], e[4 * i + 3]);
        }
      }
  }
  // valid keys of the user's segment (a segment may hold position L-1 as a padding row)
  const uint32_t* kb = &s.kbits[row0 >> 6][0];
  unsigned long long m = (((unsigned long long)kb[1] << 32) | kb[0]) >> (row0 & 63);
  if (ulen < 64) m &= (1ull << ulen) - 1ull;
  const float4* const kh = reinterpret_cast<const float4*>(s.k_hi) + row0 + h * 8 * 128;
  const float* const uh = s.uval[h] + row0;
  float mx = -INFINITY, z = 0.f, d = 0.f;
#pragma unroll 1
  for (int j = 0; j < ulen; j += 2) {
    const int j1 = min(j + 1, ulen - 1);
    float s0a = 0.f, s0b = 0.f, s1a = 0.f, s1b = 0.f;
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      const float4 k0 = kh[kc * 128 + j], k1 = kh[kc * 128 + j1];
      const float* q = &e[4 * kc];
      s0a = fmaf(q[0], k0.x, s0a); s0b = fmaf(q[1], k0.y, s0b);
      s0a = fmaf(q[2], k0.z, s0a); s0b = fmaf(q[3], k0.w, s0b);
      s1a = fmaf(q[0], k1.x, s1a); s1b = fmaf(q[1], k1.y, s1b);
      s1a = fmaf(q[2], k1.z, s1a); s1b = fmaf(q[3], k1.w, s1b);
    }
    const bool ok0 = (m >> j) & 1ull, ok1 = (j + 1 < ulen) && ((m >> (j + 1)) & 1ull);
    const float x0 = ok0 ? (s0a + s0b) * sc : -INFINITY, x1 = ok1 ? (s1a + s1b) * sc : -INFINITY;
    const float mn = fmaxf(mx, fmaxf(x0, x1));
    const float mref = (mn == -INFINITY) ? 0.f : mn;
    const float corr = ex2_approx(mx - mref), p0 = ex2_approx(x0 - mref), p1 = ex2_approx(x1 - mref);
    z = fmaf(z, corr, p0 + p1);
    d = fmaf(d, corr, fmaf(p0, uh[j], p1 * uh[j1]));
    mx = mn;
  }
  return z > 0.f ? d / z : 0.f;
}


// ---- the same decoder over PAIRS of candidates of one user -------------------------------------------------------
// The per-row loop above is bound by shared-memory operand delivery, not by FFMA issue: every FFMA takes one K value
// through an LDS.128, and a 128-bit shared load costs 4 wavefronts per warp even when all lanes read the same address
// (ncu: 4.0 wavefronts per LDS of dec_head).  Here a thread owns candidates 2p and 2p+1 of a user, so every K value it
// loads feeds two FFMAs.  With one context row per user the context part of the query is the same for all of the
// user's candidates and is folded into kc_h[j] = <K_h[j], context map> once per tile (kc = -inf marks a padding key):
// the query is the folded table row itself.
struct DecPair {
  int id0, id1, sg, t0;   // candidates t0 and t0 + 1 of segment sg; t0 < 0: no such pair; id1 = 0 if t0 + 1 == T
};
// head h of both rows: e[0..32) = row 0, e[32..64) = row 1
__device__ __forceinline__ void pair_gather(const TcArgs& a, const DecPair& r, int h, float (&e)[64]) {
  const int ids[2] = {r.id0, r.id1};
#pragma unroll
  for (int k = 0; k < 2; ++k) {
    if (ids[k] == 0) continue;
    const float* t = a.TQ + (long long)ids[k] * 64 + 32 * h;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 x = ldg_now4(t + 4 * i);
      float* o = &e[32 * k + 4 * i];
      o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
    }
  }
}
// per-candidate context (not one row per user): q += McQ ctx for head h of row k of the pair
__device__ __forceinline__ void pair_ctx_add(const TcArgs& a, const TcSmem& s, int h, int k, const float (&cv)[16],
                                             float (&e)[64]) {
#pragma unroll
  for (int kk = 0; kk < 8; ++kk)
    if (kk < a.C) {
      const float4* mp = reinterpret_cast<const float4*>(&s.mcqt[kk][32 * h]);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float4 m = mp[i];
        float* o = &e[32 * k + 4 * i];
        o[0] = fmaf(m.x, cv[8 * k + kk], o[0]); o[1] = fmaf(m.y, cv[8 * k + kk], o[1]);
        o[2] = fmaf(m.z, cv[8 * k + kk], o[2]); o[3] = fmaf(m.w, cv[8 * k + kk], o[3]);
      }
    }
}
// head h: att[k] = sum_j softmax_j((<q_k, K_h[j]> + kc_h[j]) / sqrt(dh)) u_h[j] for both rows (src/carca.py:340)
__device__ __forceinline__ void pair_head(const TcSmem& s, const DecPair& r, int h, const float (&e)[64], float sc,
                                          float& att0, float& att1) {
  const int ul = s.ulist[r.sg];
  const int row0 = ul & 0xff, ulen = ul >> 8;
  const float4* const kh = reinterpret_cast<const float4*>(s.k_hi) + row0 + h * 8 * 128;
  const float* const uh = s.uval[h] + row0;
  const float* const ch = s.kc[h] + row0;
  float mxa = -INFINITY, za = 0.f, da = 0.f, mxb = -INFINITY, zb = 0.f, db = 0.f;
#pragma unroll 1
  for (int j = 0; j < ulen; j += 2) {
    const int j1 = min(j + 1, ulen - 1);
    float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;   // a: row 0, b: row 1; 0 / 1: key j / j1
#pragma unroll
    for (int kc = 0; kc < 8; ++kc) {
      const float4 k0 = kh[kc * 128 + j], k1 = kh[kc * 128 + j1];
      const float* qa = &e[4 * kc];
      const float* qb = &e[32 + 4 * kc];
      a0 = fmaf(qa[0], k0.x, a0); a1 = fmaf(qa[0], k1.x, a1); b0 = fmaf(qb[0], k0.x, b0); b1 = fmaf(qb[0], k1.x, b1);
      a0 = fmaf(qa[1], k0.y, a0); a1 = fmaf(qa[1], k1.y, a1); b0 = fmaf(qb[1], k0.y, b0); b1 = fmaf(qb[1], k1.y, b1);
      a0 = fmaf(qa[2], k0.z, a0); a1 = fmaf(qa[2], k1.z, a1); b0 = fmaf(qb[2], k0.z, b0); b1 = fmaf(qb[2], k1.z, b1);
      a0 = fmaf(qa[3], k0.w, a0); a1 = fmaf(qa[3], k1.w, a1); b0 = fmaf(qb[3], k0.w, b0); b1 = fmaf(qb[3], k1.w, b1);
    }
    const float c0 = ch[j], c1 = (j + 1 < ulen) ? ch[j1] : -INFINITY;
    const float u0 = uh[j], u1 = uh[j1];
    {
      const float x0 = (a0 + c0) * sc, x1 = (a1 + c1) * sc;
      const float mn = fmaxf(mxa, fmaxf(x0, x1));
      const float mref = (mn == -INFINITY) ? 0.f : mn;
      const float corr = ex2_approx(mxa - mref), p0 = ex2_approx(x0 - mref), p1 = ex2_approx(x1 - mref);
      za = fmaf(za, corr, p0 + p1);
      da = fmaf(da, corr, fmaf(p0, u0, p1 * u1));
      mxa = mn;
    }
    {
      const float x0 = (b0 + c0) * sc, x1 = (b1 + c1) * sc;
      const float mn = fmaxf(mxb, fmaxf(x0, x1));
      const float mref = (mn == -INFINITY) ? 0.f : mn;
      const float corr = ex2_approx(mxb - mref), p0 = ex2_approx(x0 - mref), p1 = ex2_approx(x1 - mref);
      zb = fmaf(zb, corr, p0 + p1);
      db = fmaf(db, corr, fmaf(p0, u0, p1 * u1));
      mxb = mn;
    }
  }
  att0 = za > 0.f ? da / za : 0.f;
  att1 = zb > 0.f ? db / zb : 0.f;
}

// DEC selects the two-head cross-attention decoder: 3 = none (keys exported for decode_pairs_kernel below), 2 = fp32
// loop over candidate PAIRS, 1 = fp32 loop with one row per thread, 0 = the tcgen05 decoder loops below (also used by
// the dot decoder and by four heads).
// A template parameter, so that each kernel is register-allocated for the one decoder it contains.
template <int H, int DEC>
__global__ void __launch_bounds__(TC_THREADS, 1) fused_eval_tc_kernel(const TcArgs a) {
  static_assert(DEC == 0 || H == 2, "the row / pair decoders are written for two heads");
  constexpr bool ROW_DEC = DEC != 0;
  if (tc_auto_skip(a.auto_dec, a.n_bins, a.B)) return;
  constexpr int DH = Own<H>::DH, N2 = Own<H>::N2;
  CARCA_DYN_SMEM(unsigned char, raw);
  TcSmem& s = *reinterpret_cast<TcSmem*>(raw);
  TcCtx c;
  c.s = &s;
  c.tid = threadIdx.x;
  const int w = c.tid / 32;
  c.row = 32 * (w % 4) + (c.tid % 32);
  c.half = w / 4;
  c.pair_bar = 1 + (w % 4);
  c.nwait = 0;
  c.ncommit = 0;
  c.xslot = 0;
  c.status = a.status;
  const int L = a.L;
  const float sc = 1.4426950408889634f / sqrtf((float)DH);
  // MMAs are issued by one elected lane of warp 0; the branch on the warp index is warp-uniform so
  // descriptors stay in uniform registers (a divergent `tid == 0` branch costs a waterfall loop per MMA)
  // Two issuers (warps 0 and 1): independent accumulators of a phase (the two heads of a pair, Q and K, ...)
  // are issued concurrently — small MMAs are bound by the ~70-cycle issue path of one thread, not by the
  // tensor pipe.  Both issuers commit in every phase, so every mbarrier phase takes two arrivals.
  const int wu = __shfl_sync(kFull, w, 0);
  const int iw = wu < 2 ? wu : -1;

  if (w == 0) umma::tmem_alloc(&s.tmem_slot, 512);
  if (c.tid == 0) {
    umma::mbar_init(&s.bar[0], 2);
    umma::mbar_init(&s.bar[1], 2);
  }
  for (int i = c.tid; i < 64 * 8; i += TC_THREADS) s.mct[i % 8][i / 8] = a.Mc[i];
  for (int i = c.tid; i < (4 * a.n_blocks + 2) * 64; i += TC_THREADS) {
    const int r = i / 64, f = i % 64;
    const float* src;
    if (r < 4 * a.n_blocks) {
      const TcBlockW& wb = a.blk[r / 4];
      src = (r % 4 == 0) ? wb.ln1_g : (r % 4 == 1) ? wb.ln1_b : (r % 4 == 2) ? wb.ln2_g : wb.ln2_b;
    } else {
      src = (r == 4 * a.n_blocks) ? a.fn_g : a.fn_b;
    }
    s.ln[r][f] = src[f];
  }
  if (c.tid < 64) s.dwf[c.tid] = a.decoder == 1 ? a.dwf[c.tid] : 0.f;
  if (a.decoder == 1) {
    for (int i = c.tid; i < 64 * 8; i += TC_THREADS) s.mcqt[i % 8][i / 8] = a.McQ[i];
    if (c.tid < 8) s.mcw[c.tid] = a.mcw[c.tid];
  }
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = s.tmem_slot;                 // lane 0 (MMA operand / accumulator addresses)
  c.tmem = tmem0 + ((uint32_t)(32 * (w % 4)) << 16);  // this warp's lane quarter

  TcTicks tk;
  tk.out = (a.dbg && (a.dbg_stage == -1 || a.dbg_stage <= -2) && blockIdx.x == (a.dbg_stage <= -2 ? -2 - a.dbg_stage : 0) &&
            c.tid == 0) ? reinterpret_cast<long long*>(a.dbg) : nullptr;   // dbg_stage -2 - k: tile starts of CTA k
  tk.n = 0;
  tk.tiles_only = a.dbg_stage <= -2;
  const int n_slices = max(1, a.chunk_slices);
  // Tail slicing: with the tiles in longest-first order the last, partly filled wave (n_tiles mod gridDim tiles) holds
  // the cheapest tiles while the other CTAs idle; those tiles become S work items each (every item re-encodes the tile,
  // which only costs idle CTAs their time, and scores 1/S of its candidates), S as large as keeps them in one wave.
  const int n_tiles_real = (a.n_bins[0] + 1) / 2;
  int tail = 0, tail_s = 1;
  if (n_slices == 1 && DEC != 3 && n_tiles_real > 0) {   // (fewer tiles than CTAs: every tile is a tail tile)
    tail = n_tiles_real % (int)gridDim.x;
    tail_s = tail > 0 ? min(4, (int)gridDim.x / tail) : 1;
    if (tail_s < 2) { tail = 0; tail_s = 1; }
  }
  const int n_full = n_tiles_real - tail;
  const int n_tiles = n_slices > 1 ? n_tiles_real * n_slices : n_full + tail * tail_s;
                                                            // work items: every slice re-encodes its tile (cheap next
                                                            // to >= 16 candidate chunks per user) and scores its share
  const int u = c.row / 64, i = c.row % 64;   // bin (64-row half of the tile) and row within it
  float* const plast = s.k_hi;                // dot decoder: last-position vectors per segment (K is unused there)
  // tiles cost 62K cycles + 7K per user they hold: CTAs take the next tile from a global counter instead of
  // a fixed stride, which evens out the last wave (a.n_bins[1], zeroed by the host before the launch)
  int tile = blockIdx.x;
#pragma unroll 1
  for (bool first = true;; first = false) {
    __syncthreads();   // every thread is done with the previous tile (and has read s.next_tile)
    if (!first) tile = s.next_tile;
    if (tile >= n_tiles) break;
    if (c.half == 0) {   // constant [1,0,0,0,0,0,0,0] column block: the A operand of every bias step (per tile: the
      float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // pipelined decoder reuses these columns for scores)
      umma::tmem_st8(c.tmem + C_ONES, ones);
    }
    int tile_idx, slice, ns;   // the item's tile, its slice of the tile's candidates, slices of that tile
    if (n_slices > 1) {
      tile_idx = tile / n_slices; slice = tile % n_slices; ns = n_slices;
    } else if (tile < n_full) {
      tile_idx = tile; slice = 0; ns = 1;
    } else {
      tile_idx = n_full + (tile - n_full) / tail_s; slice = (tile - n_full) % tail_s; ns = tail_s;
    }
    if (a.order) tile_idx = a.order[tile_idx];
    const long long trow0 = (long long)tile_idx * 128;
    tick(tk, 0);
    // ---- packed rows (csrc/fused_eval_tc.cuh: pack_rows_kernel): row -> (user, position), segment of the user
    // (a tile is two bins: with an odd bin count the last tile's second bin does not exist)
    const int src = (trow0 + c.row) / 64 < a.n_bins[0] ? a.row_src[trow0 + c.row] : -1;
    const int seg = src >= 0 ? a.row_seg[trow0 + c.row] : 0;
    const int ru = src >> 8, rp = src & 255;
    const int seg0 = seg & 0xff, seglen = (seg >> 8) & 0xff;
    const int my_pid = src >= 0 ? a.p_x[(long long)ru * L + rp] : 0;
    const bool head = src >= 0 && i == seg0;
    const bool dbg_row = a.dbg != nullptr && a.dbg_stage > 0 && src >= 0 && ru < 2 && rp < 64;
    if (c.half == 0) {
      const uint32_t bits = __ballot_sync(kFull, my_pid != 0);
      const uint32_t hb = __ballot_sync(kFull, head);
      if ((c.tid & 31) == 0) {
        s.kbits[u][w & 1] = bits;
        s.headbits[w] = hb;
      }
    }
    weight_prefetch(c, 0, a.blk[0].wq);
    weight_prefetch(c, 1, a.blk[0].wk);
    __syncthreads();
    if (c.tid == 0) s.next_tile = (int)gridDim.x + atomicAdd(a.n_bins + 1, 1);   // read at the next loop top
    // segment list of the tile (one entry per user): index = number of segment heads before this one
    const int head_row = 64 * u + seg0;
    int seg_idx = __popc(s.headbits[head_row >> 5] & ((1u << (head_row & 31)) - 1u));
    for (int q = 0; q < (head_row >> 5); ++q) seg_idx += __popc(s.headbits[q]);
    const int n_seg = __popc(s.headbits[0]) + __popc(s.headbits[1]) + __popc(s.headbits[2]) + __popc(s.headbits[3]);
    if (c.half == 0 && head) {
      s.ulist[seg_idx] = c.row | (seglen << 8);
      s.uuser[seg_idx] = ru;
    }
    // allowed keys of this thread's key half in self-attention: own segment, causal (j <= i), valid
    // key, valid query
    const uint32_t self_bits = my_pid != 0 ? (s.kbits[u][c.half] & range_mask(seg0 - 32 * c.half, i - 32 * c.half)) : 0u;
    float v[32];   // the activation this thread carries from phase to phase
    {              // profile embedding (src/carca.py:415)
      float cv[8];
      embed_load<H>(a, c, my_pid, a.Tfold, a.p_c + ((long long)ru * L + rp) * a.C, v, cv);
      embed_finish<H>(a, c, my_pid, s.mct, a.pos ? a.pos + (long long)rp * 64 : nullptr, v, cv);
    }
    tick(tk, 1);

#pragma unroll 1
    for (int b = 0; b < a.n_blocks; ++b) {
      const TcBlockW& wb = a.blk[b];
      st_operand<H>(c, C_XHI, C_XLO, v);                        // block input: operand of K and V
      layernorm_rows<H>(c, v, s.ln[4 * b], s.ln[4 * b + 1]);    // LN1 (:298); v = qn from here on
      st_operand<H>(c, C_QNHI, C_QNLO, v);
      weight_wait<0>();
      publish();
      tick(tk, 2);
      dump_regs<H>(a, c, dbg_row, 1 + 10 * b, ru, rp, v);
      // Q from LN1(x), K from raw x (:238-239), separate completion events
      if (iw >= 0) {
        if (umma::elect_one()) {
          if (iw == 0) issue_proj(tmem0, C_ACCQ, C_QNHI, C_QNLO, s.w[0]);
          else issue_proj(tmem0, C_ACCK, C_XHI, C_XLO, s.w[1]);
          commit(c);
        }
      }
      c.ncommit++;
      wait_mma(c);                                              // Q and K done: both weight slots are free
      weight_prefetch(c, 0, wb.wv);
      weight_prefetch(c, 1, wb.w1);
      {   // tf32 remainder of Q (the A operand of the score products) -> QN_LO
        float q[32];
        ld_feat<H>(c, C_ACCQ, q);
#pragma unroll
        for (int j = 0; j < 32; ++j) q[j] = umma::tf32_lo(q[j]);
        st_feat<H>(c, C_QNLO, q);
      }
      tick(tk, 3);
      store_k_operand<H>(c, C_ACCK);
      weight_wait<1>();                                         // WV landed (W1 may still be in flight)
      publish();
      tick(tk, 4);
      if (iw >= 0) {
        if (umma::elect_one()) {
          if (iw == 0) issue_proj(tmem0, C_ACCV, C_XHI, C_XLO, s.w[0]);      // V (:240)
          commit(c);
        }
      }
      c.ncommit++;
      wait_mma(c);
      tick(tk, 5);
      dump_tmem<H>(a, c, dbg_row, 2 + 10 * b, ru, rp, C_ACCQ);
      dump_tmem<H>(a, c, dbg_row, 3 + 10 * b, ru, rp, C_ACCK);
      dump_tmem<H>(a, c, dbg_row, 4 + 10 * b, ru, rp, C_ACCV);
      store_v_operand<H>(c, C_ACCV);
      weight_prefetch(c, 0, wb.w2);
      if (!a.residual_sa) {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
      publish();
      tick(tk, 6);
      // ---- causal self-attention (:299), two heads per round; v accumulates qn + O (:302)
#pragma unroll
      for (int hp = 0; hp < H; hp += 2) {
        if (iw >= 0) {   // one head per issuer
          if (umma::elect_one()) {
            const int e = iw, h = hp + e;
            const uint32_t koff = (uint32_t)(h * DH / 4) * 2048u;
            issue_3x<128, DH / 8>(tmem0 + (e ? C_ACCK : C_XHI), tmem0 + C_ACCQ + h * DH, tmem0 + C_QNLO + h * DH,
                                  umma::smem_u32(s.k_hi) + koff, umma::smem_u32(s.k_lo) + koff, 2048u);
            commit(c);
          }
        }
        c.ncommit++;
        wait_mma(c);
        tick(tk, 20);
        {
          const uint32_t r0 = c.tmem + C_XHI, r1 = c.tmem + C_ACCK, k = 32 * c.half;
          softmax_pair<32>(c, self_bits, sc, r0 + 64 * u + k, r1 + 64 * u + k, r0 + k, r0 + 64 + k, r1 + k, r1 + 64 + k);
        }
        publish();
        tick(tk, 21);
        if (iw >= 0) {   // O_h = P_h V_h for both bins' V at once (each row keeps its own bin's columns)
          if (umma::elect_one()) {
            const int e = iw, h = hp + e;
            const uint32_t p = tmem0 + (e ? C_ACCK : C_XHI);
            const uint32_t voff = (uint32_t)(h * 2 * DH) * 16u;
            issue_3x<2 * DH, 8>(tmem0 + C_QNHI + e * 2 * DH, p, p + 64, umma::smem_u32(s.v_hi) + voff,
                                umma::smem_u32(s.v_lo) + voff, (uint32_t)TC_VLBO);
            commit(c);
          }
        }
        c.ncommit++;
        wait_mma(c);
        tick(tk, 22);
        {
          const uint32_t o0 = c.tmem + C_QNHI + u * DH + c.half * N2;
          if constexpr (H == 2) {
            float o[32];
            umma::tmem_ld_2x16(o0, o0 + 2 * DH, o);
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] += o[j];
          } else {
            float o[16];
            umma::tmem_ld_2x8(o0, o0 + 2 * DH, o);
#pragma unroll
            for (int j = 0; j < 16; ++j) v[hp * N2 + j] += o[j];
          }
        }
        if (hp + 2 < H) {   // the next round's PV overwrites O: every thread must have read it
          umma::fence_before_sync();
          __syncthreads();
          umma::fence_after_sync();
        }
      }
      dump_regs<H>(a, c, dbg_row, 5 + 10 * b, ru, rp, v);
      layernorm_rows<H>(c, v, s.ln[4 * b + 2], s.ln[4 * b + 3]);   // LN2 (:304); v = s2 from here on
      st_operand<H>(c, C_XHI, C_XLO, v);
      weight_wait<0>();
      publish();
      tick(tk, 7);
      if (iw >= 0) {
        if (umma::elect_one()) {
          if (iw == 0) issue_proj(tmem0, C_ACCQ, C_XHI, C_XLO, s.w[1]);     // ffn_1 (:307)
          commit(c);
        }
      }
      c.ncommit++;
      wait_mma(c);
      tick(tk, 8);
      {
        float f[32];
        ld_feat<H>(c, C_ACCQ, f);
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = f[j] > 0.f ? f[j] : kLeakySlope * f[j];   // LeakyReLU (:308)
        st_operand<H>(c, C_QNHI, C_QNLO, f);
      }
      publish();
      tick(tk, 9);
      if (iw >= 0) {
        if (umma::elect_one()) {
          if (iw == 0) issue_proj(tmem0, C_ACCK, C_QNHI, C_QNLO, s.w[0]);   // ffn_2 (:311)
          commit(c);
        }
      }
      c.ncommit++;
      wait_mma(c);
      tick(tk, 10);
      // next weights: the following block's WQ/WK, or the decoder's WK/WV
      if (b + 1 < a.n_blocks) {
        weight_prefetch(c, 0, a.blk[b + 1].wq);
        weight_prefetch(c, 1, a.blk[b + 1].wk);
      } else if (a.decoder == 1) {
        weight_prefetch(c, 0, a.dwk);
        weight_prefetch(c, 1, a.dwv);
      }
      {   // block output (+ LN2 residual, :316): the next block's input
        float f[32];
        ld_feat<H>(c, C_ACCK, f);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = a.residual_sa ? v[j] + f[j] : f[j];
      }
      dump_regs<H>(a, c, dbg_row, 9 + 10 * b, ru, rp, v);
      tick(tk, 11);
    }

    layernorm_rows<H>(c, v, s.ln[4 * a.n_blocks], s.ln[4 * a.n_blocks + 1]);   // final LayerNorm (:421)
    dump_regs<H>(a, c, dbg_row, 100, ru, rp, v);
    if (a.decoder == 1) {   // keys / values of the encoded profile (:239-240 with p as key and value)
      st_operand<H>(c, C_QNHI, C_QNLO, v);
      weight_wait<0>();
      publish();
      tick(tk, 12);
      if (iw >= 0) {
        if (umma::elect_one()) {
          if (iw == 0) issue_proj(tmem0, C_ACCK, C_QNHI, C_QNLO, s.w[0]);
          else issue_proj(tmem0, C_ACCV, C_QNHI, C_QNLO, s.w[1]);
          commit(c);
        }
      }
      c.ncommit++;
    }
    if (a.decoder == 1) {
      wait_mma(c);
      store_k_operand<H>(c, C_ACCK);
      {   // u_h[row] = <V_h[row], wf_h> (see softmax_pair_dot): the decoder needs nothing else of V
        float vv[32];
        ld_feat<H>(c, C_ACCV, vv);
        float part[H];
#pragma unroll
        for (int h = 0; h < H; ++h) {
          part[h] = 0.f;
#pragma unroll
          for (int qq = 0; qq < N2; ++qq) part[h] = fmaf(vv[h * N2 + qq], s.dwf[Own<H>::f0(h, c.half) + qq], part[h]);
        }
#pragma unroll
        for (int h = 0; h < H; h += 2) {
          const float2 o = pair_exchange(c, make_float2(part[h], part[h + 1]));
          if (c.half == 0) {
            s.uval[h][c.row] = part[h] + o.x;
            s.uval[h + 1][c.row] = part[h + 1] + o.y;
          }
        }
      }
      tick(tk, 13);
    } else if (src >= 0 && rp == L - 1) {   // dot decoder: only the last profile position is used (:362)
#pragma unroll
      for (int h = 0; h < H; ++h)
#pragma unroll
        for (int q = 0; q < N2; ++q) plast[seg_idx * 64 + Own<H>::f0(h, c.half) + q] = v[h * N2 + q];
    }

    // ---- decoder.  The candidates of a bin's users form one stream of (segment, candidate) rows; an iteration
    // takes the next 128 rows of it (T = 101 would otherwise leave 27 of 128 MMA rows idle) against the 64 keys
    // of the bin, every row masked to its own user's key segment.  Software pipeline over iterations q:
    // candidate ids of q+2 are loaded into a register, ids of q+1 sit in s.oid, and the table rows of q+1 are
    // gathered into registers while the MMAs of q run.  cross-attention: the candidate's query comes from the
    // folded table TQ (= WQ e + bq, carca_eval_prepare), so an iteration is  Q -> TMEM, scores MMA, softmax,
    // PV MMA, <O, wf> + <e, wf> + bf  (src/carca.py:338-347).
    const bool ca = a.decoder == 1;
    const float* const tab = ca ? a.TQ : a.Tfold;
    const float(*const ctab)[64] = ca ? s.mcqt : s.mct;
    // one context row per user (expanded [B,T,C] view / catalog mode): its map through the context table is
    // computed once per segment into the (now idle) weight ring instead of per candidate row
    const bool uctx = a.oc_tgt == 0;
    float* const cvecs = s.w[0];              // [n_seg][64]
    float* const cws = s.w[0] + 128 * 64;     // [n_seg]: the same through the decoder's ffn weight
    const int nb0 = __popc(s.headbits[0]) + __popc(s.headbits[1]);
    auto seg_base = [&](int bin) { return bin ? nb0 : 0; };            // segments are listed bin by bin
    auto seg_cnt = [&](int bin) { return bin ? n_seg - nb0 : nb0; };
    if (uctx) {
      for (int idx = c.tid; idx < n_seg * 64; idx += TC_THREADS) {
        const float* cu = a.o_c + (long long)s.uuser[idx >> 6] * a.oc_user;
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < a.C) acc = fmaf(ctab[k][idx & 63], __ldg(cu + k), acc);
        cvecs[idx] = acc;
      }
      if (ca) {
        for (int sg = c.tid; sg < n_seg; sg += TC_THREADS) {
          const float* cu = a.o_c + (long long)s.uuser[sg] * a.oc_user;
          float acc = 0.f;
#pragma unroll
          for (int k = 0; k < 8; ++k)
            if (k < a.C) acc = fmaf(s.mcw[k], __ldg(cu + k), acc);
          cws[sg] = acc;
        }
      }
    }
    if (DEC >= 2 && ca) {   // fp32 decoder over candidate pairs (pair_head above), or export for the split decoder
      __syncthreads();      // cvecs / K / uval of the tile are visible
      {   // kc_h[key row] = <K_h[row], context map of the row's user> (thread (row, half): head = half); -inf: padding key
        float kcv = (src >= 0 && my_pid != 0) ? 0.f : -INFINITY;
        if (uctx && kcv == 0.f) {
          const float4* kp = reinterpret_cast<const float4*>(s.k_hi) + c.half * 8 * 128 + c.row;
          const float4* cp = reinterpret_cast<const float4*>(cvecs + seg_idx * 64 + 32 * c.half);
          float p0 = 0.f, p1 = 0.f;
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            const float4 k = kp[kc * 128], m = cp[kc];
            p0 = fmaf(k.x, m.x, p0); p1 = fmaf(k.y, m.y, p1);
            p0 = fmaf(k.z, m.z, p0); p1 = fmaf(k.w, m.w, p1);
          }
          kcv = p0 + p1;
        }
        s.kc[c.half][c.row] = kcv;
      }
      __syncthreads();
      tick(tk, 40);
      if (DEC == 3) {   // export keys and per-key / per-user terms; decode_pairs_kernel scores the candidates
        const long long grow = trow0 + c.row;
        float4* kg = reinterpret_cast<float4*>(a.Kg + grow * 64 + c.half * 32);
        const float4* kp = reinterpret_cast<const float4*>(s.k_hi) + c.half * 8 * 128 + c.row;
#pragma unroll
        for (int kc = 0; kc < 8; ++kc) kg[kc] = kp[kc * 128];
        if (c.half == 0) {
          reinterpret_cast<float4*>(a.Ug)[grow] = make_float4(s.uval[0][c.row], s.uval[1][c.row], s.kc[0][c.row], s.kc[1][c.row]);
          if (head) {
            a.useg[ru] = (int)grow | (seglen << 24);
            a.cwsg[ru] = cws[seg_idx];
          }
        }
        if (!tk.tiles_only) tk.out = nullptr;
        continue;
      }
      const int P = (a.T + 1) / 2;      // candidate pairs per user
      const int total = n_seg * P;      // (segment, pair) work items of the tile
      const int n_it = (total + TC_THREADS - 1) / TC_THREADS;
      const int per = (n_it + ns - 1) / ns;
      const int it_lo = min(n_it, slice * per), it_hi = min(n_it, it_lo + per);
      const float bfv = __ldg(a.dbf);
      auto pair_at = [&](int it) {
        DecPair r;
        const int f = it * TC_THREADS + c.tid;
        r.id0 = r.id1 = 0;
        r.sg = 0;
        r.t0 = -1;
        if (it < it_hi && f < total) {
          r.sg = f / P;
          r.t0 = 2 * (f - r.sg * P);
          const bool two = r.t0 + 1 < a.T;
          if (a.cat_lo > 0) {
            r.id0 = a.cat_lo + r.t0;
            r.id1 = two ? r.id0 + 1 : 0;
          } else {
            const int* px = a.o_x + (long long)s.uuser[r.sg] * a.T + r.t0;
            r.id0 = ldg_now_i(px);
            r.id1 = two ? ldg_now_i(px + 1) : 0;
          }
        }
        return r;
      };
      // Two 64-register query buffers: A = head 0 of both rows, B = head 1.  Head 0 of the NEXT pair is gathered into A
      // as soon as this pair's head 0 is done (covered by its head 1), head 1 into B after head 1 (covered by the next
      // pair's head 0).
      DecPair cur = pair_at(it_lo), nxt = pair_at(it_lo + 1);
      float qA[64], qB[64];
      pair_gather(a, cur, 0, qA);
      pair_gather(a, cur, 1, qB);
#pragma unroll 1
      for (int it = it_lo; it < it_hi; ++it) {
        const DecPair nn = pair_at(it + 2);
        float res0 = bfv, res1 = bfv, cv[16];
        if (cur.t0 >= 0) {
          if (uctx) {
            if (a.residual_ca) {
              const float cw = cws[cur.sg];
              if (cur.id0 != 0) res0 += ldg_now(a.tw + cur.id0) + cw;
              if (cur.id1 != 0) res1 += ldg_now(a.tw + cur.id1) + cw;
            }
          } else {   // per-candidate context: read here (not prefetched) and folded into the queries
            const float* ctx = a.o_c + (long long)s.uuser[cur.sg] * a.oc_user + (long long)cur.t0 * a.oc_tgt;
            float r0 = 0.f, r1 = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) {
              cv[k] = (k < a.C && cur.id0 != 0) ? __ldg(ctx + k) : 0.f;
              cv[8 + k] = (k < a.C && cur.id1 != 0) ? __ldg(ctx + a.oc_tgt + k) : 0.f;
              r0 = fmaf(s.mcw[k], cv[k], r0);
              r1 = fmaf(s.mcw[k], cv[8 + k], r1);
            }
            if (a.residual_ca) {
              if (cur.id0 != 0) res0 += ldg_now(a.tw + cur.id0) + r0;
              if (cur.id1 != 0) res1 += ldg_now(a.tw + cur.id1) + r1;
            }
            if (cur.id0 != 0) pair_ctx_add(a, s, 0, 0, cv, qA);
            if (cur.id1 != 0) pair_ctx_add(a, s, 0, 1, cv, qA);
          }
        }
        tick(tk, 41);
        float h0a = 0.f, h0b = 0.f, h1a = 0.f, h1b = 0.f;
        if (cur.id0 != 0 || cur.id1 != 0) pair_head(s, cur, 0, qA, sc, h0a, h0b);
        pair_gather(a, nxt, 0, qA);
        if (cur.id0 != 0 || cur.id1 != 0) {
          if (!uctx) {
            if (cur.id0 != 0) pair_ctx_add(a, s, 1, 0, cv, qB);
            if (cur.id1 != 0) pair_ctx_add(a, s, 1, 1, cv, qB);
          }
          pair_head(s, cur, 1, qB, sc, h1a, h1b);
        }
        pair_gather(a, nxt, 1, qB);
        if (cur.t0 >= 0) {   // a padded candidate scores sigmoid(bf): query mask 0 -> attention row 0, o = 0 (:94, :256)
          float* yp = a.y + (long long)s.uuser[cur.sg] * a.ldy + a.col0 + cur.t0;
          yp[0] = 1.0f / (1.0f + expf(-(res0 + (cur.id0 != 0 ? h0a + h1a : 0.f))));
          if (cur.t0 + 1 < a.T) yp[1] = 1.0f / (1.0f + expf(-(res1 + (cur.id1 != 0 ? h0b + h1b : 0.f))));
        }
        tick(tk, 42);
        cur = nxt;
        nxt = nn;
      }
      if (!tk.tiles_only) tk.out = nullptr;   // first tile only
      continue;
    }
    if (DEC == 1 && ca) {   // per-row fp32 decoder (dec_head above); rows of iteration it+1 are
      __syncthreads();                  // gathered while iteration it computes, ids are fetched two iterations ahead
      tick(tk, 40);
      const int total = n_seg * a.T;    // (segment, candidate) rows of the tile
      const int n_it = (total + TC_THREADS - 1) / TC_THREADS;
      const int per = (n_it + ns - 1) / ns;
      const int it_lo = min(n_it, slice * per), it_hi = min(n_it, it_lo + per);
      const float bfv = __ldg(a.dbf);
      auto row_at = [&](int it) {
        DecRow r;
        const int f = it * TC_THREADS + c.tid;
        r.valid = it < it_hi && f < total;
        r.sg = r.valid ? f / a.T : 0;
        r.t = r.valid ? f - r.sg * a.T : 0;
        r.id = 0;
        if (r.valid) r.id = a.cat_lo > 0 ? a.cat_lo + r.t : ldg_now_i(a.o_x + (long long)s.uuser[r.sg] * a.T + r.t);
        return r;
      };
      // the query halves (one per head) are prefetched in two stages: head 0 of the next row while this row's head 0
      // runs, head 1 of the next row while its head 1 (and the next row's head 0) run — 96 live query registers
      DecRow cur = row_at(it_lo), nxt = row_at(it_lo + 1);
      float qa[32], qb[32], qc[32], cvc[8], cvn[8], twc = 0.f, twn = 0.f;
      auto step = [&](int it, float(&c0)[32], float(&c1)[32], float(&n0)[32]) {
        const DecRow nn = row_at(it + 2);
        dec_gather(a, s, nxt, 0, uctx, n0, cvn, twn);
        tick(tk, 41);
        float acc = dec_residual(a, s, cur, uctx, cws, cvc, twc) + bfv;
        acc += dec_head(a, s, cur, 0, uctx, cvecs, c0, cvc, sc);
        dec_gather(a, s, nxt, 1, uctx, c0, cvn, twn);   // c0 is free now: it receives head 1 of the next row
        acc += dec_head(a, s, cur, 1, uctx, cvecs, c1, cvc, sc);
        if (cur.valid) a.y[(long long)s.uuser[cur.sg] * a.ldy + a.col0 + cur.t] = 1.0f / (1.0f + expf(-acc));
        tick(tk, 42);
        cur = nxt;
        nxt = nn;
#pragma unroll
        for (int k = 0; k < 8; ++k) cvc[k] = cvn[k];
        twc = twn;
      };
      // buffer rotation over three 32-float arrays: (head 0, head 1, next head 0) = (A, B, C) -> (C, A, B) -> (B, C, A)
      dec_gather(a, s, cur, 0, uctx, qa, cvc, twc);
      dec_gather(a, s, cur, 1, uctx, qb, cvc, twc);
#pragma unroll 1
      for (int it = it_lo; it < it_hi; it += 3) {
        step(it, qa, qb, qc);
        if (it + 1 < it_hi) step(it + 1, qc, qa, qb);
        if (it + 2 < it_hi) step(it + 2, qb, qc, qa);
      }
      if (!tk.tiles_only) tk.out = nullptr;   // first tile only
      continue;
    }
    const bool ca_mma = ROW_DEC ? false : ca;   // (ROW_DEC: only the dot decoder gets here)
    // iterations [it_lo, it_hi) of each bin handled by this work item (all of them unless the tile is sliced)
    const int n_it_a = (seg_cnt(0) * a.T + 127) / 128, n_it_b = (seg_cnt(1) * a.T + 127) / 128;
    const int per_a = (n_it_a + ns - 1) / ns, per_b = (n_it_b + ns - 1) / ns;
    const int lo_a = min(n_it_a, slice * per_a), hi_a = min(n_it_a, lo_a + per_a);
    const int lo_b = min(n_it_b, slice * per_b), hi_b = min(n_it_b, lo_b + per_b);
    const int n_iter = (hi_a - lo_a) + (hi_b - lo_b);
    // iteration j of this work item -> (bin, iteration within the bin)
    auto locate = [&](int j, int& bin, int& it) {
      const int n0 = hi_a - lo_a;
      bin = j < n0 ? 0 : 1;
      it = j < n0 ? lo_a + j : lo_b + (j - n0);
    };
    // this thread's row of iteration (bin, it): segment index in the tile, candidate index, validity
    auto row_of = [&](int bin, int it, int& sg, int& t) -> bool {
      const int f = it * 128 + c.row;
      if (f >= seg_cnt(bin) * a.T) {
        sg = seg_base(bin);
        t = 0;
        return false;
      }
      const int qn = f / a.T;
      sg = seg_base(bin) + qn;
      t = f - qn * a.T;
      return true;
    };
    auto cand_id = [&](bool valid, int sg, int t) -> int {
      if (!valid) return 0;
      return a.cat_lo > 0 ? a.cat_lo + t : ldg_now_i(a.o_x + (long long)s.uuser[sg] * a.T + t);
    };
    auto gather = [&](int id, int sg, int t, float(&e)[32], float(&cv)[8], float& twv) {
      embed_load<H>(a, c, id, tab,
                    uctx ? nullptr : a.o_c + (long long)s.uuser[sg] * a.oc_user + (long long)t * a.oc_tgt, e, cv);
      if (ca_mma && id != 0 && c.half == 0) twv = ldg_now(a.tw + id);
    };
    int bin0 = 0, it0 = 0, bin1 = 0, it1 = 0, bin2 = 0, it2 = 0;
    int sg0 = 0, t0r = 0, sg1 = 0, t1r = 0;
    bool v0 = false;
    if (n_iter > 0) {
      locate(0, bin0, it0);
      v0 = row_of(bin0, it0, sg0, t0r);
    }
    if (c.half == 1) s.oid[c.row] = n_iter > 0 ? cand_id(v0, sg0, t0r) : 0;
    __syncthreads();   // s.oid, K/V/plast/cvecs stores are visible
    int oid = s.oid[c.row];
    __syncthreads();   // s.oid is rewritten at the top of iteration 0
    float e[32], cv[8], twv = 0.f;
    gather(oid, sg0, t0r, e, cv, twv);
    int idn = 0;
    if (n_iter > 1) {
      locate(1, bin1, it1);
      const bool v1 = row_of(bin1, it1, sg1, t1r);
      if (c.half == 1) idn = cand_id(v1, sg1, t1r);
    }
    // ---- cross-attention with two heads: SOFTWARE-PIPELINED loop.  Iteration q writes its query operand, hands it
    // to the iteration's two issuing warps (a 256-count named barrier on which every other warp only ARRIVES),
    // starts the gather of q+1 and then does softmax + score of iteration q-1, whose score MMAs were issued one
    // iteration ago: MMA issue (~1.1 K cycles of one thread) and execution overlap the row work instead of sitting
    // between two CTA-wide barriers.  Query operands and scores are double-buffered in TMEM (free since the
    // attention output is folded into uval):  buffer 0: Q hi/lo at ACC_Q / QN_LO, scores at X_HI / X_LO;
    // buffer 1: Q hi/lo at ACC_K / ACC_V, scores at QN_HI / ONES.  The issuer role rotates over the four warp pairs.
    const bool pipelined = ca_mma && H == 2;
    if (pipelined) {
      float p_acc = 0.f;
      uint32_t p_bits = 0;
      int p_W = 32, p_kw0 = 0, p_ubin = 0, p_usr = 0, p_t0r = 0;
      bool p_v0 = false;
#pragma unroll 1
      for (int q = 0; q <= n_iter; ++q) {
        const bool cur = q < n_iter;   // q == n_iter only drains iteration n_iter - 1
        const bool has1 = q + 1 < n_iter, has2 = q + 2 < n_iter;
        const int ubin = bin0;
        float acc = 0.f;
        uint32_t cross_bits = 0;
        int W = 32, kw0 = 0, usr = 0, oid_next = 0, sg2 = 0, t2r = 0;
        if (cur) {
          usr = s.uuser[sg0];
          const int ul = s.ulist[sg0];
          const int useg0 = ul & 63, ulen = ul >> 8;
          if (uctx) embed_finish_user<H>(c, oid, cvecs + sg0 * 64, e);
          else embed_finish<H>(a, c, oid, ctab, nullptr, e, cv);
          tick(tk, 14);
          const int f_first = it0 * 128, f_last = min(f_first + 127, seg_cnt(ubin) * a.T - 1);
          const int ul_a = s.ulist[seg_base(ubin) + f_first / a.T], ul_b = s.ulist[seg_base(ubin) + f_last / a.T];
          const int w_lo = (ul_a & 63) & ~7, w_hi = (((ul_b & 63) + (ul_b >> 8)) + 7) & ~7;
          const int wl = w_hi - w_lo;
          W = wl <= 16 ? 8 : (wl <= 32 ? 16 : 32);
          kw0 = min(w_lo, 64 - 2 * W);
          if (oid != 0) {
            const unsigned long long valid = ((unsigned long long)s.kbits[ubin][1] << 32) | s.kbits[ubin][0];
            const unsigned long long segm = (ulen >= 64 ? ~0ull : ((1ull << ulen) - 1ull)) << useg0;
            cross_bits = (uint32_t)((valid & segm) >> (kw0 + c.half * W));
            if (W < 32) cross_bits &= (1u << W) - 1u;
            if (a.residual_ca && c.half == 0) {   // residual term <o, wf> (:343,:345) from the folded tables
              acc = twv;
              if (uctx) {
                acc += cws[sg0];
              } else {
#pragma unroll
                for (int k = 0; k < 8; ++k) acc = fmaf(s.mcw[k], cv[k], acc);
              }
            }
          }
          const int qb = q & 1;
          st_operand<H>(c, qb ? C_ACCK : C_ACCQ, qb ? C_ACCV : C_QNLO, e);
          bool v2 = false;
          if (has2) {
            locate(q + 2, bin2, it2);
            v2 = row_of(bin2, it2, sg2, t2r);
          }
          if (c.half == 1) {   // s.oid <- ids of q+1; fetch ids of q+2
            s.oid[c.row] = has1 ? idn : 0;
            if (has2) idn = cand_id(v2, sg2, t2r);
          }
          tick(tk, 30);
          umma::tmem_st_wait();
          umma::fence_before_sync();
          const int iwq = ((wu & 3) == (q & 3)) ? (wu >> 2) : -1;   // this iteration's issuers: one warp pair, one head each
          if (iwq >= 0) {
            named_bar_sync(5 + qb, TC_THREADS);
            umma::fence_after_sync();
            if (umma::elect_one()) {
              const int h = iwq;
              const uint32_t koff = (uint32_t)(h * DH / 4) * 2048u + (uint32_t)(ubin * 64 + kw0) * 16u;
              const uint32_t dcol = tmem0 + (qb ? (h ? C_ONES : C_QNHI) : (h ? C_XLO : C_XHI)) + kw0;
              const uint32_t qh = tmem0 + (qb ? C_ACCK : C_ACCQ) + h * DH, ql = tmem0 + (qb ? C_ACCV : C_QNLO) + h * DH;
              const uint32_t kh = umma::smem_u32(s.k_hi) + koff, kl = umma::smem_u32(s.k_lo) + koff;
              if (W == 8) issue_3x<16, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              else if (W == 16) issue_3x<32, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              else issue_3x<64, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              commit(c);
            }
          } else {
            named_bar_arrive(5 + qb, TC_THREADS);
          }
          c.ncommit++;
          tick(tk, 33);
          // s.oid was written by the row's half-1 thread above: pair barrier, then both halves read it
          named_bar_sync(c.pair_bar, 64);
          if (has1) {
            oid_next = s.oid[c.row];
            gather(oid_next, sg1, t1r, e, cv, twv);
          }
          if (q == 0) named_bar_sync(c.pair_bar, 64);   // (later iterations: the pair exchanges below order the next write)
          tick(tk, 32);
        }
        if (q > 0) {   // softmax + score of iteration q-1
          wait_mma(c);
          tick(tk, 20);
          const int pb = (q - 1) & 1;
          const uint32_t t = c.tmem + p_kw0 + p_W * c.half;
          const uint32_t s0 = t + (pb ? C_QNHI : C_XHI), s1 = t + (pb ? C_ONES : C_XLO);
          const float* u0 = &s.uval[0][p_ubin * 64 + p_kw0 + p_W * c.half];
          const float* u1 = &s.uval[1][p_ubin * 64 + p_kw0 + p_W * c.half];
          if (p_W == 8) p_acc += softmax_pair_dot<8>(c, p_bits, sc, s0, s1, u0, u1);
          else if (p_W == 16) p_acc += softmax_pair_dot<16>(c, p_bits, sc, s0, s1, u0, u1);
          else p_acc += softmax_pair_dot<32>(c, p_bits, sc, s0, s1, u0, u1);
          tick(tk, 21);
          p_acc += pair_exchange(c, make_float2(p_acc, 0.f)).x;
          p_acc += __ldg(a.dbf);
          if (c.half == 0 && p_v0) a.y[(long long)p_usr * a.ldy + a.col0 + p_t0r] = 1.0f / (1.0f + expf(-p_acc));
          tick(tk, 23);
        }
        p_acc = acc; p_bits = cross_bits; p_W = W; p_kw0 = kw0; p_ubin = ubin; p_usr = usr; p_t0r = t0r;
        p_v0 = cur && v0;
        if (cur) {   // shift the pipeline: (q+1) becomes current, (q+2) becomes next
          if (has1) v0 = row_of(bin1, it1, sg0, t0r);
          bin0 = bin1; it0 = it1;
          bin1 = bin2; it1 = it2;
          sg1 = sg2; t1r = t2r;
          oid = oid_next;
        }
      }
    }
#pragma unroll 1
    for (int q = 0; q < (pipelined ? 0 : n_iter); ++q) {
      const bool has1 = q + 1 < n_iter, has2 = q + 2 < n_iter;
      const int ubin = bin0;
      const int usr = s.uuser[sg0], ul = s.ulist[sg0];
      const int useg0 = ul & 63, ulen = ul >> 8;
      // candidate embedding (:426), or its query for `ca`
      if (uctx) embed_finish_user<H>(c, oid, cvecs + sg0 * 64, e);
      else embed_finish<H>(a, c, oid, ctab, nullptr, e, cv);
      tick(tk, 14);
      float acc = 0.f;
      uint32_t cross_bits = 0;
      int W = 32, kw0 = 0;
      if (ca_mma) {
        // key window of the iteration: the thread pair covers 2W consecutive keys starting at kw0 (multiple of 8)
        // that contain the segments of every user present in these 128 rows; softmax and PV touch only it
        const int f_first = it0 * 128, f_last = min(f_first + 127, seg_cnt(ubin) * a.T - 1);
        const int ul_a = s.ulist[seg_base(ubin) + f_first / a.T], ul_b = s.ulist[seg_base(ubin) + f_last / a.T];
        const int w_lo = (ul_a & 63) & ~7, w_hi = (((ul_b & 63) + (ul_b >> 8)) + 7) & ~7;
        const int wl = w_hi - w_lo;
        W = wl <= 16 ? 8 : (wl <= 32 ? 16 : 32);
        kw0 = min(w_lo, 64 - 2 * W);
        if (oid != 0) {
          const unsigned long long valid = ((unsigned long long)s.kbits[ubin][1] << 32) | s.kbits[ubin][0];
          const unsigned long long segm = (ulen >= 64 ? ~0ull : ((1ull << ulen) - 1ull)) << useg0;
          cross_bits = (uint32_t)((valid & segm) >> (kw0 + c.half * W));
          if (W < 32) cross_bits &= (1u << W) - 1u;
          if (a.residual_ca && c.half == 0) {   // residual term <o, wf> (:343,:345) from the folded tables
            acc = twv;
            if (uctx) {
              acc += cws[sg0];
            } else {
#pragma unroll
              for (int k = 0; k < 8; ++k) acc = fmaf(s.mcw[k], cv[k], acc);
            }
          }
        }
        st_operand<H>(c, C_ACCQ, C_QNLO, e);
      } else {   // dot product with the last profile position (:362)
#pragma unroll
        for (int h = 0; h < H; ++h)
#pragma unroll
          for (int qq = 0; qq < N2; ++qq) acc = fmaf(e[h * N2 + qq], plast[sg0 * 64 + Own<H>::f0(h, c.half) + qq], acc);
      }
      int sg2 = 0, t2r = 0;
      bool v2 = false;
      if (has2) {
        locate(q + 2, bin2, it2);
        v2 = row_of(bin2, it2, sg2, t2r);
      }
      if (c.half == 1) {   // s.oid <- ids of q+1 (every thread read the ids of q one sync ago); fetch ids of q+2
        s.oid[c.row] = has1 ? idn : 0;
        if (has2) idn = cand_id(v2, sg2, t2r);
      }
      tick(tk, 30);
      if (ca_mma) publish();
      else __syncthreads();
      tick(tk, 31);
      int oid_next = 0;
      if (ca_mma) {
#pragma unroll
        for (int hp = 0; hp < H; hp += 2) {
          if (iw >= 0) {   // one head per issuer
            if (umma::elect_one()) {
              const int ee = iw, h = hp + ee;
              // only the 2W keys of the iteration's window (softmax reads nothing else): N = 2W instead of 64
              const uint32_t koff = (uint32_t)(h * DH / 4) * 2048u + (uint32_t)(ubin * 64 + kw0) * 16u;
              const uint32_t dcol = tmem0 + (ee ? C_XLO : C_XHI) + kw0;
              const uint32_t qh = tmem0 + C_ACCQ + h * DH, ql = tmem0 + C_QNLO + h * DH;
              const uint32_t kh = umma::smem_u32(s.k_hi) + koff, kl = umma::smem_u32(s.k_lo) + koff;
              if (W == 8) issue_3x<16, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              else if (W == 16) issue_3x<32, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              else issue_3x<64, DH / 8>(dcol, qh, ql, kh, kl, 2048u);
              commit(c);
            }
          }
          c.ncommit++;
          tick(tk, 33);
          // candidate rows of the next iteration: in flight while this iteration's attention runs (issued after
          // the MMAs so that the issuing warp does not wait on them first)
          if (hp == 0 && has1) {
            oid_next = s.oid[c.row];
            gather(oid_next, sg1, t1r, e, cv, twv);
          }
          tick(tk, 32);
          wait_mma(c);
          tick(tk, 20);
          {
            const uint32_t t = c.tmem + kw0 + W * c.half;
            const float* u0 = &s.uval[hp][ubin * 64 + kw0 + W * c.half];
            const float* u1 = &s.uval[hp + 1][ubin * 64 + kw0 + W * c.half];
            if (W == 8) acc += softmax_pair_dot<8>(c, cross_bits, sc, t + C_XHI, t + C_XLO, u0, u1);
            else if (W == 16) acc += softmax_pair_dot<16>(c, cross_bits, sc, t + C_XHI, t + C_XLO, u0, u1);
            else acc += softmax_pair_dot<32>(c, cross_bits, sc, t + C_XHI, t + C_XLO, u0, u1);
          }
          tick(tk, 21);
          if (hp + 2 < H) {   // the next head pair's score MMAs overwrite the TMEM columns just read
            umma::fence_before_sync();
            __syncthreads();
            umma::fence_after_sync();
          }
        }
        acc += pair_exchange(c, make_float2(acc, 0.f)).x;
        acc += __ldg(a.dbf);
      } else {
        acc += pair_exchange(c, make_float2(acc, 0.f)).x;
        if (has1) {
          oid_next = s.oid[c.row];
          gather(oid_next, sg1, t1r, e, cv, twv);
        }
        __syncthreads();   // s.oid is rewritten at the top of the next iteration
      }
      if (c.half == 0 && v0) a.y[(long long)usr * a.ldy + a.col0 + t0r] = 1.0f / (1.0f + expf(-acc));
      tick(tk, 23);
      // shift the pipeline: (q+1) becomes current, (q+2) becomes next
      if (has1) v0 = row_of(bin1, it1, sg0, t0r);
      bin0 = bin1; it0 = it1;
      bin1 = bin2; it1 = it2;
      sg1 = sg2; t1r = t2r;
      oid = oid_next;
    }
    if (!tk.tiles_only) tk.out = nullptr;   // first tile only
  }
  tick(tk, 0);   // (tiles_only: end of the CTA's last tile)
  umma::fence_before_sync();
  __syncthreads();
  if (w == 0) umma::tmem_free(tmem0, 512);
}


// ---- split decoder (DEC == 3): the candidates of ALL users as one grid of (user, candidate pair) work items -------
// The in-kernel row / pair loops run on the encoder kernel's 8 warps per SM (255 registers, one CTA per SM).  This
// kernel has no TMEM, no shared memory and half the registers (16 warps per SM); keys come from the rows the encoder
// kernel exported (one 128-byte line per (key, head), read through L1 — lanes of one user read the same line),
// everything else is pair_head's arithmetic.  One context row per user only (the context is folded into kc).
// Measured: the default for catalog mode (8.4 vs 7.0 G scores/s: consecutive item ids, so the table rows of a warp
// are contiguous); for 101 sampled candidates it is no faster than the in-kernel row loop (21.9 vs 22.5 M users/s:
// one item per thread leaves the id -> table-row -> first-FFMA latency chain exposed, ncu: 72 % long-scoreboard
// stalls, and prefetching it costs the registers the higher occupancy was bought with).
struct DecPairsArgs {
  const float *Kg, *Ug, *cwsg, *TQ, *tw, *dbf;
  const int *useg, *o_x;
  float* y;
  long long ldy;
  int col0, B, T, cat_lo, residual_ca;
  float sc;
  const int* n_bins;
  int auto_dec;
};
__global__ void __launch_bounds__(256, 2) decode_pairs_kernel(const DecPairsArgs a) {
  if (tc_auto_skip(a.auto_dec, a.n_bins, a.B)) return;
  const int P = (a.T + 1) / 2;
  const long long total = (long long)a.B * P;
  const float bfv = __ldg(a.dbf);
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < total; p += (long long)gridDim.x * blockDim.x) {
    const int u = (int)(p / P);
    const int t0 = 2 * (int)(p - (long long)u * P);
    const bool two = t0 + 1 < a.T;
    int id0, id1;
    if (a.cat_lo > 0) {
      id0 = a.cat_lo + t0;
      id1 = two ? id0 + 1 : 0;
    } else {
      const int* px = a.o_x + (long long)u * a.T + t0;
      id0 = __ldg(px);
      id1 = two ? __ldg(px + 1) : 0;
    }
    const int seg = __ldg(a.useg + u);
    const int row0 = seg & 0xffffff, ulen = seg >> 24;
    float res0 = bfv, res1 = bfv, att0 = 0.f, att1 = 0.f;
    if (a.residual_ca) {
      const float cw = __ldg(a.cwsg + u);
      if (id0 != 0) res0 += __ldg(a.tw + id0) + cw;
      if (id1 != 0) res1 += __ldg(a.tw + id1) + cw;
    }
    if (id0 != 0 || id1 != 0) {
      const float4* const ug = reinterpret_cast<const float4*>(a.Ug) + row0;
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        float e[64];   // query features of head h: row 0, row 1 (zeros for a padded candidate)
        const int ids[2] = {id0, id1};
#pragma unroll
        for (int k = 0; k < 2; ++k) {
          const float4* t = reinterpret_cast<const float4*>(a.TQ + (long long)ids[k] * 64 + 32 * h);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float4 x = ids[k] != 0 ? __ldg(t + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            float* o = &e[32 * k + 4 * i];
            o[0] = x.x; o[1] = x.y; o[2] = x.z; o[3] = x.w;
          }
        }
        const float4* const kh = reinterpret_cast<const float4*>(a.Kg + (long long)row0 * 64 + 32 * h);
        float mxa = -INFINITY, za = 0.f, da = 0.f, mxb = -INFINITY, zb = 0.f, db = 0.f;
#pragma unroll 1
        for (int j = 0; j < ulen; j += 2) {
          const int j1 = min(j + 1, ulen - 1);
          float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;   // a: row 0, b: row 1; 0 / 1: key j / j1
#pragma unroll
          for (int kc = 0; kc < 8; ++kc) {
            const float4 k0 = __ldg(kh + j * 16 + kc), k1 = __ldg(kh + j1 * 16 + kc);
            const float* qa = &e[4 * kc];
            const float* qb = &e[32 + 4 * kc];
            a0 = fmaf(qa[0], k0.x, a0); a1 = fmaf(qa[0], k1.x, a1); b0 = fmaf(qb[0], k0.x, b0); b1 = fmaf(qb[0], k1.x, b1);
            a0 = fmaf(qa[1], k0.y, a0); a1 = fmaf(qa[1], k1.y, a1); b0 = fmaf(qb[1], k0.y, b0); b1 = fmaf(qb[1], k1.y, b1);
            a0 = fmaf(qa[2], k0.z, a0); a1 = fmaf(qa[2], k1.z, a1); b0 = fmaf(qb[2], k0.z, b0); b1 = fmaf(qb[2], k1.z, b1);
            a0 = fmaf(qa[3], k0.w, a0); a1 = fmaf(qa[3], k1.w, a1); b0 = fmaf(qb[3], k0.w, b0); b1 = fmaf(qb[3], k1.w, b1);
          }
          const float4 g0 = __ldg(ug + j), g1 = __ldg(ug + j1);
          const float u0 = h ? g0.y : g0.x, u1 = h ? g1.y : g1.x;
          const float c0 = h ? g0.w : g0.z, c1 = (j + 1 < ulen) ? (h ? g1.w : g1.z) : -INFINITY;
          {
            const float x0 = (a0 + c0) * a.sc, x1 = (a1 + c1) * a.sc;
            const float mn = fmaxf(mxa, fmaxf(x0, x1));
            const float mref = (mn == -INFINITY) ? 0.f : mn;
            const float corr = ex2_approx(mxa - mref), p0 = ex2_approx(x0 - mref), p1 = ex2_approx(x1 - mref);
            za = fmaf(za, corr, p0 + p1);
            da = fmaf(da, corr, fmaf(p0, u0, p1 * u1));
            mxa = mn;
          }
          {
            const float x0 = (b0 + c0) * a.sc, x1 = (b1 + c1) * a.sc;
            const float mn = fmaxf(mxb, fmaxf(x0, x1));
            const float mref = (mn == -INFINITY) ? 0.f : mn;
            const float corr = ex2_approx(mxb - mref), p0 = ex2_approx(x0 - mref), p1 = ex2_approx(x1 - mref);
            zb = fmaf(zb, corr, p0 + p1);
            db = fmaf(db, corr, fmaf(p0, u0, p1 * u1));
            mxb = mn;
          }
        }
        if (id0 != 0 && za > 0.f) att0 += da / za;
        if (id1 != 0 && zb > 0.f) att1 += db / zb;
      }
    }
    // a padded candidate scores sigmoid(bf): query mask 0 -> attention row 0, o = 0 (src/carca.py:94, :256)
    float* yp = a.y + (long long)u * a.ldy + a.col0 + t0;
    yp[0] = 1.0f / (1.0f + expf(-(res0 + att0)));
    if (two) yp[1] = 1.0f / (1.0f + expf(-(res1 + att1)));
  }
}

// Packs the valid profile positions of every user into 64-row bins (two bins = one 128-row tile of
// the kernel above).  Beauty-shaped profiles are mostly left padding (src/data.py:112-113, mean ~8
// of 50 positions valid): padded rows never influence a valid row (they are masked as keys,
// src/carca.py:246-251) and only valid rows (ca decoder: keys/values; dot decoder: position L-1)
// are read by the decoder, so the encoder runs on valid rows only.  A user's rows stay contiguous
// and in sequence order inside one bin (causal mask = row order); position L-1 is always included
// (the dot decoder reads it even when it is padding, src/carca.py:362).
//   row_src[bin*64 + r] = user*256 + position (or -1), row_seg = segment start | length << 8.
// One 128-thread block packs 128 consecutive users (next-fit) and claims its bins with one atomicAdd; it writes
// every row of its bins (-1 for the unused ones), so the scratch needs no clearing between calls.
__global__ void __launch_bounds__(128) pack_rows_kernel(int* __restrict__ row_src, int* __restrict__ row_seg,
                                                        int* __restrict__ n_bins, int* __restrict__ status,
                                                        const int* __restrict__ p_x, int B, int L,
                                                        int* __restrict__ bin_users) {
  __shared__ int cnt[128], bin_of[128], start_of[128], users_of[128], fill_of[128], base, nbins;
  const int t = threadIdx.x, usr = blockIdx.x * 128 + t;
  int n = 0;
  if (usr < B) {
    const int* x = p_x + (long long)usr * L;
    for (int j = 0; j < L; ++j) n += (x[j] != 0) || (j == L - 1);
  }
  int skip = 0;   // a bin holds 64 rows: longer profiles (possible only when L > 64) keep their LAST 64 rows
  if (n > 64) {   // and are flagged — the caller must route such batches to the per-op kernels
    skip = n - 64;
    n = 64;
    if (status) atomicOr(status, 2);
  }
  cnt[t] = n;
  __syncthreads();
  if (t == 0) {
    int bin = 0, fill = 0, in_bin = 0;
    const int users = min(128, B - blockIdx.x * 128);
    for (int q = 0; q < users; ++q) {
      if (fill + cnt[q] > 64) {
        users_of[bin] = in_bin;
        fill_of[bin] = fill;
        ++bin;
        fill = 0;
        in_bin = 0;
      }
      bin_of[q] = bin;
      start_of[q] = fill;
      fill += cnt[q];
      ++in_bin;
    }
    users_of[bin] = in_bin;
    fill_of[bin] = fill;
    nbins = bin + 1;
    base = atomicAdd(n_bins, bin + 1);
  }
  __syncthreads();
  if (bin_users && t < nbins) bin_users[base + t] = users_of[t];   // users per bin: the tile's decoder work
  if (usr < B) {
    const int* x = p_x + (long long)usr * L;
    const long long o = (long long)(base + bin_of[t]) * 64;
    int r = start_of[t];
    const int seg = start_of[t] | (n << 8);
    for (int j = 0; j < L; ++j) {
      if (x[j] != 0 || j == L - 1) {
        if (skip > 0) {
          --skip;
          continue;
        }
        row_src[o + r] = usr * 256 + j;
        row_seg[o + r] = seg;
        ++r;
      }
    }
  }
  // unused rows of the block's bins (the fused kernel treats the bin past the last one as empty itself)
  for (int i = t; i < nbins * 64; i += 128) {
    const int b = i / 64, r = i % 64;
    if (r >= fill_of[b]) row_src[(long long)(base + b) * 64 + r] = -1;
  }
}

// Tiles in descending order of their user count (= decoder work; the encoder part of a tile is constant): CTAs take
// work items from a global counter, so with the longest tiles first the last, partly filled wave consists of the
// cheapest ones (502 tiles on 148 CTAs run as 4 waves: profiles/r01/tc_tile_times.log).  Counting sort, one CTA.
__global__ void __launch_bounds__(256) tile_order_kernel(int* __restrict__ order, const int* __restrict__ bin_users,
                                                         const int* __restrict__ n_bins) {
  __shared__ int hist[130], cursor[130];
  const int nb = n_bins[0], n_tiles = (nb + 1) / 2;
  for (int i = threadIdx.x; i < 130; i += blockDim.x) hist[i] = 0;
  __syncthreads();
  auto bucket = [&](int t) {   // 128 - users, users <= 128 (a bin holds at most 64 users)
    const int u = bin_users[2 * t] + (2 * t + 1 < nb ? bin_users[2 * t + 1] : 0);
    return 128 - min(u, 128);
  };
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) atomicAdd(&hist[bucket(t)], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int i = 0; i < 129; ++i) {
      cursor[i] = acc;
      acc += hist[i];
    }
  }
  __syncthreads();
  for (int t = threadIdx.x; t < n_tiles; t += blockDim.x) order[atomicAdd(&cursor[bucket(t)], 1)] = t;
}


// Packs W [64 out, 64 in] (+ bias [64]) into the K-major B operand of umma.cuh with the bias as
// K step 8: dst = [hi: 18 chunks x 64 rows x 4 | lo: same].
__global__ void __launch_bounds__(256) pack_weight_tc_kernel(float* __restrict__ dst, const float* __restrict__ W,
                                                             const float* __restrict__ bias) {
  for (int e = threadIdx.x + blockIdx.x * blockDim.x; e < TC_WFLOATS; e += blockDim.x * gridDim.x) {
    const int j = e % 4, n = (e / 4) % 64, kc = e / 256;
    float v = 0.f;
    if (kc < 16) v = W[n * 64 + 4 * kc + j];
    else if (kc == 16 && j == 0) v = bias[n];
    dst[e] = v;
    dst[TC_WFLOATS + e] = umma::tf32_lo(v);
  }
}

}  // namespace carca
#endif  // CARCA_EMU
