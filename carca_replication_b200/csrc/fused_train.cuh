// Fused TRAINING core of CARCA for d = 64, L <= 64 (the Beauty configuration): everything between the
// embeddings and the probabilities, forward and backward, in two kernels
//   forward : dropout(p_e) -> n_blocks x SelfAttentionBlock -> final LayerNorm -> decoder -> y
//   backward: dy -> d p_e, d o_e, every parameter gradient of the blocks / final norm / decoder
// Reference path replaced: CARCA.forward in train mode, src/carca.py:416-431, with
// SelfAttentionBlock :297-318, MultiHeadAttention :228-265, CrossAttentionBlock :338-349 (causal -1),
// DotProduct :358-360, and autograd's backward of all of them (src/train.py:95).
//
// Why: at the reference batch size the per-op path is ~250 launches of microsecond kernels over
// [B*L, 64] tensors that are mostly LEFT PADDING (src/data.py:112-113; Beauty: ~8 of 50 positions).
// A padded position reaches no loss term: it is masked as a key (:246-251), its target rows carry a
// zero loss mask (src/train.py:92), so its weight-gradient contribution is exactly zero.  The kernels
// therefore run on ACTIVE positions only (profile id != 0 or a target id != 0), packed user after
// user into 64-row bins (train_pack_kernel); a CTA owns a bin and keeps its activations in shared
// memory as row-major [64][68] fp32 tiles.  Row r of a bin is the same (user, position) for the
// profile and for every target tuple, so the decoder's causal rule "target i sees profile keys j < i"
// is "row sees earlier rows of its own segment".
//
// All products are fp32 FFMA thread-tiled GEMMs on shared-memory operands (exactly the reference's
// arithmetic type; no tf32 rounding in gradients):
//   gemm_nt  C[r][n] = sum_k A[r][k] W[n][k]      projections (weights as stored, [out][in])
//   gemm_nn  C[r][c] = sum_k A[r][k] B[k][c]      input gradients, P.V, dS.K
//   gemm_tn  C[n][c] = sum_r A[r][n] X[r][c]      weight gradients (atomics into the grad tensors), dK, dV
// The forward stores six [rows, 64] activations per block (+ K, V, s of the decoder) for the backward;
// LayerNorm outputs, Q of the decoder and every dropout mask (Philox, same element indexing as the
// per-op kernels, so both paths draw identical masks) are recomputed.
#pragma once
#include "common.cuh"

namespace carca {

constexpr int TR = 64;            // rows per bin
constexpr int TD = 64;            // model width
constexpr int TLD = 68;           // row stride of a tile in shared memory (floats)
constexpr int TBUF = TR * TLD;    // floats per tile
constexpr int TTHREADS = 512;
constexpr int TWARPS = TTHREADS / 32;
constexpr int TMAXB = 8;          // encoder blocks
constexpr int TMAXT = 2;          // target tuples (positives | negatives, src/train.py:86-88)
constexpr int TNBUF = 10;
constexpr int TMAXC = 8;          // context features the in-kernel embedding handles

struct TrainBlockW { const float *ln1_g, *ln1_b, *wq, *bq, *wk, *bk, *wv, *bv, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2; };
struct TrainBlockG { float *ln1_g, *ln1_b, *wq, *bq, *wk, *bk, *wv, *bv, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2; };
struct TrainCrossW { const float *wq, *bq, *wk, *bk, *wv, *bv, *wf, *bf; };
struct TrainCrossG { float *wq, *bq, *wk, *bk, *wv, *bv, *wf, *bf; };

// row_info bits
constexpr int TI_PVALID = 1 << 16;
constexpr int TI_TVALID0 = 1 << 17;   // tuple t: TI_TVALID0 << t

struct TrainArgs {
  int B, L, H, n_blocks, n_tuples, decoder;   // decoder: 0 dot-product, 1 cross-attention
  int n_sms;                                  // CTAs the bins should spread over (packing heuristic)
  int residual_sa, residual_ca;
  DropCfg drop;                               // p / seed / device seed word; the site is set per use
  const int* p_x;                             // [B, L]
  const int* o_x[TMAXT];                      // [B, L] each
  const float* p_e;                           // [B, L, 64] embedded profile (masked, before dropout)
  const float* o_e[TMAXT];                    // [B, L, 64] embedded targets (masked)
  float* y;                                   // [B, ldy]; tuple t at columns t*L
  long long ldy;
  int* row_src;                               // [B*64] bin row -> u*L + i, or -1
  int* row_info;                              // [B*64] seg start | seg last << 8 | validity bits
  int* n_bins;                                // [1]
  float* sv;                                  // saved activations: tensor k at sv + k*sv_stride, [B*64, 64]
  long long sv_stride;
  TrainBlockW blk[TMAXB];
  const float *fn_g, *fn_b;
  TrainCrossW dec;
  // folded in-kernel embedding (embed_mode == 1): e = mask * (Wj_z sqrt(d) E[x] + sum_attr val GT[attr] + sum_k c_k GT[A+k]
  //                                                              + cst (+ pos)),  GT = (Wj_q Wf)^T, cst = Wj_q bf + bj
  int embed_mode, C, A, ldj;                  // ldj = d + g, row stride of joint_embed.weight
  float sqrt_d;
  const float* E;                             // items_embed.weight [n_items, 64]
  const float* Wj;                            // joint_embed.weight [64, ldj]; columns 0..63 multiply sqrt(d) E[x]
  const float* fold;                          // [A + C + 1, 64]: GT rows, then cst
  const float* pos_table;                     // optional [>= L, 64], profile rows only
  const int *csr_rowptr, *csr_cols;
  const float* csr_vals;
  const float* p_c;                           // [B, L, C]
  const float* o_c[TMAXT];
  float *gE, *gWj, *dfold, *gpos;             // backward: d items_embed, d joint_embed.weight, d fold, d pos
  long long* ticks;                           // optional (development): (label, clock64) pairs of CTA 0, thread 0
  int ticks_cap;
  // backward only
  const float* dy;                            // [B, ldy]
  float* d_pe;                                // [B, L, 64] zero-initialised by the caller
  float* d_oe[TMAXT];
  TrainBlockG gblk[TMAXB];
  float *g_fn_g, *g_fn_b;
  TrainCrossG gdec;
};

// saved-tensor slots
__host__ __device__ __forceinline__ int sv_block(int b, int k) { return 6 * b + k; }     // k: 0 x, 1 Q, 2 K, 3 V, 4 s, 5 a1
__host__ __device__ __forceinline__ int sv_final(int nb) { return 6 * nb; }              // last block's output
__host__ __device__ __forceinline__ int sv_dec(int nb, int k) { return 6 * nb + 1 + k; } // 0 K, 1 V, 2 + t: s of tuple t
__host__ __device__ __forceinline__ int sv_count(int nb, int nt) { return 6 * nb + 3 + 2 * nt; }   // + nt + t: embedded targets

struct TrainSmem {
  float buf[TNBUF][TBUF];
  float w[TBUF];                   // weight staging, [n][k] with row stride TLD
  float red[TWARPS][2][TD];        // per-warp partial sums of vector gradients
  float gsc[TR];                   // per-row scalars
  int src[TR], info[TR], pos[TR], usr[TR];
  int ids[1 + TMAXT][TR];          // embed_mode: item id of the row in the profile / each target tuple
  float gtc[TMAXC + 1][TD];        // embed_mode: context rows of the folded table, then cst (loaded once per CTA)
  float ctxv[TR][TMAXC];           // embed_mode: context values of the row set being embedded
  int n, npad;
};

struct TrainTicks {
  long long* out;
  int n, cap;
};
__device__ __forceinline__ TrainTicks ticks_begin(const TrainArgs& a) {
  TrainTicks t;
  t.out = (a.ticks && blockIdx.x == 0 && threadIdx.x == 0) ? a.ticks : nullptr;
  t.n = 0;
  t.cap = a.ticks_cap;
  return t;
}
__device__ __forceinline__ void tick(TrainTicks& t, int label) {
#ifndef CARCA_EMU
  if (t.out && t.n < t.cap) {
    t.out[2 * t.n] = label;
    t.out[2 * t.n + 1] = clock64();
    ++t.n;
  }
#endif
}

// -------------------------------------------------------------------------------------------------- packing
// One CTA per 128 users: counts each user's active positions, packs users greedily into 64-row bins,
// writes the row maps, and fills y of every INACTIVE position with the value the reference computes
// there (sigmoid(ffn bias) for the cross-attention decoder: s = 0; 0.5 for the dot product).
constexpr int TMAXL = 256;        // longest window the packing pass handles (a user's ACTIVE positions must fit a bin)

__global__ void __launch_bounds__(1024) train_pack_kernel(const TrainArgs a) {
  __shared__ unsigned char flg[128][TMAXL];   // per (user of this CTA, position): validity bits
  __shared__ int cnt[128], skip_of[128], bin_of[128], start_of[128], base;
  const int t = threadIdx.x;
  const int L = a.L;
  const int user0 = blockIdx.x * 128;
  const int users = min(128, a.B - user0);
  const float ydef = a.decoder == 1 ? 1.0f / (1.0f + expf(-a.dec.bf[0])) : 0.5f;
  for (int e = t; e < users * L; e += blockDim.x) {          // coalesced sweep over the CTA's positions
    const int u = e / L, j = e % L;
    const long long g = (long long)(user0 + u) * L + j;
    int f = a.p_x[g] != 0 ? 1 : 0;
    for (int q = 0; q < a.n_tuples; ++q) f |= (a.o_x[q][g] != 0 ? 2 : 0) << q;
    flg[u][j] = (unsigned char)f;
    if (!f)
      for (int q = 0; q < a.n_tuples; ++q) a.y[(long long)(user0 + u) * a.ldy + q * L + j] = ydef;
  }
  __syncthreads();
  if (t < 128) {
    int n = 0;
    if (t < users)
      for (int j = 0; j < L; ++j) n += flg[t][j] != 0;
    // a bin holds 64 rows: a user with more active positions (possible only when L > 64) keeps the LAST 64 and
    // is flagged in n_bins[1] — the host routes such batches to the per-op kernels before calling (carca.py)
    skip_of[t] = n > TR ? n - TR : 0;
    if (n > TR) {
      n = TR;
      atomicAdd(a.n_bins + 1, 1);
    }
    cnt[t] = n;
  }
  __syncthreads();
  if (t == 0) {
    // bin capacity: 64 rows when there is enough work for every SM, fewer (>= 16) for small batches so the rows
    // spread over more CTAs and each CTA's dependent chain of phases gets shorter; a user always fits (<= 64 rows)
    int total = 0;
    for (int q = 0; q < users; ++q) total += cnt[q];
    const long long est = (long long)total * a.B / users;            // rows of the whole batch, extrapolated
    int cap = (int)((est / a.n_sms + 7) / 8 * 8);
    cap = cap < 16 ? 16 : (cap > TR ? TR : cap);
    int bin = 0, fill = 0;
    for (int q = 0; q < users; ++q) {
      if (fill > 0 && fill + cnt[q] > cap) {
        ++bin;
        fill = 0;
      }
      bin_of[q] = bin;
      start_of[q] = fill;
      fill += cnt[q];
    }
    base = atomicAdd(a.n_bins, bin + 1);
  }
  __syncthreads();
  if (t < users && cnt[t] > 0) {
    const long long o = (long long)(base + bin_of[t]) * TR;
    int r = start_of[t];
    int skip = skip_of[t];
    const int seg = start_of[t] | ((start_of[t] + cnt[t] - 1) << 8);
    for (int j = 0; j < L; ++j) {
      const int f = flg[t][j];
      if (f) {
        if (skip > 0) {
          --skip;
          for (int q = 0; q < a.n_tuples; ++q) a.y[(long long)(user0 + t) * a.ldy + q * L + j] = ydef;   // (flagged)
          continue;
        }
        a.row_src[o + r] = (user0 + t) * L + j;
        a.row_info[o + r] = seg | (f << 16);
        ++r;
      }
    }
  }
}

// Thread mapping of the tile products (512 threads = 16 warps): warp w owns the 8 rows [8 (w >> 1), +8) and the 32
// columns [32 (w & 1), +32) of a 64 x 64 output; lane = (ty, tx) = (lane >> 3, lane & 7) owns rows 2 ty, 2 ty + 1 of
// the warp's rows.  A warp thus reads 8 operand rows and 32 operand columns per step: every shared-memory request
// is one conflict-free wavefront and the products are FFMA-bound, not shared-memory-bound.
static_assert(TTHREADS == 512 && TR == 64 && TD == 64, "tile mapping below assumes 16 warps over a 64 x 64 tile");
constexpr int TNR = 2;   // rows per thread

struct TileMap {
  int r0, c0, tx;
  __device__ __forceinline__ TileMap() {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    tx = lane & 7;
    r0 = 8 * (w >> 1) + 2 * (lane >> 3);
    c0 = 32 * (w & 1);
  }
};

// C[r][n] = sum_k A[r][k] W[n][k], K % 4 == 0; thread owns rows r0, r0 + 1 and the strided columns c0 + tx + 8 j
// (j < 4), so the 128-bit loads of W rows are bank-conflict free.  `need(r0, col)` lets attention skip columns
// outside the rows' key range.
template <int NCOLS, class Need, class Epi>
__device__ __forceinline__ void gemm_nt(const float* __restrict__ A, int lda, const float* __restrict__ W, int ldw,
                                        int K, int nrows, Need need, Epi epi) {
  static_assert(NCOLS == 64, "gemm_nt covers 64 output columns");
  const TileMap m;
  constexpr int NJ = 4;
  if (m.r0 < nrows) {
    float acc[TNR][NJ];
    bool on[NJ];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      on[j] = need(m.r0, m.c0 + m.tx + 8 * j);
#pragma unroll
      for (int i = 0; i < TNR; ++i) acc[i][j] = 0.f;
    }
    const float* ap = A + m.r0 * lda;
    const float* wp = W + (m.c0 + m.tx) * ldw;
#pragma unroll 2
    for (int k = 0; k < K; k += 4) {
      float4 av[TNR];
#pragma unroll
      for (int i = 0; i < TNR; ++i) av[i] = *reinterpret_cast<const float4*>(ap + i * lda + k);
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        if (on[j]) {
          const float4 b = *reinterpret_cast<const float4*>(wp + 8 * j * ldw + k);
#pragma unroll
          for (int i = 0; i < TNR; ++i) {
            acc[i][j] = fmaf(av[i].x, b.x, acc[i][j]);
            acc[i][j] = fmaf(av[i].y, b.y, acc[i][j]);
            acc[i][j] = fmaf(av[i].z, b.z, acc[i][j]);
            acc[i][j] = fmaf(av[i].w, b.w, acc[i][j]);
          }
        }
      }
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j)
      if (on[j]) {
#pragma unroll
        for (int i = 0; i < TNR; ++i) epi(m.r0 + i, m.c0 + m.tx + 8 * j, acc[i][j]);
      }
  }
}

struct NeedAll {
  __device__ __forceinline__ bool operator()(int, int) const { return true; }
};

// Output rows / columns of a thread for products with NC output columns (4 contiguous columns per thread):
//   NC = 64: the TileMap above, 2 rows;   NC = 32: warp w owns rows 4w .. 4w+3, 1 row per thread;
//   NC = 16: warps 0..7 own rows 8w .. 8w+7, 1 row per thread.
template <int NC>
struct OutMap {
  static constexpr int RT = NC == 64 ? 2 : 1;
  int r0, c0;
  bool active;
  __device__ __forceinline__ OutMap() {
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (NC == 64) {
      r0 = 8 * (w >> 1) + 2 * (lane >> 3);
      c0 = 32 * (w & 1) + 4 * (lane & 7);
      active = true;
    } else if (NC == 32) {
      r0 = 4 * w + (lane >> 3);
      c0 = 4 * (lane & 7);
      active = true;
    } else {
      r0 = 8 * w + (lane >> 2);
      c0 = 4 * (lane & 3);
      active = w < 8;
    }
  }
};

// C[r][c] = sum_{k in [klo, khi)} A[r][k] B[k][c] for c < NC (NC in {16, 32, 64}).  range(r0, r1, klo, khi) gives the
// reduction range of rows r0..r1; it is widened to multiples of 4 (A is read 128 bits at a time), so A must hold
// zeros and B finite values just outside it.
template <int NC, class Range, class Epi>
__device__ __forceinline__ void gemm_nn(const float* __restrict__ A, int lda, const float* __restrict__ Bm, int ldb,
                                        int nrows, Range range, Epi epi) {
  const OutMap<NC> m;
  constexpr int RT = OutMap<NC>::RT;
  if (m.active && m.r0 < nrows) {
    int klo, khi;
    range(m.r0, m.r0 + RT - 1, klo, khi);
    klo &= ~3;
    khi = min((khi + 3) & ~3, TR);
    float acc[RT][4];
#pragma unroll
    for (int i = 0; i < RT; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
    const float* ap = A + m.r0 * lda;
    const float* bp = Bm + m.c0;
#pragma unroll 2
    for (int k = klo; k < khi; k += 4) {
      float4 av[RT];
#pragma unroll
      for (int i = 0; i < RT; ++i) av[i] = *reinterpret_cast<const float4*>(ap + i * lda + k);
      const float4 b0 = *reinterpret_cast<const float4*>(bp + (k + 0) * ldb);
      const float4 b1 = *reinterpret_cast<const float4*>(bp + (k + 1) * ldb);
      const float4 b2 = *reinterpret_cast<const float4*>(bp + (k + 2) * ldb);
      const float4 b3 = *reinterpret_cast<const float4*>(bp + (k + 3) * ldb);
#pragma unroll
      for (int i = 0; i < RT; ++i) {
        acc[i][0] = fmaf(av[i].x, b0.x, acc[i][0]); acc[i][1] = fmaf(av[i].x, b0.y, acc[i][1]);
        acc[i][2] = fmaf(av[i].x, b0.z, acc[i][2]); acc[i][3] = fmaf(av[i].x, b0.w, acc[i][3]);
        acc[i][0] = fmaf(av[i].y, b1.x, acc[i][0]); acc[i][1] = fmaf(av[i].y, b1.y, acc[i][1]);
        acc[i][2] = fmaf(av[i].y, b1.z, acc[i][2]); acc[i][3] = fmaf(av[i].y, b1.w, acc[i][3]);
        acc[i][0] = fmaf(av[i].z, b2.x, acc[i][0]); acc[i][1] = fmaf(av[i].z, b2.y, acc[i][1]);
        acc[i][2] = fmaf(av[i].z, b2.z, acc[i][2]); acc[i][3] = fmaf(av[i].z, b2.w, acc[i][3]);
        acc[i][0] = fmaf(av[i].w, b3.x, acc[i][0]); acc[i][1] = fmaf(av[i].w, b3.y, acc[i][1]);
        acc[i][2] = fmaf(av[i].w, b3.z, acc[i][2]); acc[i][3] = fmaf(av[i].w, b3.w, acc[i][3]);
      }
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) epi(m.r0 + i, m.c0, acc[i]);
  }
}

// C[n][c] = sum_{r in [rlo, rhi)} A[r][n] X[r][c] for n < 64, c < NC.  range(n0, n1, rlo, rhi) gives the reduction
// range for output rows n0..n1.
template <int NC, class Range, class Epi>
__device__ __forceinline__ void gemm_tn(const float* __restrict__ A, int lda, const float* __restrict__ X, int ldx,
                                        Range range, Epi epi) {
  const OutMap<NC> m;
  constexpr int RT = OutMap<NC>::RT;
  if (!m.active) return;
  int rlo, rhi;
  range(m.r0, m.r0 + RT - 1, rlo, rhi);
  float acc[RT][4];
#pragma unroll
  for (int i = 0; i < RT; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const float* ap = A + m.r0;
  const float* xp = X + m.c0;
#pragma unroll 4
  for (int r = rlo; r < rhi; ++r) {
    const float4 x = *reinterpret_cast<const float4*>(xp + r * ldx);
    float av[RT];
    if (RT == 2) {
      const float2 t = *reinterpret_cast<const float2*>(ap + r * lda);
      av[0] = t.x;
      av[RT - 1] = t.y;
    } else {
      av[0] = ap[r * lda];
    }
#pragma unroll
    for (int i = 0; i < RT; ++i) {
      acc[i][0] = fmaf(av[i], x.x, acc[i][0]);
      acc[i][1] = fmaf(av[i], x.y, acc[i][1]);
      acc[i][2] = fmaf(av[i], x.z, acc[i][2]);
      acc[i][3] = fmaf(av[i], x.w, acc[i][3]);
    }
  }
#pragma unroll
  for (int i = 0; i < RT; ++i) epi(m.r0 + i, m.c0, acc[i]);
}

// -------------------------------------------------------------------------------------------------- helpers
// four keep-scales of the aligned flat elements 4q .. 4q+3 of a site's tensor (one Philox call)
__device__ __forceinline__ void drop4(const DropCfg& c, unsigned long long q, float (&f)[4]) {
  if (c.p <= 0.f) {
    f[0] = f[1] = f[2] = f[3] = 1.0f;
    return;
  }
  unsigned k0 = c.seed_lo, k1 = c.seed_hi;
  if (c.seed_dev) {
    const unsigned long long x = *c.seed_dev;
    k0 ^= (unsigned)x;
    k1 ^= (unsigned)(x >> 32);
  }
  const Philox4 r = philox4x32_10((unsigned)q, (unsigned)(q >> 32), c.site, 0u, k0, k1);
  const unsigned wds[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float u = (float)(wds[i] >> 8) * 5.9604644775390625e-08f;
    f[i] = u >= c.p ? c.scale : 0.0f;
  }
}

__device__ __forceinline__ DropCfg at_site(DropCfg c, unsigned site) {
  c.site = site;
  return c;
}

// loads the 64 x 64 weight matrix Wg (row-major, as stored) into s.w with row stride TLD
__device__ __forceinline__ void load_w(TrainSmem& s, const float* __restrict__ Wg, int ldg = TD) {
#pragma unroll
  for (int i = 0; i < TR * 16 / TTHREADS; ++i) {
    const int e = threadIdx.x + i * TTHREADS;   // float4 index, 16 per row
    const int n = e >> 4, k4 = e & 15;
    *reinterpret_cast<float4*>(s.w + n * TLD + k4 * 4) = *reinterpret_cast<const float4*>(Wg + (long long)n * ldg + k4 * 4);
  }
}

// tile <-> saved activation [bin*64 + r][64]
__device__ __forceinline__ void save_tile(const TrainArgs& a, const TrainSmem& s, int slot, int bin, const float* tile) {
  float* g = a.sv + slot * a.sv_stride + (long long)bin * TR * TD;
  for (int e = threadIdx.x; e < s.n * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    *reinterpret_cast<float4*>(g + r * TD + c4 * 4) = *reinterpret_cast<const float4*>(tile + r * TLD + c4 * 4);
  }
}
__device__ __forceinline__ void load_tile(const TrainArgs& a, const TrainSmem& s, int slot, int bin, float* tile) {
  const float* g = a.sv + slot * a.sv_stride + (long long)bin * TR * TD;
  for (int e = threadIdx.x; e < TR * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < s.n) v = *reinterpret_cast<const float4*>(g + r * TD + c4 * 4);
    *reinterpret_cast<float4*>(tile + r * TLD + c4 * 4) = v;
  }
}

// rows of a [B, L, 64] tensor -> tile (zeros below the bin's fill), optional dropout of site `site`
__device__ __forceinline__ void gather_rows(const TrainSmem& s, const float* __restrict__ src, float* tile,
                                            const DropCfg* drop) {
  for (int e = threadIdx.x; e < TR * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < s.n) {
      v = *reinterpret_cast<const float4*>(src + (long long)s.src[r] * TD + c4 * 4);
      if (drop) {
        float f[4];
        drop4(*drop, (unsigned long long)s.src[r] * 16 + c4, f);
        v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
      }
    }
    *reinterpret_cast<float4*>(tile + r * TLD + c4 * 4) = v;
  }
}
// tile rows -> rows of a [B, L, 64] tensor (only the bin's rows), optional dropout factor (the backward of it)
__device__ __forceinline__ void scatter_rows(const TrainSmem& s, float* __restrict__ dst, const float* tile,
                                             const DropCfg* drop) {
  for (int e = threadIdx.x; e < s.n * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    float4 v = *reinterpret_cast<const float4*>(tile + r * TLD + c4 * 4);
    if (drop) {
      float f[4];
      drop4(*drop, (unsigned long long)s.src[r] * 16 + c4, f);
      v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
    }
    *reinterpret_cast<float4*>(dst + (long long)s.src[r] * TD + c4 * 4) = v;
  }
}

// in-place elementwise pass over a tile: fn(r, c4, float4&) for r < rows
template <class F>
__device__ __forceinline__ void tile_pass(float* tile, int rows, F fn) {
  for (int e = threadIdx.x; e < rows * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    float4 v = *reinterpret_cast<float4*>(tile + r * TLD + c4 * 4);
    fn(r, c4, v);
    *reinterpret_cast<float4*>(tile + r * TLD + c4 * 4) = v;
  }
}

// Row passes give every warp TRW = 4 consecutive rows and run them TOGETHER (independent shuffle / exp / Philox
// chains interleave), instead of one row after the other: with 16 warps per SM these passes are latency-bound.
constexpr int TRW = TR / TWARPS;

template <int NR>
__device__ __forceinline__ void warp_sum_n(float (&v)[NR]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < NR; ++q) v[q] += __shfl_xor_sync(kFull, v[q], o);
}
template <int NR>
__device__ __forceinline__ void warp_max_n(float (&v)[NR]) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int q = 0; q < NR; ++q) v[q] = fmaxf(v[q], __shfl_xor_sync(kFull, v[q], o));
}

// LayerNorm of the bin's rows (same arithmetic as layernorm_fwd_kernel)
__device__ __forceinline__ void ln_rows(float* out, const float* in, int rows, const float* __restrict__ gamma,
                                        const float* __restrict__ beta) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r0 = w * TRW;
  if (r0 >= rows) return;
  const float g0 = gamma[lane], g1 = gamma[lane + 32], b0 = beta[lane], b1 = beta[lane + 32];
  float x0[TRW], x1[TRW], t[TRW];
#pragma unroll
  for (int q = 0; q < TRW; ++q) {
    x0[q] = in[(r0 + q) * TLD + lane];
    x1[q] = in[(r0 + q) * TLD + lane + 32];
    t[q] = x0[q] + x1[q];
  }
  warp_sum_n(t);
#pragma unroll
  for (int q = 0; q < TRW; ++q) {
    const float mean = t[q] / (float)TD;
    x0[q] -= mean;
    x1[q] -= mean;
    t[q] = fmaf(x0[q], x0[q], x1[q] * x1[q]);
  }
  warp_sum_n(t);
#pragma unroll
  for (int q = 0; q < TRW; ++q) {
    const float rstd = 1.0f / sqrtf(t[q] / (float)TD + kLnEps);
    if (r0 + q < rows) {
      out[(r0 + q) * TLD + lane] = x0[q] * rstd * g0 + b0;
      out[(r0 + q) * TLD + lane + 32] = x1[q] * rstd * g1 + b1;
    }
  }
}

// LayerNorm backward over the bin's rows: dx (=|+=) ..., dgamma/dbeta accumulated with atomics.
// dy and x are tiles; dx may alias dy.
__device__ __forceinline__ void ln_bwd_rows(TrainSmem& s, float* dx, const float* dy, const float* x, int rows,
                                            const float* __restrict__ gamma, float* __restrict__ dgamma,
                                            float* __restrict__ dbeta, bool accumulate) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int r0 = w * TRW;
  const float g0 = gamma[lane], g1 = gamma[lane + 32];
  float ag0 = 0.f, ag1 = 0.f, ab0 = 0.f, ab1 = 0.f;
  if (r0 < rows) {
    float h0[TRW], h1[TRW], d0[TRW], d1[TRW], t[TRW], u[TRW], rstd[TRW];
#pragma unroll
    for (int q = 0; q < TRW; ++q) {
      const bool live = r0 + q < rows;
      h0[q] = live ? x[(r0 + q) * TLD + lane] : 0.f;
      h1[q] = live ? x[(r0 + q) * TLD + lane + 32] : 0.f;
      d0[q] = live ? dy[(r0 + q) * TLD + lane] : 0.f;
      d1[q] = live ? dy[(r0 + q) * TLD + lane + 32] : 0.f;
      t[q] = h0[q] + h1[q];
    }
    warp_sum_n(t);
#pragma unroll
    for (int q = 0; q < TRW; ++q) {
      const float mean = t[q] / (float)TD;
      h0[q] -= mean;
      h1[q] -= mean;
      t[q] = fmaf(h0[q], h0[q], h1[q] * h1[q]);
    }
    warp_sum_n(t);
#pragma unroll
    for (int q = 0; q < TRW; ++q) {
      rstd[q] = 1.0f / sqrtf(t[q] / (float)TD + kLnEps);
      h0[q] *= rstd[q];
      h1[q] *= rstd[q];
      t[q] = d0[q] * g0 + d1[q] * g1;
      u[q] = fmaf(d0[q] * g0, h0[q], d1[q] * g1 * h1[q]);
      ag0 = fmaf(d0[q], h0[q], ag0); ag1 = fmaf(d1[q], h1[q], ag1);
      ab0 += d0[q]; ab1 += d1[q];
    }
    warp_sum_n(t);
    warp_sum_n(u);
#pragma unroll
    for (int q = 0; q < TRW; ++q) {
      if (r0 + q < rows) {
        const float m1 = t[q] / (float)TD, m2 = u[q] / (float)TD;
        const float v0 = rstd[q] * (d0[q] * g0 - m1 - h0[q] * m2), v1 = rstd[q] * (d1[q] * g1 - m1 - h1[q] * m2);
        float* o = dx + (r0 + q) * TLD;
        o[lane] = accumulate ? o[lane] + v0 : v0;
        o[lane + 32] = accumulate ? o[lane + 32] + v1 : v1;
      }
    }
  }
  s.red[w][0][lane] = ag0; s.red[w][0][lane + 32] = ag1;
  s.red[w][1][lane] = ab0; s.red[w][1][lane + 32] = ab1;
  __syncthreads();
  if (threadIdx.x < 2 * TD) {
    const int which = threadIdx.x / TD, c = threadIdx.x % TD;
    float acc = 0.f;
#pragma unroll
    for (int ww = 0; ww < TWARPS; ++ww) acc += s.red[ww][which][c];
    atomicAdd((which == 0 ? dgamma : dbeta) + c, acc);
  }
  __syncthreads();
}

// db[c] += sum_r tile[r][c]
__device__ __forceinline__ void colsum_tile(TrainSmem& s, const float* tile, int rows, float* __restrict__ db) {
  constexpr int PARTS = TTHREADS / TD;                      // row slices
  static_assert(PARTS <= TWARPS, "red[] holds one row of partials per slice");
  const int c = threadIdx.x % TD, part = threadIdx.x / TD;
  float acc = 0.f;
  for (int r = part; r < rows; r += PARTS) acc += tile[r * TLD + c];
  s.red[part][0][c] = acc;
  __syncthreads();
  if (threadIdx.x < TD) {
    float t = 0.f;
#pragma unroll
    for (int q = 0; q < PARTS; ++q) t += s.red[q][0][c];
    atomicAdd(db + c, t);
  }
  __syncthreads();
}

// dW[n][k] += sum_r dY[r][n] X[r][k]  (atomics into the gradient tensor), db[n] += colsum(dY)
__device__ __forceinline__ void weight_grad(TrainSmem& s, const float* dY, const float* X, int rows,
                                            float* __restrict__ dW, float* __restrict__ db, int ldw = TD) {
  gemm_tn<TD>(dY, TLD, X, TLD,
              [&](int, int, int& lo, int& hi) { lo = 0; hi = rows; },
              [&](int n, int c, const float (&v)[4]) {
#pragma unroll
                for (int j = 0; j < 4; ++j) atomicAdd(dW + (long long)n * ldw + c + j, v[j]);
              });
  if (db) colsum_tile(s, dY, rows, db);
}

struct AttnCfg {
  int H, dh;
  float sqrt_dh;
  bool cross;        // false: self (keys j <= i), true: decoder in training (keys j < i)
  int qbit;          // info bit that validates a query row
  int L;
  DropCfg drop;
};

// rows' key range: keys of the segment up to the row itself
__device__ __forceinline__ void key_range(const TrainSmem& s, int r0, int r1, int& lo, int& hi) {
  lo = s.info[r0] & 0x7f;
  hi = min(r1 + 1, s.n);
  if (r0 >= s.n) { lo = 0; hi = 0; }
}
// keys' query range: from the key itself to the last row of its segment
__device__ __forceinline__ void query_range(const TrainSmem& s, int j0, int j1, int& lo, int& hi) {
  lo = j0;
  hi = j1 < s.n ? ((s.info[j1] >> 8) & 0x7f) + 1 : s.n;
  if (j0 >= s.n) { lo = 0; hi = 0; }
}

__device__ __forceinline__ bool attn_need(const TrainSmem& s, int r0, int j) {
  // rows r0 .. r0+TNR-1: any of them may attend key j?
  if (r0 >= s.n || j >= s.n) return false;
  const int r1 = min(r0 + TNR - 1, s.n - 1);
  return j <= r1 && j >= (s.info[r0] & 0x7f);
}

// softmax of rows i0 .. i0+TRW-1 of S (tile, logits before scaling) restricted to the allowed keys; returns p for
// the lane's two keys (lane, lane + 32) of every row and the keep-scales of the attention dropout
__device__ __forceinline__ void softmax_rows(const TrainSmem& s, const AttnCfg& c, const float* S, int i0, int h, int lane,
                                             float (&p)[TRW][2], float (&f)[TRW][2]) {
  float sc[TRW][2], mx[TRW], sum[TRW];
#pragma unroll
  for (int r = 0; r < TRW; ++r) {
    const int i = i0 + r;
    const int info = s.info[i];
    const int seg0 = info & 0x7f;
    const bool qv = (info & c.qbit) != 0;
    const int jmax = c.cross ? i - 1 : i;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q;
      const bool ok = qv && j >= seg0 && j <= jmax && (s.info[j] & TI_PVALID);
      sc[r][q] = ok ? S[i * TLD + j] / c.sqrt_dh : -INFINITY;
    }
    mx[r] = fmaxf(sc[r][0], sc[r][1]);
  }
  warp_max_n(mx);
#pragma unroll
  for (int r = 0; r < TRW; ++r) {
#pragma unroll
    for (int q = 0; q < 2; ++q) p[r][q] = sc[r][q] == -INFINITY ? 0.f : expf(sc[r][q] - mx[r]);
    sum[r] = p[r][0] + p[r][1];
  }
  warp_sum_n(sum);
#pragma unroll
  for (int r = 0; r < TRW; ++r) {
    const int i = i0 + r;
    const float inv = sum[r] > 0.f ? 1.0f / sum[r] : 0.f;
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int j = lane + 32 * q;
      p[r][q] *= inv;
      f[r][q] = 1.0f;
      if (c.drop.p > 0.f) {
        const unsigned long long elem =
            (((unsigned long long)s.usr[i] * c.H + h) * c.L + s.pos[i]) * c.L + s.pos[j];
        f[r][q] = drop_factor(c.drop, elem);
      }
    }
  }
}

// out[:, head columns] (=|+=) dropout(softmax(mask(Q K^T))) V for every head.  Sb: scratch tile.
__device__ __forceinline__ void attention_fwd_tile(TrainSmem& s, const AttnCfg& c, float* out, const float* Q,
                                                   const float* K, const float* V, float* Sb, bool accumulate) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  for (int h = 0; h < c.H; ++h) {
    const int hc = h * c.dh;
    gemm_nt<TR>(Q + hc, TLD, K + hc, TLD, c.dh, s.npad, [&](int r0, int j) { return attn_need(s, r0, j); },
                [&](int r, int j, float v) { Sb[r * TLD + j] = v; });
    __syncthreads();
    if (w * TRW < s.npad) {
      float p[TRW][2], f[TRW][2];
      softmax_rows(s, c, Sb, w * TRW, h, lane, p, f);
#pragma unroll
      for (int r = 0; r < TRW; ++r) {
        Sb[(w * TRW + r) * TLD + lane] = p[r][0] * f[r][0];
        Sb[(w * TRW + r) * TLD + lane + 32] = p[r][1] * f[r][1];
      }
    }
    __syncthreads();
    auto range = [&](int r0, int r1, int& lo, int& hi) { key_range(s, r0, r1, lo, hi); };
    auto epi = [&](int r, int cc, const float (&v)[4]) {
      float4* o = reinterpret_cast<float4*>(out + r * TLD + hc + cc);
      float4 t = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
      t.x += v[0]; t.y += v[1]; t.z += v[2]; t.w += v[3];
      *o = t;
    };
    if (c.dh == 64) gemm_nn<64>(Sb, TLD, V + hc, TLD, s.npad, range, epi);
    else if (c.dh == 32) gemm_nn<32>(Sb, TLD, V + hc, TLD, s.npad, range, epi);
    else gemm_nn<16>(Sb, TLD, V + hc, TLD, s.npad, range, epi);
    __syncthreads();
  }
}

// Backward of attention_fwd_tile.  dQ is overwritten (head columns), dK / dV are accumulated when
// `acc_kv`, else overwritten.  Sb, Db: two scratch tiles.
__device__ __forceinline__ void attention_bwd_tile(TrainSmem& s, const AttnCfg& c, float* dQ, float* dK, float* dV,
                                                   const float* dO, const float* Q, const float* K, const float* V,
                                                   float* Sb, float* Db, bool acc_kv) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  for (int h = 0; h < c.H; ++h) {
    const int hc = h * c.dh;
    auto need = [&](int r0, int j) { return attn_need(s, r0, j); };
    gemm_nt<TR>(Q + hc, TLD, K + hc, TLD, c.dh, s.npad, need, [&](int r, int j, float v) { Sb[r * TLD + j] = v; });
    gemm_nt<TR>(dO + hc, TLD, V + hc, TLD, c.dh, s.npad, need, [&](int r, int j, float v) { Db[r * TLD + j] = v; });
    __syncthreads();
    {
      float p[TRW][2], f[TRW][2], dp[TRW][2], D[TRW];
      softmax_rows(s, c, Sb, w * TRW, h, lane, p, f);      // rows beyond the bin's fill come out as zeros
#pragma unroll
      for (int r = 0; r < TRW; ++r) {
        const int i = w * TRW + r;
        D[r] = 0.f;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          dp[r][q] = p[r][q] != 0.f ? Db[i * TLD + lane + 32 * q] * f[r][q] : 0.f;
          D[r] = fmaf(p[r][q], dp[r][q], D[r]);
        }
      }
      warp_sum_n(D);
#pragma unroll
      for (int r = 0; r < TRW; ++r) {
        const int i = w * TRW + r;
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          Db[i * TLD + lane + 32 * q] = p[r][q] * (dp[r][q] - D[r]) / c.sqrt_dh;   // d(Q K^T)
          Sb[i * TLD + lane + 32 * q] = p[r][q] * f[r][q];                          // dropout(W)
        }
      }
    }
    __syncthreads();
    auto krange = [&](int r0, int r1, int& lo, int& hi) { key_range(s, r0, r1, lo, hi); };
    auto qrange = [&](int j0, int j1, int& lo, int& hi) { query_range(s, j0, j1, lo, hi); };
    auto set_q = [&](int r, int cc, const float (&v)[4]) {
      *reinterpret_cast<float4*>(dQ + r * TLD + hc + cc) = make_float4(v[0], v[1], v[2], v[3]);
    };
    auto put = [&](float* dst) {
      return [=](int j, int cc, const float (&v)[4]) {
        float4* o = reinterpret_cast<float4*>(dst + j * TLD + hc + cc);
        float4 t = acc_kv ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
        t.x += v[0]; t.y += v[1]; t.z += v[2]; t.w += v[3];
        *o = t;
      };
    };
    if (c.dh == 64) {
      gemm_nn<64>(Db, TLD, K + hc, TLD, TR, krange, set_q);
      gemm_tn<64>(Db, TLD, Q + hc, TLD, qrange, put(dK));
      gemm_tn<64>(Sb, TLD, dO + hc, TLD, qrange, put(dV));
    } else if (c.dh == 32) {
      gemm_nn<32>(Db, TLD, K + hc, TLD, TR, krange, set_q);
      gemm_tn<32>(Db, TLD, Q + hc, TLD, qrange, put(dK));
      gemm_tn<32>(Sb, TLD, dO + hc, TLD, qrange, put(dV));
    } else {
      gemm_nn<16>(Db, TLD, K + hc, TLD, TR, krange, set_q);
      gemm_tn<16>(Db, TLD, Q + hc, TLD, qrange, put(dK));
      gemm_tn<16>(Sb, TLD, dO + hc, TLD, qrange, put(dV));
    }
    __syncthreads();
  }
}

// Y = X W^T + b into a tile (all 64 columns)
__device__ __forceinline__ void project(TrainSmem& s, float* Y, const float* X, const float* __restrict__ Wg,
                                        const float* __restrict__ bias) {
  load_w(s, Wg);
  __syncthreads();
  gemm_nt<TD>(X, TLD, s.w, TLD, TD, s.npad, NeedAll(),
              [&](int r, int n, float v) { Y[r * TLD + n] = v + bias[n]; });
  __syncthreads();
}

// dX (=|+=) dY W into a tile
__device__ __forceinline__ void project_bwd(TrainSmem& s, float* dX, const float* dY, const float* __restrict__ Wg,
                                            bool accumulate) {
  load_w(s, Wg);
  __syncthreads();
  gemm_nn<TD>(dY, TLD, s.w, TLD, s.npad,
              [&](int, int, int& lo, int& hi) { lo = 0; hi = TD; },
              [&](int r, int cc, const float (&v)[4]) {
                float4* o = reinterpret_cast<float4*>(dX + r * TLD + cc);
                float4 t = accumulate ? *o : make_float4(0.f, 0.f, 0.f, 0.f);
                t.x += v[0]; t.y += v[1]; t.z += v[2]; t.w += v[3];
                *o = t;
              });
  __syncthreads();
}

__device__ __forceinline__ void load_bin(const TrainArgs& a, TrainSmem& s, int bin) {
  __syncthreads();
  if (threadIdx.x < TR) {
    const int r = threadIdx.x;
    const int src = a.row_src[(long long)bin * TR + r];
    s.src[r] = src;
    s.info[r] = src >= 0 ? a.row_info[(long long)bin * TR + r] : 0;
    s.pos[r] = src >= 0 ? src % a.L : 0;
    s.usr[r] = src >= 0 ? src / a.L : 0;
    if (a.embed_mode) {
      s.ids[0][r] = src >= 0 ? a.p_x[src] : 0;
      for (int t = 0; t < a.n_tuples; ++t) s.ids[1 + t][r] = src >= 0 ? a.o_x[t][src] : 0;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int n = 0;
    while (n < TR && s.src[n] >= 0) ++n;
    s.n = n;
    s.npad = (n + 3) & ~3;
  }
  __syncthreads();
}

__device__ __forceinline__ AttnCfg self_cfg(const TrainArgs& a, int b) {
  AttnCfg c;
  c.H = a.H;
  c.dh = TD / a.H;
  c.sqrt_dh = sqrtf((float)c.dh);
  c.cross = false;
  c.qbit = TI_PVALID;
  c.L = a.L;
  c.drop = at_site(a.drop, 1u + 3u * (unsigned)b);
  return c;
}
__device__ __forceinline__ AttnCfg cross_cfg(const TrainArgs& a, int t) {
  AttnCfg c = self_cfg(a, 0);
  c.cross = true;
  c.qbit = TI_TVALID0 << t;
  c.drop = at_site(a.drop, 1000u + (unsigned)t);
  return c;
}

// y = sigmoid(<s_row, wf> + bf) for the bin's rows (warp per row); also usable to recompute y in the backward
__device__ __forceinline__ float row_dot(const float* a, const float* __restrict__ b, int lane) {
  return warp_sum(fmaf(a[lane], b[lane], a[lane + 32] * b[lane + 32]));
}

// -------------------------------------------------------------------------------------------------- folded embedding
// Z[r] = sqrt(d) E[id_r] (zero rows for padding ids and below the bin's fill)
__device__ __forceinline__ void gather_items(const TrainArgs& a, const TrainSmem& s, const int* ids, float* Z) {
  for (int e = threadIdx.x; e < TR * 16; e += TTHREADS) {
    const int r = e >> 4, c4 = e & 15;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    const int id = ids[r];
    if (id != 0) {
      v = *reinterpret_cast<const float4*>(a.E + (long long)id * TD + c4 * 4);
      v.x *= a.sqrt_d; v.y *= a.sqrt_d; v.z *= a.sqrt_d; v.w *= a.sqrt_d;
    }
    *reinterpret_cast<float4*>(Z + r * TLD + c4 * 4) = v;
  }
}

// context rows of the folded table and cst -> shared memory (once per CTA)
__device__ __forceinline__ void load_fold_consts(const TrainArgs& a, TrainSmem& s) {
  for (int e = threadIdx.x; e < (TMAXC + 1) * TD; e += TTHREADS) {
    const int k = e / TD, c = e % TD;
    float v = 0.f;
    if (k < a.C) v = a.fold[(long long)(a.A + k) * TD + c];
    else if (k == TMAXC) v = a.fold[(long long)(a.A + a.C) * TD + c];
    s.gtc[k][c] = v;
  }
}

// context values of the bin's rows -> shared memory (zeros for padding ids / unused slots)
__device__ __forceinline__ void stage_ctx(const TrainArgs& a, TrainSmem& s, const int* ids, const float* __restrict__ ctx) {
  for (int e = threadIdx.x; e < TR * TMAXC; e += TTHREADS) {
    const int r = e / TMAXC, k = e % TMAXC;
    float v = 0.f;
    if (r < s.n && k < a.C && ids[r] != 0) v = ctx[(long long)s.src[r] * a.C + k];
    s.ctxv[r][k] = v;
  }
}

// AllEmbedding.forward (src/carca.py:85-95) of one row set through the folded tables.  Eo: output tile,
// Z: scratch tile.  Rows with id 0 come out exactly 0 (the final `* mask`, :94).
// The attribute gather runs one row per HALF-warp (16 lanes x 128 bits = one 64-float table row per load), the
// row's column ids fetched by the lanes at once and four table rows in flight per step; the loop bounds are
// made warp-uniform so both halves stay converged.
__device__ __forceinline__ void embed_rows(const TrainArgs& a, TrainSmem& s, const int* ids, const float* __restrict__ ctx,
                                           bool add_pos, float* Eo, float* Z) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int half = lane >> 4, hl = lane & 15;
  gather_items(a, s, ids, Z);
  load_w(s, a.Wj, a.ldj);
  stage_ctx(a, s, ids, ctx);
  __syncthreads();
  gemm_nt<TD>(Z, TLD, s.w, TLD, TD, s.npad, NeedAll(), [&](int r, int n, float v) { Eo[r * TLD + n] = v; });
  __syncthreads();
  for (int r0 = 2 * w; r0 < s.npad; r0 += 2 * TWARPS) {
    const int r = r0 + half;
    const int id = ids[r];
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    int k0 = 0, nnz = 0;
    if (id != 0) {
      k0 = a.csr_rowptr[id];
      nnz = a.csr_rowptr[id + 1] - k0;
      v = *reinterpret_cast<const float4*>(Eo + r * TLD + hl * 4);
      const float4 c4 = *reinterpret_cast<const float4*>(&s.gtc[TMAXC][hl * 4]);
      v.x += c4.x; v.y += c4.y; v.z += c4.z; v.w += c4.w;
    }
    const int nmax = max(nnz, __shfl_xor_sync(kFull, nnz, 16));
    for (int kb = 0; kb < nmax; kb += 16) {
      int col = 0;
      float val = 0.f;
      if (kb + hl < nnz) {
        col = a.csr_cols[k0 + kb + hl];
        val = a.csr_vals[k0 + kb + hl];
      }
      const int mm = min(16, nmax - kb);
      for (int q = 0; q < mm; q += 4) {
        float4 g[4];
        float vv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const int from = (lane & 16) | ((q + u) & 15);
          const int cq = __shfl_sync(kFull, col, from);
          vv[u] = __shfl_sync(kFull, val, from);
          g[u] = *reinterpret_cast<const float4*>(a.fold + (long long)cq * TD + hl * 4);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          v.x = fmaf(vv[u], g[u].x, v.x);
          v.y = fmaf(vv[u], g[u].y, v.y);
          v.z = fmaf(vv[u], g[u].z, v.z);
          v.w = fmaf(vv[u], g[u].w, v.w);
        }
      }
    }
    if (id != 0) {
#pragma unroll
      for (int k = 0; k < TMAXC; ++k) {
        const float cv = s.ctxv[r][k];
        const float4 g = *reinterpret_cast<const float4*>(&s.gtc[k][hl * 4]);
        v.x = fmaf(cv, g.x, v.x);
        v.y = fmaf(cv, g.y, v.y);
        v.z = fmaf(cv, g.z, v.z);
        v.w = fmaf(cv, g.w, v.w);
      }
      if (add_pos) {
        const float4 pv = *reinterpret_cast<const float4*>(a.pos_table + (long long)s.pos[r] * TD + hl * 4);
        v.x += pv.x; v.y += pv.y; v.z += pv.z; v.w += pv.w;
      }
    }
    *reinterpret_cast<float4*>(Eo + r * TLD + hl * 4) = v;
  }
  __syncthreads();
}

// Backward of embed_rows for the row gradients DE (tile, modified in place): accumulates d fold (GT rows, context
// rows, cst), d pos, d joint_embed.weight[:, :64] and d items_embed.  Z, T2: scratch tiles.
__device__ __forceinline__ void embed_rows_bwd(const TrainArgs& a, TrainSmem& s, const int* ids,
                                               const float* __restrict__ ctx, bool is_profile, float* DE, float* Z,
                                               float* T2) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int half = lane >> 4, hl = lane & 15;
  const int n = s.n;
  tile_pass(DE, TR, [&](int r, int, float4& v) {                     // the `* mask` of :94
    if (r >= n || ids[r] == 0) v = make_float4(0.f, 0.f, 0.f, 0.f);
  });
  gather_items(a, s, ids, Z);
  stage_ctx(a, s, ids, ctx);
  __syncthreads();
  colsum_tile(s, DE, n, a.dfold + (long long)(a.A + a.C) * TD);      // d cst
  for (int e = threadIdx.x; e < a.C * TD; e += TTHREADS) {            // d (context rows of GT)
    const int k = e / TD, c = e % TD;
    float acc = 0.f;
    for (int r = 0; r < n; ++r) acc = fmaf(s.ctxv[r][k], DE[r * TLD + c], acc);
    atomicAdd(a.dfold + (long long)(a.A + k) * TD + c, acc);
  }
  for (int r0 = 2 * w; r0 < s.npad; r0 += 2 * TWARPS) {              // d (attribute rows of GT), d pos
    const int r = r0 + half;
    const int id = ids[r];
    // lane hl of the half-warp owns features hl, hl+16, hl+32, hl+48: each RED instruction of the half-warp then
    // covers 16 consecutive floats (two sectors) instead of 16 sectors
    float dv[4] = {0.f, 0.f, 0.f, 0.f};
    int k0 = 0, nnz = 0;
    if (id != 0) {
      k0 = a.csr_rowptr[id];
      nnz = a.csr_rowptr[id + 1] - k0;
#pragma unroll
      for (int j = 0; j < 4; ++j) dv[j] = DE[r * TLD + hl + 16 * j];
    }
    const int nmax = max(nnz, __shfl_xor_sync(kFull, nnz, 16));
    for (int kb = 0; kb < nmax; kb += 16) {
      int col = 0;
      float val = 0.f;
      if (kb + hl < nnz) {
        col = a.csr_cols[k0 + kb + hl];
        val = a.csr_vals[k0 + kb + hl];
      }
      const int mm = min(16, nmax - kb);
      for (int q = 0; q < mm; ++q) {
        const int from = (lane & 16) | q;
        const int cq = __shfl_sync(kFull, col, from);
        const float vq = __shfl_sync(kFull, val, from);
        if (kb + q < nnz) {
          float* g = a.dfold + (long long)cq * TD + hl;
#pragma unroll
          for (int j = 0; j < 4; ++j) atomicAdd(g + 16 * j, vq * dv[j]);
        }
      }
    }
    if (id != 0 && is_profile && a.gpos) {
      float* g = a.gpos + (long long)s.pos[r] * TD + hl;
#pragma unroll
      for (int j = 0; j < 4; ++j) atomicAdd(g + 16 * j, dv[j]);
    }
  }
  weight_grad(s, DE, Z, n, a.gWj, nullptr, a.ldj);                    // d Wj[:, :64]
  __syncthreads();
  load_w(s, a.Wj, a.ldj);
  __syncthreads();
  gemm_nn<TD>(DE, TLD, s.w, TLD, s.npad,
              [&](int, int, int& lo, int& hi) { lo = 0; hi = TD; },
              [&](int r, int cc, const float (&v)[4]) {
                *reinterpret_cast<float4*>(T2 + r * TLD + cc) = make_float4(v[0], v[1], v[2], v[3]);
              });
  __syncthreads();
  for (int e = threadIdx.x; e < n * TD; e += TTHREADS) {              // d items_embed (padding_idx row gets none)
    const int r = e / TD, c = e % TD;
    const int id = ids[r];
    if (id != 0) atomicAdd(a.gE + (long long)id * TD + c, a.sqrt_d * T2[r * TLD + c]);
  }
  __syncthreads();
}

// -------------------------------------------------------------------------------------------------- forward
__global__ void __launch_bounds__(TTHREADS, 1) fused_train_fwd_kernel(const TrainArgs a) {
  CARCA_DYN_SMEM(unsigned char, raw);
  TrainSmem& s = *reinterpret_cast<TrainSmem*>(raw);
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* X = s.buf[0];
  float* QN = s.buf[1];
  float* Qb = s.buf[2];
  float* Kb = s.buf[3];
  float* Vb = s.buf[4];
  float* Sb = s.buf[5];
  for (int e = threadIdx.x; e < 6 * TBUF; e += TTHREADS) s.buf[0][e] = 0.f;
  if (a.embed_mode) load_fold_consts(a, s);
  const int nbins = *a.n_bins;
  TrainTicks tk = ticks_begin(a);
  for (int bin = blockIdx.x; bin < nbins; bin += gridDim.x) {
    tick(tk, 0);
    load_bin(a, s, bin);
    const int n = s.n;
    tick(tk, 1);
    {  // x0 = dropout(p_e)   (src/carca.py:415-416, site 0)
      const DropCfg d0 = at_site(a.drop, 0u);
      if (a.embed_mode) {
        embed_rows(a, s, s.ids[0], a.p_c, a.pos_table != nullptr, X, QN);
        if (a.drop.p > 0.f) {
          tile_pass(X, n, [&](int r, int c4, float4& v) {
            float f[4];
            drop4(d0, (unsigned long long)s.src[r] * 16 + c4, f);
            v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
          });
        }
      } else {
        gather_rows(s, a.p_e, X, a.drop.p > 0.f ? &d0 : nullptr);
      }
    }
    __syncthreads();
    tick(tk, 2);
    for (int b = 0; b < a.n_blocks; ++b) {
      const TrainBlockW& wb = a.blk[b];
      save_tile(a, s, sv_block(b, 0), bin, X);
      ln_rows(QN, X, n, wb.ln1_g, wb.ln1_b);                          // :298
      __syncthreads();
      tick(tk, 3);
      project(s, Qb, QN, wb.wq, wb.bq);                               // :238 (query = LN1(x))
      project(s, Kb, X, wb.wk, wb.bk);                                // :239 (key = raw x)
      project(s, Vb, X, wb.wv, wb.bv);                                // :240
      tick(tk, 4);
      save_tile(a, s, sv_block(b, 1), bin, Qb);
      save_tile(a, s, sv_block(b, 2), bin, Kb);
      save_tile(a, s, sv_block(b, 3), bin, Vb);
      tick(tk, 5);
      // s = MHA(..., causal=0) (+ LN1(x)), in place over QN (:299-302)
      attention_fwd_tile(s, self_cfg(a, b), QN, Qb, Kb, Vb, Sb, a.residual_sa != 0);
      tick(tk, 6);
      save_tile(a, s, sv_block(b, 4), bin, QN);
      ln_rows(Qb, QN, n, wb.ln2_g, wb.ln2_b);                         // s2 (:304)
      __syncthreads();
      tick(tk, 7);
      // a1 = dropout(LeakyReLU(ffn_1(s2)))  (:307-309, site 2 + 3b)
      project(s, Kb, Qb, wb.w1, wb.b1);
      {
        const DropCfg d1 = at_site(a.drop, 2u + 3u * (unsigned)b);
        tile_pass(Kb, n, [&](int r, int c4, float4& v) {
          float f[4];
          drop4(d1, (unsigned long long)s.src[r] * 16 + c4, f);
          v.x = (v.x > 0.f ? v.x : kLeakySlope * v.x) * f[0];
          v.y = (v.y > 0.f ? v.y : kLeakySlope * v.y) * f[1];
          v.z = (v.z > 0.f ? v.z : kLeakySlope * v.z) * f[2];
          v.w = (v.w > 0.f ? v.w : kLeakySlope * v.w) * f[3];
        });
      }
      __syncthreads();
      save_tile(a, s, sv_block(b, 5), bin, Kb);
      // x' = dropout(ffn_2(a1)) (+ s2)  (:311-316, site 3 + 3b)
      project(s, X, Kb, wb.w2, wb.b2);
      {
        const DropCfg d2 = at_site(a.drop, 3u + 3u * (unsigned)b);
        const bool res = a.residual_sa != 0;
        tile_pass(X, n, [&](int r, int c4, float4& v) {
          float f[4];
          drop4(d2, (unsigned long long)s.src[r] * 16 + c4, f);
          const float4 q = *reinterpret_cast<const float4*>(Qb + r * TLD + c4 * 4);
          v.x = v.x * f[0] + (res ? q.x : 0.f);
          v.y = v.y * f[1] + (res ? q.y : 0.f);
          v.z = v.z * f[2] + (res ? q.z : 0.f);
          v.w = v.w * f[3] + (res ? q.w : 0.f);
        });
      }
      __syncthreads();
      tick(tk, 8);
    }
    save_tile(a, s, sv_final(a.n_blocks), bin, X);
    float* PE = QN;
    ln_rows(PE, X, n, a.fn_g, a.fn_b);                                // :421
    __syncthreads();
    tick(tk, 9);
    if (a.decoder == 1) {
      project(s, Kb, PE, a.dec.wk, a.dec.bk);
      project(s, Vb, PE, a.dec.wv, a.dec.bv);
      save_tile(a, s, sv_dec(a.n_blocks, 0), bin, Kb);
      save_tile(a, s, sv_dec(a.n_blocks, 1), bin, Vb);
    }
    tick(tk, 10);
    for (int t = 0; t < a.n_tuples; ++t) {
      float* O = X;
      if (a.embed_mode) {
        embed_rows(a, s, s.ids[1 + t], a.o_c[t], false, O, Qb);       // :426
        save_tile(a, s, sv_dec(a.n_blocks, 2 + a.n_tuples + t), bin, O);
      } else {
        gather_rows(s, a.o_e[t], O, nullptr);                         // (embedded by the caller)
      }
      __syncthreads();
      tick(tk, 11);
      if (a.decoder == 1) {
        project(s, Qb, O, a.dec.wq, a.dec.bq);
        if (!a.residual_ca) {
          for (int e = threadIdx.x; e < TBUF; e += TTHREADS) O[e] = 0.f;
          __syncthreads();
        }
        tick(tk, 12);
        attention_fwd_tile(s, cross_cfg(a, t), O, Qb, Kb, Vb, Sb, true);   // :339-343 (causal -1)
        tick(tk, 13);
        save_tile(a, s, sv_dec(a.n_blocks, 2 + t), bin, O);
        const float bf = a.dec.bf[0];
        for (int r = w; r < n; r += TWARPS) {                         // :345-347
          const float z = row_dot(O + r * TLD, a.dec.wf, lane) + bf;
          if (lane == 0)
            a.y[(long long)s.usr[r] * a.ldy + t * a.L + s.pos[r]] = 1.0f / (1.0f + expf(-z));
        }
      } else {
        for (int r = w; r < n; r += TWARPS) {                         // :360
          const float z = row_dot(O + r * TLD, PE + r * TLD, lane);
          if (lane == 0)
            a.y[(long long)s.usr[r] * a.ldy + t * a.L + s.pos[r]] = 1.0f / (1.0f + expf(-z));
        }
      }
      __syncthreads();
      tick(tk, 14);
    }
  }
}

// -------------------------------------------------------------------------------------------------- backward
__global__ void __launch_bounds__(TTHREADS, 1) fused_train_bwd_kernel(const TrainArgs a) {
  CARCA_DYN_SMEM(unsigned char, raw);
  TrainSmem& s = *reinterpret_cast<TrainSmem*>(raw);
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* B0 = s.buf[0];
  float* B1 = s.buf[1];
  float* B2 = s.buf[2];
  float* B3 = s.buf[3];
  float* B4 = s.buf[4];
  float* B5 = s.buf[5];
  float* B6 = s.buf[6];
  float* B7 = s.buf[7];
  float* Sb = s.buf[8];
  float* Db = s.buf[9];
  for (int e = threadIdx.x; e < TNBUF * TBUF; e += TTHREADS) s.buf[0][e] = 0.f;
  if (a.embed_mode) load_fold_consts(a, s);
  const int nbins = *a.n_bins;
  const int nb = a.n_blocks;

  // d(ffn bias) of the cross-attention decoder from INACTIVE positions (y = sigmoid(bf) there); zero in the
  // reference loop (the loss mask kills dy), kept for exactness with any upstream gradient
  if (a.decoder == 1) {
    const float y0 = 1.0f / (1.0f + expf(-a.dec.bf[0]));
    float acc = 0.f;
    const long long total = (long long)a.B * a.L;
    for (long long e = (long long)blockIdx.x * TTHREADS + threadIdx.x; e < total; e += (long long)gridDim.x * TTHREADS) {
      bool act = a.p_x[e] != 0;
      for (int q = 0; q < a.n_tuples; ++q) act = act || a.o_x[q][e] != 0;
      if (!act) {
        const long long u = e / a.L, i = e % a.L;
        for (int q = 0; q < a.n_tuples; ++q) acc += a.dy[u * a.ldy + q * a.L + i];
      }
    }
    acc = warp_sum(acc);
    if (lane == 0 && acc != 0.f) atomicAdd(a.gdec.bf, acc * y0 * (1.0f - y0));
  }

  TrainTicks tk = ticks_begin(a);
  for (int bin = blockIdx.x; bin < nbins; bin += gridDim.x) {
    tick(tk, 100);
    load_bin(a, s, bin);
    const int n = s.n;
    float* DX = B0;   // gradient w.r.t. the output of the stage being processed

    // ---------------- decoder
    if (a.decoder == 1) {
      float* Kd = B3;
      float* Vd = B4;
      float* dKd = B5;
      float* dVd = B6;
      load_tile(a, s, sv_dec(nb, 0), bin, Kd);
      load_tile(a, s, sv_dec(nb, 1), bin, Vd);
      for (int t = 0; t < a.n_tuples; ++t) {
        float* St = B0;
        float* DS = B1;
        tick(tk, 101);
        load_tile(a, s, sv_dec(nb, 2 + t), bin, St);
        __syncthreads();
        // g = dy * y (1 - y); d wf += g s; d bf += g; ds = g wf   (:345-347)
        const float bf = a.dec.bf[0];
        float awf0 = 0.f, awf1 = 0.f, abf = 0.f;
        for (int r = w; r < TR; r += TWARPS) {
          float g = 0.f;
          if (r < n) {
            const float z = row_dot(St + r * TLD, a.dec.wf, lane) + bf;
            const float yv = 1.0f / (1.0f + expf(-z));
            g = a.dy[(long long)s.usr[r] * a.ldy + t * a.L + s.pos[r]] * yv * (1.0f - yv);
            awf0 = fmaf(g, St[r * TLD + lane], awf0);
            awf1 = fmaf(g, St[r * TLD + lane + 32], awf1);
            abf += g;
          }
          DS[r * TLD + lane] = g * a.dec.wf[lane];
          DS[r * TLD + lane + 32] = g * a.dec.wf[lane + 32];
        }
        s.red[w][0][lane] = awf0; s.red[w][0][lane + 32] = awf1;
        if (lane == 0) s.red[w][1][0] = abf;
        __syncthreads();
        if (threadIdx.x < TD) {
          float acc = 0.f;
          for (int ww = 0; ww < TWARPS; ++ww) acc += s.red[ww][0][threadIdx.x];
          atomicAdd(a.gdec.wf + threadIdx.x, acc);
        } else if (threadIdx.x == TD) {
          float acc = 0.f;
          for (int ww = 0; ww < TWARPS; ++ww) acc += s.red[ww][1][0];
          atomicAdd(a.gdec.bf, acc);
        }
        __syncthreads();
        // recompute Q of the targets
        float* O = B0;
        float* Qt = B2;
        float* dQ = B7;
        tick(tk, 102);
        if (a.embed_mode) load_tile(a, s, sv_dec(nb, 2 + a.n_tuples + t), bin, O);   // embedded in the forward
        else gather_rows(s, a.o_e[t], O, nullptr);
        __syncthreads();
        tick(tk, 103);
        project(s, Qt, O, a.dec.wq, a.dec.bq);
        tick(tk, 104);
        attention_bwd_tile(s, cross_cfg(a, t), dQ, dKd, dVd, DS, Qt, Kd, Vd, Sb, Db, t > 0);
        tick(tk, 105);
        weight_grad(s, dQ, O, n, a.gdec.wq, a.gdec.bq);
        __syncthreads();
        tick(tk, 106);
        // d o = dQ WQ (+ ds through the residual) -> d_oe rows
        if (a.residual_ca) project_bwd(s, DS, dQ, a.dec.wq, true);
        else project_bwd(s, DS, dQ, a.dec.wq, false);
        tick(tk, 107);
        if (a.embed_mode) embed_rows_bwd(a, s, s.ids[1 + t], a.o_c[t], false, DS, B0, B2);
        else scatter_rows(s, a.d_oe[t], DS, nullptr);
        __syncthreads();
        tick(tk, 108);
      }
      // keys / values: PE = LN_f(x_nb) recomputed
      float* XF = B1;
      float* PE = B2;
      load_tile(a, s, sv_final(nb), bin, XF);
      __syncthreads();
      ln_rows(PE, XF, n, a.fn_g, a.fn_b);
      __syncthreads();
      weight_grad(s, dKd, PE, n, a.gdec.wk, a.gdec.bk);
      weight_grad(s, dVd, PE, n, a.gdec.wv, a.gdec.bv);
      float* DPE = B7;
      project_bwd(s, DPE, dKd, a.dec.wk, false);
      project_bwd(s, DPE, dVd, a.dec.wv, true);
      // final LayerNorm backward -> DX
      ln_bwd_rows(s, DX, DPE, XF, n, a.fn_g, a.g_fn_g, a.g_fn_b, false);
      tick(tk, 109);
    } else {
      // dot product (:360): y = sigmoid(<pe_r, o_r>); d pe_r = sum_t g o_r; d o_r = g pe_r
      float* XF = B1;
      float* PE = B2;
      float* DPE = B7;
      load_tile(a, s, sv_final(nb), bin, XF);
      __syncthreads();
      ln_rows(PE, XF, n, a.fn_g, a.fn_b);
      for (int e = threadIdx.x; e < TBUF; e += TTHREADS) DPE[e] = 0.f;
      __syncthreads();
      for (int t = 0; t < a.n_tuples; ++t) {
        float* O = B3;
        if (a.embed_mode) load_tile(a, s, sv_dec(nb, 2 + a.n_tuples + t), bin, O);
        else gather_rows(s, a.o_e[t], O, nullptr);
        __syncthreads();
        for (int r = w; r < n; r += TWARPS) {
          const float z = row_dot(O + r * TLD, PE + r * TLD, lane);
          const float yv = 1.0f / (1.0f + expf(-z));
          const float g = a.dy[(long long)s.usr[r] * a.ldy + t * a.L + s.pos[r]] * yv * (1.0f - yv);
          const float o0 = O[r * TLD + lane], o1 = O[r * TLD + lane + 32];
          DPE[r * TLD + lane] = fmaf(g, o0, DPE[r * TLD + lane]);
          DPE[r * TLD + lane + 32] = fmaf(g, o1, DPE[r * TLD + lane + 32]);
          O[r * TLD + lane] = g * PE[r * TLD + lane];
          O[r * TLD + lane + 32] = g * PE[r * TLD + lane + 32];
        }
        __syncthreads();
        if (a.embed_mode) embed_rows_bwd(a, s, s.ids[1 + t], a.o_c[t], false, O, B4, B5);
        else scatter_rows(s, a.d_oe[t], O, nullptr);
        __syncthreads();
      }
      ln_bwd_rows(s, DX, DPE, XF, n, a.fn_g, a.g_fn_g, a.g_fn_b, false);
    }

    // ---------------- encoder blocks, last to first.  DX (B0) = d(block output)
    for (int b = nb - 1; b >= 0; --b) {
      const TrainBlockW& wb = a.blk[b];
      const TrainBlockG& gb = a.gblk[b];
      const bool res = a.residual_sa != 0;
      tick(tk, 110);
      // FFN (:305-316)
      float* DF2 = B1;   // d ffn_2 output = dout * dropout2
      {
        const DropCfg d2 = at_site(a.drop, 3u + 3u * (unsigned)b);
        for (int e = threadIdx.x; e < TR * 16; e += TTHREADS) {
          const int r = e >> 4, c4 = e & 15;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (r < n) {
            v = *reinterpret_cast<const float4*>(DX + r * TLD + c4 * 4);
            float f[4];
            drop4(d2, (unsigned long long)s.src[r] * 16 + c4, f);
            v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
          }
          *reinterpret_cast<float4*>(DF2 + r * TLD + c4 * 4) = v;
        }
      }
      float* A1 = B2;
      load_tile(a, s, sv_block(b, 5), bin, A1);
      __syncthreads();
      weight_grad(s, DF2, A1, n, gb.w2, gb.b2);
      float* DF1 = B3;
      project_bwd(s, DF1, DF2, wb.w2, false);                         // d a1
      {
        const DropCfg d1 = at_site(a.drop, 2u + 3u * (unsigned)b);
        tile_pass(DF1, n, [&](int r, int c4, float4& v) {             // through dropout1 and LeakyReLU
          float f[4];
          drop4(d1, (unsigned long long)s.src[r] * 16 + c4, f);
          const float4 q = *reinterpret_cast<const float4*>(A1 + r * TLD + c4 * 4);
          v.x *= f[0] * (q.x > 0.f ? 1.0f : kLeakySlope);
          v.y *= f[1] * (q.y > 0.f ? 1.0f : kLeakySlope);
          v.z *= f[2] * (q.z > 0.f ? 1.0f : kLeakySlope);
          v.w *= f[3] * (q.w > 0.f ? 1.0f : kLeakySlope);
        });
      }
      __syncthreads();
      // s2 = LN2(s) recomputed
      float* Sx = B4;    // s (LN2 input)
      float* S2 = B2;    // over A1 (no longer needed)
      load_tile(a, s, sv_block(b, 4), bin, Sx);
      __syncthreads();
      ln_rows(S2, Sx, n, wb.ln2_g, wb.ln2_b);
      __syncthreads();
      weight_grad(s, DF1, S2, n, gb.w1, gb.b1);
      float* DS2 = B1;   // over DF2
      project_bwd(s, DS2, DF1, wb.w1, false);
      if (res) {
        tile_pass(DS2, n, [&](int r, int c4, float4& v) {
          const float4 q = *reinterpret_cast<const float4*>(DX + r * TLD + c4 * 4);
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        });
        __syncthreads();
      }
      tick(tk, 111);
      // LN2 backward -> d s (B0)
      float* DSx = B0;
      ln_bwd_rows(s, DSx, DS2, Sx, n, wb.ln2_g, gb.ln2_g, gb.ln2_b, false);
      tick(tk, 112);
      // attention backward
      float* Qb = B1;
      float* Kb = B2;
      float* Vb = B3;
      float* dQ = B4;
      float* dK = B5;
      float* dV = B6;
      load_tile(a, s, sv_block(b, 1), bin, Qb);
      load_tile(a, s, sv_block(b, 2), bin, Kb);
      load_tile(a, s, sv_block(b, 3), bin, Vb);
      __syncthreads();
      tick(tk, 113);
      attention_bwd_tile(s, self_cfg(a, b), dQ, dK, dV, DSx, Qb, Kb, Vb, Sb, Db, false);
      tick(tk, 114);
      // x and qn = LN1(x) recomputed
      float* Xb = B1;
      float* QN = B2;
      load_tile(a, s, sv_block(b, 0), bin, Xb);
      __syncthreads();
      ln_rows(QN, Xb, n, wb.ln1_g, wb.ln1_b);
      __syncthreads();
      weight_grad(s, dQ, QN, n, gb.wq, gb.bq);
      weight_grad(s, dK, Xb, n, gb.wk, gb.bk);
      weight_grad(s, dV, Xb, n, gb.wv, gb.bv);
      tick(tk, 115);
      // d qn = dQ WQ (+ d s through the residual)
      float* DQN = B3;
      project_bwd(s, DQN, dQ, wb.wq, false);
      if (res) {
        tile_pass(DQN, n, [&](int r, int c4, float4& v) {
          const float4 q = *reinterpret_cast<const float4*>(DSx + r * TLD + c4 * 4);
          v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
        });
        __syncthreads();
      }
      // d x = dK WK + dV WV + LN1 backward(d qn)
      float* DXn = B7;
      project_bwd(s, DXn, dK, wb.wk, false);
      project_bwd(s, DXn, dV, wb.wv, true);
      ln_bwd_rows(s, DXn, DQN, Xb, n, wb.ln1_g, gb.ln1_g, gb.ln1_b, true);
      tick(tk, 116);
      for (int e = threadIdx.x; e < TR * 16; e += TTHREADS) {         // DX <- DXn
        const int r = e >> 4, c4 = e & 15;
        *reinterpret_cast<float4*>(DX + r * TLD + c4 * 4) = *reinterpret_cast<const float4*>(DXn + r * TLD + c4 * 4);
      }
      __syncthreads();
    }
    // d p_e = d x0 through the embedding dropout (site 0)
    {
      const DropCfg d0 = at_site(a.drop, 0u);
      if (a.embed_mode) {
        if (a.drop.p > 0.f) {
          tile_pass(DX, n, [&](int r, int c4, float4& v) {
            float f[4];
            drop4(d0, (unsigned long long)s.src[r] * 16 + c4, f);
            v.x *= f[0]; v.y *= f[1]; v.z *= f[2]; v.w *= f[3];
          });
          __syncthreads();
        }
        tick(tk, 117);
        embed_rows_bwd(a, s, s.ids[0], a.p_c, true, DX, B1, B2);
      } else {
        scatter_rows(s, a.d_pe, DX, a.drop.p > 0.f ? &d0 : nullptr);
      }
    }
    __syncthreads();
    tick(tk, 118);
  }
}

// -------------------------------------------------------------------------------------------------- fold helpers
// cst[c] = sum_g Wj[c][d + g] bf[g] + bj[c]: the constant row of the folded embedding tables (one warp per output)
__global__ void __launch_bounds__(256) fold_cst_kernel(float* __restrict__ cst, const float* __restrict__ Wj,
                                                       const float* __restrict__ bf, const float* __restrict__ bj,
                                                       int d, int g) {
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  for (int c = blockIdx.x * (blockDim.x / kWarp) + w; c < d; c += gridDim.x * (blockDim.x / kWarp)) {
    const float* row = Wj + (long long)c * (d + g) + d;
    float acc = 0.f;
    for (int k = lane; k < g; k += kWarp) acc = fmaf(row[k], bf[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) cst[c] = acc + bj[c];
  }
}

// Un-folds d cst: d bj = dcst; d bf[g] = sum_c dcst[c] Wj[c][d + g]; d Wj[c][d + g] += dcst[c] bf[g].
// Grid: d blocks (one per c); block 0 also writes d bf and d bj.
__global__ void __launch_bounds__(256) unfold_cst_kernel(float* __restrict__ gWj, float* __restrict__ gbf,
                                                         float* __restrict__ gbj, const float* __restrict__ dcst,
                                                         const float* __restrict__ Wj, const float* __restrict__ bf,
                                                         int d, int g) {
  const int c = blockIdx.x;
  const float dc = dcst[c];
  for (int k = threadIdx.x; k < g; k += blockDim.x) atomicAdd(gWj + (long long)c * (d + g) + d + k, dc * bf[k]);
  if (c == 0) {
    for (int k = threadIdx.x; k < g; k += blockDim.x) {
      float acc = 0.f;
      for (int cc = 0; cc < d; ++cc) acc = fmaf(dcst[cc], Wj[(long long)cc * (d + g) + d + k], acc);
      gbf[k] = acc;
    }
    for (int k = threadIdx.x; k < d; k += blockDim.x) gbj[k] = dcst[k];
  }
}

inline size_t fused_train_smem() { return sizeof(TrainSmem) + 16; }

}  // namespace carca
