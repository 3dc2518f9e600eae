// Fused inference forward of CARCA for d = 64, L <= 52 (the Beauty configuration):
//   ids/context -> embedding (folded item table gather) -> n_blocks x SelfAttentionBlock -> final
//   LayerNorm -> decoder (cross-attention or dot-product) -> probabilities y[B, T]
// in ONE kernel.  Reference path replaced: CARCA.forward in eval mode, src/carca.py:411-431, with
// AllEmbedding :85-95, SelfAttentionBlock :297-318, MultiHeadAttention :228-265,
// CrossAttentionBlock :338-349, DotProduct :358-365.
//
// Why: the modular kernels move every [B*L, d] intermediate through HBM (about 25 round trips per
// forward); here a CTA owns U = 2 users and keeps all of their activations in shared memory, so
// HBM traffic is the algorithmic minimum (ids + context in, scores out; the folded item table and
// the weights are L2-resident).
//
// Data layout in shared memory (floats).  Activations that feed a projection are stored TRANSPOSED,
// [feature k][row r] with row stride RP = 104, rows of user u at [u*LP, u*LP + L) (LP = 52 keeps
// every user 16-byte aligned), so the thread-tiled FFMA GEMM reads 8 consecutive rows with two
// 128-bit loads.  V is row-major [row][68] because it is the B operand of the P.V product.
// Five regions of 7072 floats are recycled through the block:
//   R0: x (block input)   -> S^T / P^T of the head being processed -> FFN output (next block's x)
//   R1: LN1(x)            -> attention + residual (in place)
//   R2: Q^T               -> LN2 output
//   R3: K^T               -> FFN hidden
//   R4: V (row-major)
// Weights are pre-transposed ([k][n], see carca_eval_prepare) and stream through a 2-slot ring,
// the next one prefetched into registers while the current GEMM runs.
#pragma once
#include "common.cuh"

namespace carca {

constexpr int FD = 64;          // model width handled by this kernel
constexpr int FLP = 52;         // padded rows per user
constexpr int FU = 2;           // users per CTA
constexpr int FRP = FU * FLP;   // 104 rows per tile
constexpr int FVS = 68;         // row stride of the row-major V region
constexpr int FREG = FRP * FVS; // 7072 floats per region
constexpr int FSM = 56;         // row stride of S^T per user (7 row groups of 8)
constexpr int FMAXB = 8;        // encoder blocks supported
constexpr int FTHREADS = 256;

struct FusedBlockW {
  const float *ln1_g, *ln1_b, *wqT, *bq, *wkT, *bk, *wvT, *bv, *ln2_g, *ln2_b, *w1T, *b1, *w2T, *b2;
};

struct FusedArgs {
  const float* Tfold;   // [n_items, 64] folded item table
  const float* Mc;      // [64, C] folded context projection
  const float* pos;     // optional [>=L, 64]
  const int* p_x;       // [B, L]
  const float* p_c;     // [B, L, C]
  const int* o_x;       // [B, T]
  const float* o_c;     // [B, T, C]
  float* y;             // [B, ldy]
  long long ldy;
  int col0;
  int B, L, T, C, H, n_blocks, residual_sa, residual_ca, decoder;  // decoder: 0 dot, 1 cross-attention
  FusedBlockW blk[FMAXB];
  const float *fn_g, *fn_b;
  const float *dwqT, *dbq, *dwkT, *dbk, *dwvT, *dbv, *dwf, *dbf;
  int cat_lo;           // > 0: full-catalog mode — candidate t of every user is item cat_lo + t
  long long oc_user, oc_tgt;   // strides (floats) of o_c over users / candidates ([B,T,C]: T*C, C; per-user rows: C, 0)
};

struct FusedSmem {
  float reg[5][FREG];
  float w[2][FD * FD];
  float mc[FD * 8];
  float pmask[FRP];
  float tmask[FRP];
  int pid[FRP];
  int tid_[FRP];
};

// acc[i][j] += sum_k A[k*lda + row0 + i] * Bm[k*ldb + col0 + j]
__device__ __forceinline__ void tile_fma(float (&acc)[8][4], const float* __restrict__ A, int lda,
                                         const float* __restrict__ Bm, int ldb, int K, int row0, int col0) {
  const float* ap = A + row0;
  const float* bp = Bm + col0;
#pragma unroll 4
  for (int k = 0; k < K; ++k) {
    const float4 a0 = *reinterpret_cast<const float4*>(ap + (long long)k * lda);
    const float4 a1 = *reinterpret_cast<const float4*>(ap + (long long)k * lda + 4);
    const float4 b = *reinterpret_cast<const float4*>(bp + (long long)k * ldb);
    const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
    const float bb[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], bb[j], acc[i][j]);
  }
}

__device__ __forceinline__ void zero_acc(float (&acc)[8][4]) {
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
}

struct WeightRing {
  float4 pre[4];
  int slot;
};

__device__ __forceinline__ void ring_prefetch(WeightRing& r, const float* __restrict__ wT) {
  const float4* src = reinterpret_cast<const float4*>(wT);
#pragma unroll
  for (int i = 0; i < 4; ++i) r.pre[i] = src[threadIdx.x + i * FTHREADS];
}
__device__ __forceinline__ void ring_commit(WeightRing& r, FusedSmem& s) {
  r.slot ^= 1;
  float4* dst = reinterpret_cast<float4*>(s.w[r.slot]);
#pragma unroll
  for (int i = 0; i < 4; ++i) dst[threadIdx.x + i * FTHREADS] = r.pre[i];
}

// out = A(rows x 64, transposed in smem) * W^T + bias for the 104-row tile; thread (ty, tx) owns
// rows 8ty..8ty+7, cols 4tx..4tx+3.  Epilogue variants write transposed or row-major.
enum { OUT_T = 0, OUT_ROW = 1 };

template <int MODE, bool LRELU, bool RESID>
__device__ __forceinline__ void project(float* __restrict__ out, const float* __restrict__ A_T,
                                        const float* __restrict__ W, const float* __restrict__ bias,
                                        const float* __restrict__ resid_T) {
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  if (ty >= FRP / 8) return;
  float acc[8][4];
  zero_acc(acc);
  tile_fma(acc, A_T, FRP, W, FD, FD, ty * 8, tx * 4);
  const float4 bv = *reinterpret_cast<const float4*>(bias + tx * 4);
  const float bb[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int c = tx * 4 + j;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float t = acc[i][j] + bb[j];
      if (LRELU) t = t > 0.f ? t : kLeakySlope * t;
      v[i] = t;
    }
    if (RESID) {
      const float4 r0 = *reinterpret_cast<const float4*>(resid_T + c * FRP + ty * 8);
      const float4 r1 = *reinterpret_cast<const float4*>(resid_T + c * FRP + ty * 8 + 4);
      v[0] += r0.x; v[1] += r0.y; v[2] += r0.z; v[3] += r0.w;
      v[4] += r1.x; v[5] += r1.y; v[6] += r1.z; v[7] += r1.w;
    }
    if (MODE == OUT_T) {
      *reinterpret_cast<float4*>(out + c * FRP + ty * 8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(out + c * FRP + ty * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) out[(ty * 8 + i) * FVS + c] = v[i];
    }
  }
}

// LayerNorm over the feature dim of a transposed tile: one thread per row.
__device__ __forceinline__ void layernorm_T(float* __restrict__ out_T, const float* __restrict__ in_T,
                                            const float* __restrict__ gamma, const float* __restrict__ beta) {
  const int r = threadIdx.x;
  if (r >= FRP) return;
  float s = 0.f;
#pragma unroll 8
  for (int k = 0; k < FD; ++k) s += in_T[k * FRP + r];
  const float mean = s / (float)FD;
  float v = 0.f;
#pragma unroll 8
  for (int k = 0; k < FD; ++k) {
    const float c = in_T[k * FRP + r] - mean;
    v = fmaf(c, c, v);
  }
  const float rstd = 1.0f / sqrtf(v / (float)FD + kLnEps);
#pragma unroll 8
  for (int k = 0; k < FD; ++k) out_T[k * FRP + r] = (in_T[k * FRP + r] - mean) * rstd * gamma[k] + beta[k];
}

// Gathers embeddings of `n` positions (ids in s_ids, masks in s_mask, ctx rows from global) into a
// transposed tile: e = mask * (Tfold[id] + Mc ctx (+ pos)).  Warp per row, lanes over features.
__device__ __forceinline__ void gather_embed(float* __restrict__ out_T, const FusedArgs& a, const FusedSmem& s,
                                             const int* __restrict__ ids, const float* __restrict__ msk,
                                             const float* __restrict__ ctx_base, int rows_per_user,
                                             long long ctx_user_stride, bool add_pos, int ctx_pos_stride) {
  // tile row r belongs to user r / rows_per_user at sequence position r % rows_per_user; ids/msk are
  // already laid out per tile row and ctx_base points at the first user's first position.
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  for (int r = w; r < FRP; r += FTHREADS / kWarp) {
    const int id = ids[r];
    const float m = msk[r];
    float v0 = 0.f, v1 = 0.f;
    if (m != 0.f) {
      const int u = r / rows_per_user, pos_in_user = r % rows_per_user;
      const float* trow = a.Tfold + (long long)id * FD;
      v0 = trow[lane];
      v1 = trow[lane + 32];
      const float* ctx = ctx_base + (long long)u * ctx_user_stride + (long long)pos_in_user * ctx_pos_stride;
      for (int c = 0; c < a.C; ++c) {
        const float cv = ctx[c];
        v0 = fmaf(s.mc[lane * 8 + c], cv, v0);
        v1 = fmaf(s.mc[(lane + 32) * 8 + c], cv, v1);
      }
      if (add_pos) {
        v0 += a.pos[(long long)pos_in_user * FD + lane];
        v1 += a.pos[(long long)pos_in_user * FD + lane + 32];
      }
      v0 *= m;
      v1 *= m;
    }
    out_T[lane * FRP + r] = v0;
    out_T[(lane + 32) * FRP + r] = v1;
  }
}

// Attention for one head over tiles laid out as described in the header.
//   self   : queries = keys = the U users' profile rows, causal (j <= i)
//   cross  : queries = `nq` target rows of ONE user (tile rows 0..nq), keys = that user's profile rows
// Q_T/K_T transposed [64][FRP], V row-major [FRP][FVS]; S region holds S^T (then P^T).
// Output: out_T[c][row] (+)= sum_j P[row][j] V[j][c] for the head's 32/64/16 columns.
template <bool CROSS>
__device__ __forceinline__ void attention_head(float* __restrict__ io_T, float* __restrict__ S, const float* __restrict__ Q_T,
                                               const float* __restrict__ K_T, const float* __restrict__ V,
                                               const float* __restrict__ qmask, const float* __restrict__ kmask,
                                               int h, int dh, int L, int nq, int key_row0, float sqrt_dh,
                                               bool accumulate_into_io) {
  const int tid = threadIdx.x;
  // ---- S^T[j][i] = sum_c Q[i][c] K[j][c]
  {
    const int col_groups = FLP / 4;                       // 13 key groups of 4
    if (!CROSS) {
      const int per_user = (FSM / 8) * col_groups;        // 7 * 13 = 91 thread tiles per user
      if (tid < FU * per_user) {
        const int u = tid / per_user, rem = tid % per_user;
        const int ty = rem / col_groups, tx = rem % col_groups;
        if (tx * 4 <= ty * 8 + 7) {                       // tile not entirely above the causal diagonal
          float acc[8][4];
          zero_acc(acc);
          tile_fma(acc, Q_T + h * dh * FRP + u * FLP, FRP, K_T + h * dh * FRP + u * FLP, FRP, dh, ty * 8, tx * 4);
          float* dst = S + u * (FLP * FSM);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float* p = dst + (tx * 4 + j) * FSM + ty * 8;
            *reinterpret_cast<float4*>(p) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
            *reinterpret_cast<float4*>(p + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
          }
        }
      }
    } else {
      const int row_groups = FRP / 8;                     // 13 query groups of 8
      if (tid < row_groups * col_groups) {
        const int ty = tid / col_groups, tx = tid % col_groups;
        float acc[8][4];
        zero_acc(acc);
        tile_fma(acc, Q_T + h * dh * FRP, FRP, K_T + h * dh * FRP + key_row0, FRP, dh, ty * 8, tx * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* p = S + (tx * 4 + j) * FRP + ty * 8;
          *reinterpret_cast<float4*>(p) = make_float4(acc[0][j], acc[1][j], acc[2][j], acc[3][j]);
          *reinterpret_cast<float4*>(p + 4) = make_float4(acc[4][j], acc[5][j], acc[6][j], acc[7][j]);
        }
      }
    }
  }
  __syncthreads();
  // ---- masked softmax, one thread per query row, in place S^T -> P^T (zeros where not allowed)
  {
    const int n_rows = CROSS ? FRP : FU * FSM;
    if (tid < n_rows) {
      int u = 0, i = tid;
      float* col;
      int ld;
      bool live;
      float qm;
      const float* km;
      if (!CROSS) {
        u = tid / FSM;
        i = tid % FSM;
        col = S + u * (FLP * FSM) + i;
        ld = FSM;
        live = i < L;
        qm = live ? qmask[u * FLP + i] : 0.f;
        km = kmask + u * FLP;
      } else {
        col = S + i;
        ld = FRP;
        live = i < nq;
        qm = live ? qmask[i] : 0.f;
        km = kmask + key_row0;
      }
      const int jmax = CROSS ? L : min(L, i + 1);        // causal: keys j <= i
      float mx = -INFINITY;
      if (qm != 0.f)
        for (int j = 0; j < jmax; ++j)
          if (km[j] != 0.f) mx = fmaxf(mx, col[j * ld] / sqrt_dh);
      float sum = 0.f;
      if (mx != -INFINITY)
        for (int j = 0; j < jmax; ++j)
          if (km[j] != 0.f) sum += expf(col[j * ld] / sqrt_dh - mx);
      const float inv = sum > 0.f ? 1.0f / sum : 0.f;
      for (int j = 0; j < FLP; ++j) {
        float p = 0.f;
        if (inv != 0.f && j < jmax && km[j] != 0.f) p = expf(col[j * ld] / sqrt_dh - mx) * inv;
        col[j * ld] = p;
      }
    }
  }
  __syncthreads();
  // ---- O[i][c] = sum_j P[i][j] V[j][c]  -> io_T[h*dh + c][row] (+= when it already holds the residual)
  {
    const int cg = dh / 4;
    if (!CROSS) {
      const int per_user = (FSM / 8) * cg;
      if (tid < FU * per_user) {
        const int u = tid / per_user, rem = tid % per_user;
        const int ty = rem / cg, tx = rem % cg;
        float acc[8][4];
        zero_acc(acc);
        tile_fma(acc, S + u * (FLP * FSM), FSM, V + (u * FLP) * FVS + h * dh, FVS, L, ty * 8, tx * 4);
        if (ty * 8 < FLP) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float* p = io_T + (h * dh + tx * 4 + j) * FRP + u * FLP + ty * 8;
            // rows 48..55 of the last group: only 48..51 belong to this user's slot
            const int lim = min(8, FLP - ty * 8);
            for (int i = 0; i < lim; ++i) p[i] = accumulate_into_io ? p[i] + acc[i][j] : acc[i][j];
          }
        }
      }
    } else {
      const int row_groups = FRP / 8;
      if (tid < row_groups * cg) {
        const int ty = tid / cg, tx = tid % cg;
        float acc[8][4];
        zero_acc(acc);
        tile_fma(acc, S, FRP, V + key_row0 * FVS + h * dh, FVS, L, ty * 8, tx * 4);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float* p = io_T + (h * dh + tx * 4 + j) * FRP + ty * 8;
#pragma unroll
          for (int i = 0; i < 8; ++i) p[i] = accumulate_into_io ? p[i] + acc[i][j] : acc[i][j];
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(FTHREADS, 1) fused_eval_kernel(const FusedArgs a) {
  CARCA_DYN_SMEM(unsigned char, raw);
  FusedSmem& s = *reinterpret_cast<FusedSmem*>(raw);
  const int tid = threadIdx.x;
  const int L = a.L, dh = FD / a.H;
  const float sqrt_dh = sqrtf((float)dh);
  float* R0 = s.reg[0];
  float* R1 = s.reg[1];
  float* R2 = s.reg[2];
  float* R3 = s.reg[3];
  float* R4 = s.reg[4];
  const int n_tiles = (a.B + FU - 1) / FU;

  for (int i = tid; i < FD * 8; i += FTHREADS) s.mc[i] = a.Mc[i];   // [64][8], zero-padded beyond C

  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int user0 = tile * FU;
    __syncthreads();
    // ---- profile ids / masks for the tile rows
    for (int r = tid; r < FRP; r += FTHREADS) {
      const int u = r / FLP, i = r % FLP;
      int id = 0;
      if (user0 + u < a.B && i < L) id = a.p_x[(long long)(user0 + u) * L + i];
      s.pid[r] = id;
      s.pmask[r] = id != 0 ? 1.f : 0.f;
    }
    WeightRing ring;
    ring.slot = 1;
    ring_prefetch(ring, a.blk[0].wqT);
    __syncthreads();
    // ---- embedding of the profile (src/carca.py:415; dropout :416 is identity in eval)
    gather_embed(R0, a, s, s.pid, s.pmask, a.p_c + (long long)user0 * L * a.C, FLP, (long long)L * a.C,
                 a.pos != nullptr, a.C);
    ring_commit(ring, s);   // slot 0 <- block 0 WQ^T
    __syncthreads();

    for (int b = 0; b < a.n_blocks; ++b) {
      const FusedBlockW& w = a.blk[b];
      // LN1 (:298)
      layernorm_T(R1, R0, w.ln1_g, w.ln1_b);
      __syncthreads();
      // Q from LN1(x), K and V from the raw x (:299, :238-240)
      ring_prefetch(ring, w.wkT);
      project<OUT_T, false, false>(R2, R1, s.w[ring.slot], w.bq, nullptr);
      ring_commit(ring, s);
      __syncthreads();
      ring_prefetch(ring, w.wvT);
      project<OUT_T, false, false>(R3, R0, s.w[ring.slot], w.bk, nullptr);
      ring_commit(ring, s);
      __syncthreads();
      ring_prefetch(ring, w.w1T);
      project<OUT_ROW, false, false>(R4, R0, s.w[ring.slot], w.bv, nullptr);
      ring_commit(ring, s);
      __syncthreads();
      // attention, head by head; R1 holds LN1(x) = the residual (:302), or is cleared first
      if (!a.residual_sa) {
        for (int i = tid; i < FD * FRP; i += FTHREADS) R1[i] = 0.f;
        __syncthreads();
      }
      for (int h = 0; h < a.H; ++h)
        attention_head<false>(R1, R0, R2, R3, R4, s.pmask, s.pmask, h, dh, L, 0, 0, sqrt_dh, true);
      // LN2 (:304) -> R2
      layernorm_T(R2, R1, w.ln2_g, w.ln2_b);
      __syncthreads();
      // FFN (:305-316): hidden -> R3, output (+ LN2 residual) -> R0
      ring_prefetch(ring, w.w2T);
      project<OUT_T, true, false>(R3, R2, s.w[ring.slot], w.b1, nullptr);
      ring_commit(ring, s);
      __syncthreads();
      const float* next_w = (b + 1 < a.n_blocks) ? a.blk[b + 1].wqT : (a.decoder == 1 ? a.dwkT : a.blk[0].wqT);
      ring_prefetch(ring, next_w);
      if (a.residual_sa) project<OUT_T, false, true>(R0, R3, s.w[ring.slot], w.b2, R2);
      else project<OUT_T, false, false>(R0, R3, s.w[ring.slot], w.b2, nullptr);
      ring_commit(ring, s);
      __syncthreads();
    }
    // ---- final LayerNorm (:421) -> R1 (transposed)
    layernorm_T(R1, R0, a.fn_g, a.fn_b);
    __syncthreads();

    if (a.decoder == 1) {
      // keys / values of the encoded profile for both users (:239-240 with p as key and value)
      ring_prefetch(ring, a.dwvT);
      project<OUT_T, false, false>(R3, R1, s.w[ring.slot], a.dbk, nullptr);
      ring_commit(ring, s);
      __syncthreads();
      ring_prefetch(ring, a.dwqT);
      project<OUT_ROW, false, false>(R4, R1, s.w[ring.slot], a.dbv, nullptr);
      ring_commit(ring, s);     // ring slot now holds WQ^T of the decoder and stays there
      __syncthreads();
    }

    for (int u = 0; u < FU; ++u) {
      if (user0 + u >= a.B) break;
      for (int t0 = 0; t0 < a.T; t0 += FRP) {
        const int nq = min(FRP, a.T - t0);
        __syncthreads();
        for (int r = tid; r < FRP; r += FTHREADS) {
          int id = 0;
          if (r < nq) id = a.cat_lo > 0 ? a.cat_lo + t0 + r : a.o_x[(long long)(user0 + u) * a.T + t0 + r];
          s.tid_[r] = id;
          s.tmask[r] = id != 0 ? 1.f : 0.f;
        }
        __syncthreads();
        // target embeddings (:426, no positional encoding) -> R0 transposed
        gather_embed(R0, a, s, s.tid_, s.tmask, a.o_c + (long long)(user0 + u) * a.oc_user + (long long)t0 * a.oc_tgt,
                     FRP, 0, false, (int)a.oc_tgt);
        __syncthreads();
        if (a.decoder == 1) {
          project<OUT_T, false, false>(R2, R0, s.w[ring.slot], a.dbq, nullptr);   // Q of the targets
          __syncthreads();
          // s = attention (+ o) accumulated in place over the target embeddings (:340-343)
          if (!a.residual_ca) {   // no residual: the output is the attention alone
            for (int i = tid; i < FD * FRP; i += FTHREADS) R0[i] = 0.f;
            __syncthreads();
          }
          // S^T goes to R1: the final-LN output there was consumed by the K/V projections above
          for (int h = 0; h < a.H; ++h)
            attention_head<true>(R0, R1, R2, R3, R4, s.tmask, s.pmask, h, dh, L, nq, u * FLP, sqrt_dh, true);
          // y = sigmoid(<s, wf> + bf) (:345-347), one thread per target row
          if (tid < nq) {
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < FD; ++k) acc = fmaf(R0[k * FRP + tid], a.dwf[k], acc);
            a.y[(long long)(user0 + u) * a.ldy + a.col0 + t0 + tid] = 1.0f / (1.0f + expf(-(acc + a.dbf[0])));
          }
        } else {
          // dot-product with the last profile position (:362); R1 = final LN output, transposed
          if (tid < nq) {
            const int last = u * FLP + L - 1;
            float acc = 0.f;
#pragma unroll 8
            for (int k = 0; k < FD; ++k) acc = fmaf(R0[k * FRP + tid], R1[k * FRP + last], acc);
            a.y[(long long)(user0 + u) * a.ldy + a.col0 + t0 + tid] = 1.0f / (1.0f + expf(-acc));
          }
        }
      }
    }
  }
}

}  // namespace carca
