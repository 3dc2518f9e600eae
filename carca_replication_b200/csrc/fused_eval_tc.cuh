// Tensor-core (tcgen05) version of the fused inference forward for d = 64, L <= 64, heads of
// width 16 or 32.  Same contract as fused_eval_kernel (fused_eval.cuh); reference path replaced:
// CARCA.forward in eval mode, src/carca.py:411-431 with :85-95, :228-265, :297-318, :338-365.
//
// Design (per CTA: 2 users = 128 rows, one TMEM lane per row, 256 threads = 2 threads per row,
// each owning 32 of the 64 feature columns):
//   * ACTIVATIONS LIVE IN TENSOR MEMORY.  Every [128 x 64] activation is a block of 64 TMEM
//     columns; LayerNorm, softmax, bias/activation epilogues are thread-local row operations on
//     registers loaded with tcgen05.ld and written back with tcgen05.st.  Activations never touch
//     shared memory or HBM.
//   * every projection / QK^T / PV product is a tcgen05.mma with M = 128, the A operand read
//     straight from TMEM, the B operand (weights, K, V) from shared memory in the K-major
//     no-swizzle layout of umma.cuh, accumulating into other TMEM columns.
//   * fp32-grade accuracy on the tf32 tensor cores via the 3xTF32 split (hi*hi + lo*hi + hi*lo);
//     biases ride along as one extra K step against a constant [1,0,..] column block.
//   * shared memory holds only B operands: a 2-slot weight ring (cp.async prefetch), K (64 KB),
//     V (2 users x 16.6 KB x hi/lo, K-major over keys with a padded chunk stride so the
//     thread-per-key scalar stores are bank-conflict free).
// TMEM column map: X_HI 0, X_LO 64, QN_HI 128, QN_LO 192, ACC_Q 256, ACC_K 320, ACC_V 384, ONES 448.
#pragma once
#include "common.cuh"
#include "fused_eval.cuh"
#include "umma.cuh"

#ifndef CARCA_EMU
namespace carca {

constexpr int TC_THREADS = 256;
constexpr int TC_WFLOATS = 18 * 64 * 4;              // packed weight: 16 k-chunks + 2 bias chunks, hi or lo
constexpr int TC_VLBO = 64 * 16 + 16;                // bytes between key chunks of the V operand (padded)
constexpr int TC_VUSER = 16 * (TC_VLBO / 4);         // floats per user per hi/lo
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
  const float *dwq, *dwk, *dwv;                      // packed decoder projections
  const float *dwf, *dbf;
  int* status;                                       // [0] set to 1 if an MMA wait timed out
  float* dbg;                                        // optional [128, 64] dump of the stage `dbg_stage`
  int dbg_stage;
};

struct TcSmem {
  float w[2][2 * TC_WFLOATS];                        // weight ring: [slot][hi | lo]
  float k_hi[16 * 128 * 4];
  float k_lo[16 * 128 * 4];
  float v_hi[2 * TC_VUSER];
  float v_lo[2 * TC_VUSER];
  float mc[64 * 8];
  float xch[4][2][128];
  float plast[2][64];
  float pmask[128];
  float tmask[128];
  int pid[128];
  int tid_[128];
  uint64_t bar;
  uint32_t tmem_slot;
};

struct TcCtx {
  TcSmem* s;
  uint32_t tmem;       // TMEM base
  uint32_t lane_base;  // this warp's lane quarter << 16
  int row, half, tid;
  uint32_t phase;      // mbarrier phase parity
  int xslot;
  int* status;
};

__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
    umma::tmem_ld8(taddr + 8 * i, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[8 * i + j] = t[j];
  }
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    float t[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) t[j] = v[8 * i + j];
    umma::tmem_st8(taddr + 8 * i, t);
  }
}
// store a half row as a tf32 operand pair: hi = the fp32 value, lo = its tf32 remainder
__device__ __forceinline__ void st_operand(const TcCtx& c, int col_hi, int col_lo, const float (&v)[32]) {
  float lo[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) lo[j] = umma::tf32_lo(v[j]);
  tmem_st32(c.tmem + c.lane_base + col_hi + 32 * c.half, v);
  tmem_st32(c.tmem + c.lane_base + col_lo + 32 * c.half, lo);
}
__device__ __forceinline__ void ld_half(const TcCtx& c, int col, float (&v)[32]) {
  tmem_ld32(c.tmem + c.lane_base + col + 32 * c.half, v);
}

// all TMEM / shared-memory writes of every thread become visible to the MMA issued afterwards
__device__ __forceinline__ void publish(TcCtx& c) {
  umma::tmem_st_wait();
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
}
// the issuing thread commits; everyone waits for the MMAs to land in TMEM
__device__ __forceinline__ void commit_and_wait(TcCtx& c) {
  if (c.tid == 0) umma::commit(&c.s->bar);
  if (!umma::mbar_wait(&c.s->bar, c.phase & 1u)) c.status[0] = 1;
  c.phase++;
  umma::fence_after_sync();
}
// combine a per-row partial of the two half-row threads (sum or max)
template <bool MAX>
__device__ __forceinline__ float row_combine(TcCtx& c, float v) {
  float* x = &c.s->xch[c.xslot][0][0];
  x[c.half * 128 + c.row] = v;
  __syncthreads();
  const float o = x[(c.half ^ 1) * 128 + c.row];
  c.xslot = (c.xslot + 1) & 3;
  return MAX ? fmaxf(v, o) : v + o;
}

// ---- MMA issue helpers (one thread) ---------------------------------------------------------
// D[d_col .. d_col+N) (=) A(tmem, K cols at a_hi/a_lo) x B(smem hi/lo)^T, 3xTF32
__device__ __forceinline__ void issue_3x(uint32_t tmem, int d_col, int a_hi, int a_lo, uint32_t b_hi, uint32_t b_lo,
                                         int N, int K, uint32_t b_lbo, bool first) {
  const uint32_t idesc = umma::idesc_tf32(N);
  for (int p = 0; p < 3; ++p) {
    const int a = (p == 1) ? a_lo : a_hi;
    const uint32_t b = (p == 2) ? b_lo : b_hi;
    for (int ks = 0; ks < K / 8; ++ks)
      umma::mma_tf32_ts(tmem + d_col, tmem + a + 8 * ks, umma::smem_desc(b + ks * 2 * b_lbo, b_lbo, 128), idesc,
                        !(first && p == 0 && ks == 0));
  }
}
// projection with a packed weight (bias folded in as K step 8 against the ONES block)
__device__ __forceinline__ void issue_proj(uint32_t tmem, int d_col, int a_hi, int a_lo, const float* w_slot) {
  const uint32_t b_hi = umma::smem_u32(w_slot), b_lo = umma::smem_u32(w_slot + TC_WFLOATS);
  const uint32_t lbo = 64 * 16;
  issue_3x(tmem, d_col, a_hi, a_lo, b_hi, b_lo, 64, 64, lbo, true);
  const uint32_t idesc = umma::idesc_tf32(64);
  umma::mma_tf32_ts(tmem + d_col, tmem + C_ONES, umma::smem_desc(b_hi + 16 * lbo, lbo, 128), idesc, true);
  umma::mma_tf32_ts(tmem + d_col, tmem + C_ONES, umma::smem_desc(b_lo + 16 * lbo, lbo, 128), idesc, true);
}

// ---- weight ring: cp.async global -> shared, 36,864 bytes per packed weight ---------------------
__device__ __forceinline__ void weight_prefetch(TcCtx& c, int slot, const float* __restrict__ w) {
  const uint32_t dst = umma::smem_u32(c.s->w[slot]);
  for (int i = c.tid; i < 2 * TC_WFLOATS / 4; i += TC_THREADS)
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + i * 16), "l"(w + i * 4) : "memory");
  asm volatile("cp.async.commit_group;" ::: "memory");
}
__device__ __forceinline__ void weight_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

// LayerNorm of the half rows held in v (two-pass, as torch), gamma/beta from global
__device__ __forceinline__ void layernorm_rows(TcCtx& c, float (&v)[32], const float* __restrict__ g,
                                               const float* __restrict__ b) {
  float s = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) s += v[j];
  const float mean = row_combine<false>(c, s) / 64.0f;
  float q = 0.f;
#pragma unroll
  for (int j = 0; j < 32; ++j) {
    const float d = v[j] - mean;
    q = fmaf(d, d, q);
  }
  const float rstd = 1.0f / sqrtf(row_combine<false>(c, q) / 64.0f + kLnEps);
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = (v[j] - mean) * rstd * __ldg(g + 32 * c.half + j) + __ldg(b + 32 * c.half + j);
}

// e = mask * (Tfold[id] + Mc ctx (+ pos)) for this thread's half row
__device__ __forceinline__ void embed_row(const TcArgs& a, const TcCtx& c, int id, float m, const float* ctx,
                                          const float* pos_row, float (&v)[32]) {
  if (m == 0.f) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = 0.f;
    return;
  }
  const float4* t = reinterpret_cast<const float4*>(a.Tfold + (long long)id * 64 + 32 * c.half);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float4 x = __ldg(t + i);
    v[4 * i] = x.x; v[4 * i + 1] = x.y; v[4 * i + 2] = x.z; v[4 * i + 3] = x.w;
  }
  for (int k = 0; k < a.C; ++k) {
    const float cv = ctx[k];
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = fmaf(c.s->mc[(32 * c.half + j) * 8 + k], cv, v[j]);
  }
  if (pos_row) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += pos_row[32 * c.half + j];
  }
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] *= m;
}

__device__ __forceinline__ void dump_stage(const TcArgs& a, TcCtx& c, int stage, int col) {
  if (a.dbg && a.dbg_stage == stage && blockIdx.x == 0) {
    float v[32];
    ld_half(c, col, v);
#pragma unroll
    for (int j = 0; j < 32; ++j) a.dbg[c.row * 64 + 32 * c.half + j] = v[j];
  }
}

// K operand: this thread's half row of the accumulator at `col` -> shared K-major chunks (hi/lo)
__device__ __forceinline__ void store_k_operand(TcCtx& c, int col) {
  float v[32];
  ld_half(c, col, v);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int kc = 8 * c.half + i;
    const float4 hi = make_float4(v[4 * i], v[4 * i + 1], v[4 * i + 2], v[4 * i + 3]);
    const float4 lo = make_float4(umma::tf32_lo(hi.x), umma::tf32_lo(hi.y), umma::tf32_lo(hi.z), umma::tf32_lo(hi.w));
    reinterpret_cast<float4*>(c.s->k_hi)[kc * 128 + c.row] = hi;
    reinterpret_cast<float4*>(c.s->k_lo)[kc * 128 + c.row] = lo;
  }
}
// V operand: K-major over keys, [user][key/4][feature][key%4] with the padded chunk stride
__device__ __forceinline__ void store_v_operand(TcCtx& c, int col) {
  float v[32];
  ld_half(c, col, v);
  const int u = c.row / 64, j = c.row % 64;
  float* hi = c.s->v_hi + u * TC_VUSER + (j / 4) * (TC_VLBO / 4) + (j % 4);
  float* lo = c.s->v_lo + u * TC_VUSER + (j / 4) * (TC_VLBO / 4) + (j % 4);
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const int f = 32 * c.half + t;
    hi[f * 4] = v[t];
    lo[f * 4] = umma::tf32_lo(v[t]);
  }
}

// One attention head.  Queries: the 128 tile rows (ACC_Q hi / QN_LO lo, head columns h*dh..).
//   SELF : keys of both users (N = 128), row m uses its own user's 64 score columns, causal.
//   CROSS: keys of user `ku` only (N = 64), no causal mask, query mask = tmask.
// P overwrites the score columns region-locally; O_h lands at o_col (+ user * dh for SELF).
template <bool CROSS>
__device__ __forceinline__ void attention_head_tc(const TcArgs& a, TcCtx& c, int h, int dh, int ku, int s_col,
                                                  int p_hi_col, int p_lo_col, int o_col, float sqrt_dh) {
  TcSmem& s = *c.s;
  const int L = a.L;
  // ---- scores
  if (c.tid == 0) {
    const uint32_t koff = (uint32_t)(h * dh / 4) * 2048u + (CROSS ? (uint32_t)ku * 64u * 16u : 0u);
    issue_3x(c.tmem, s_col, C_ACCQ + h * dh, C_QNLO + h * dh, umma::smem_u32(s.k_hi) + koff,
             umma::smem_u32(s.k_lo) + koff, CROSS ? 64 : 128, dh, 2048u, true);
  }
  commit_and_wait(c);
  // ---- masked softmax over this row's 64 keys, 32 per thread
  const int u = CROSS ? ku : c.row / 64;
  const int i = c.row % 64;
  float v[32];
  tmem_ld32(c.tmem + c.lane_base + s_col + (CROSS ? 0 : 64 * (c.row / 64)) + 32 * c.half, v);
  const float qm = CROSS ? s.tmask[c.row] : s.pmask[c.row];
  float mx = -INFINITY;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    const int j = 32 * c.half + t;
    const bool ok = qm != 0.f && j < L && s.pmask[u * 64 + j] != 0.f && (CROSS || j <= i);
    v[t] = ok ? v[t] / sqrt_dh : -INFINITY;
    mx = fmaxf(mx, v[t]);
  }
  mx = row_combine<true>(c, mx);
  float sum = 0.f;
#pragma unroll
  for (int t = 0; t < 32; ++t) {
    v[t] = (v[t] == -INFINITY) ? 0.f : expf(v[t] - mx);
    sum += v[t];
  }
  sum = row_combine<false>(c, sum);
  const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
  for (int t = 0; t < 32; ++t) v[t] *= inv;
  st_operand(c, p_hi_col, p_lo_col, v);
  publish(c);
  // ---- O_h = P V_h  (SELF: once per user's V; each row later picks its own user's result)
  if (c.tid == 0) {
    for (int vu = CROSS ? ku : 0; vu <= (CROSS ? ku : 1); ++vu) {
      const uint32_t voff = (uint32_t)(vu * TC_VUSER) * 4u + (uint32_t)(h * dh) * 16u;
      issue_3x(c.tmem, o_col + (CROSS ? 0 : vu * dh), p_hi_col, p_lo_col, umma::smem_u32(s.v_hi) + voff,
               umma::smem_u32(s.v_lo) + voff, dh, 64, (uint32_t)TC_VLBO, true);
    }
  }
  commit_and_wait(c);
}

__global__ void __launch_bounds__(TC_THREADS, 1) fused_eval_tc_kernel(const TcArgs a) {
  CARCA_DYN_SMEM(unsigned char, raw);
  TcSmem& s = *reinterpret_cast<TcSmem*>(raw);
  TcCtx c;
  c.s = &s;
  c.tid = threadIdx.x;
  const int w = c.tid / 32;
  c.row = 32 * (w % 4) + (c.tid % 32);
  c.half = w / 4;
  c.lane_base = (uint32_t)(32 * (w % 4)) << 16;
  c.phase = 0;
  c.xslot = 0;
  c.status = a.status;
  const int L = a.L, dh = 64 / a.H;
  const float sqrt_dh = sqrtf((float)dh);

  if (w == 0) umma::tmem_alloc(&s.tmem_slot, 512);
  if (c.tid == 0) umma::mbar_init(&s.bar, 1);
  for (int i = c.tid; i < 64 * 8; i += TC_THREADS) s.mc[i] = a.Mc[i];
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  c.tmem = s.tmem_slot;
  if (c.half == 0) {   // constant [1,0,0,0,0,0,0,0] column block: the A operand of every bias step
    float ones[8] = {1.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    umma::tmem_st8(c.tmem + c.lane_base + C_ONES, ones);
  }

  const int n_tiles = (a.B + 1) / 2;
  for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int user0 = tile * 2;
    const int u = c.row / 64, i = c.row % 64;
    __syncthreads();
    if (c.half == 0) {
      int id = 0;
      if (user0 + u < a.B && i < L) id = a.p_x[(long long)(user0 + u) * L + i];
      s.pid[c.row] = id;
      s.pmask[c.row] = id != 0 ? 1.f : 0.f;
    }
    weight_prefetch(c, 0, a.blk[0].wq);
    weight_prefetch(c, 1, a.blk[0].wk);
    __syncthreads();
    {   // profile embedding (src/carca.py:415) -> X
      float v[32];
      const float* ctx = a.p_c + ((long long)(user0 + u) * L + i) * a.C;
      embed_row(a, c, s.pid[c.row], s.pmask[c.row], ctx, a.pos ? a.pos + (long long)i * 64 : nullptr, v);
      st_operand(c, C_XHI, C_XLO, v);
      umma::tmem_st_wait();   // LN1 below reads these columns back
    }

    for (int b = 0; b < a.n_blocks; ++b) {
      const TcBlockW& wb = a.blk[b];
      {   // LN1 (:298) -> QN
        float v[32];
        ld_half(c, C_XHI, v);
        layernorm_rows(c, v, wb.ln1_g, wb.ln1_b);
        st_operand(c, C_QNHI, C_QNLO, v);
      }
      weight_wait_all();
      publish(c);
      dump_stage(a, c, 1 + 10 * b, C_QNHI);
      // Q from LN1(x), K from raw x (:238-239); V's weights are loaded while these run
      if (c.tid == 0) {
        issue_proj(c.tmem, C_ACCQ, C_QNHI, C_QNLO, s.w[0]);
        issue_proj(c.tmem, C_ACCK, C_XHI, C_XLO, s.w[1]);
      }
      commit_and_wait(c);
      weight_prefetch(c, 0, wb.wv);
      weight_wait_all();
      publish(c);
      if (c.tid == 0) issue_proj(c.tmem, C_ACCV, C_XHI, C_XLO, s.w[0]);
      // overlap with the V projection: K operand to shared memory, Q remainder to QN_LO
      weight_prefetch(c, 1, wb.w1);
      store_k_operand(c, C_ACCK);
      {
        float v[32], lo[32];
        ld_half(c, C_ACCQ, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) lo[j] = umma::tf32_lo(v[j]);
        tmem_st32(c.tmem + c.lane_base + C_QNLO + 32 * c.half, lo);
      }
      commit_and_wait(c);
      dump_stage(a, c, 2 + 10 * b, C_ACCQ);
      dump_stage(a, c, 3 + 10 * b, C_ACCK);
      dump_stage(a, c, 4 + 10 * b, C_ACCV);
      store_v_operand(c, C_ACCV);
      weight_prefetch(c, 0, wb.w2);
      if (!a.residual_sa) {   // no residual: the attention output replaces LN1(x)
        float z[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) z[j] = 0.f;
        tmem_st32(c.tmem + c.lane_base + C_QNHI + 32 * c.half, z);
      }
      publish(c);
      for (int h = 0; h < a.H; ++h) {
        attention_head_tc<false>(a, c, h, dh, 0, C_XHI, C_XHI, C_XLO, C_ACCV, sqrt_dh);
        // s[:, head cols] = LN1(x) + O_h   (in place in QN_HI; :302)
        const int n = dh / 2;   // columns per thread
        const int src = C_ACCV + u * dh + c.half * n, dst = C_QNHI + h * dh + c.half * n;
        for (int c0 = 0; c0 < n; c0 += 8) {
          float o[8], q[8];
          umma::tmem_ld8(c.tmem + c.lane_base + src + c0, o);
          umma::tmem_ld8(c.tmem + c.lane_base + dst + c0, q);
#pragma unroll
          for (int j = 0; j < 8; ++j) q[j] += o[j];
          umma::tmem_st8(c.tmem + c.lane_base + dst + c0, q);
        }
        publish(c);
      }
      dump_stage(a, c, 5 + 10 * b, C_QNHI);
      float s2[32];
      {   // LN2 (:304) -> X (operand of ffn_1 and the FFN residual)
        ld_half(c, C_QNHI, s2);
        layernorm_rows(c, s2, wb.ln2_g, wb.ln2_b);
        st_operand(c, C_XHI, C_XLO, s2);
      }
      weight_wait_all();
      publish(c);
      if (c.tid == 0) issue_proj(c.tmem, C_ACCQ, C_XHI, C_XLO, s.w[1]);    // ffn_1 (:307)
      commit_and_wait(c);
      {
        float v[32];
        ld_half(c, C_ACCQ, v);
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = v[j] > 0.f ? v[j] : kLeakySlope * v[j];   // LeakyReLU (:308)
        st_operand(c, C_QNHI, C_QNLO, v);
      }
      publish(c);
      if (c.tid == 0) issue_proj(c.tmem, C_ACCK, C_QNHI, C_QNLO, s.w[0]);  // ffn_2 (:311)
      commit_and_wait(c);
      // next weights: the following block's WQ/WK, or the decoder's WK/WV
      if (b + 1 < a.n_blocks) {
        weight_prefetch(c, 0, a.blk[b + 1].wq);
        weight_prefetch(c, 1, a.blk[b + 1].wk);
      } else if (a.decoder == 1) {
        weight_prefetch(c, 0, a.dwk);
        weight_prefetch(c, 1, a.dwv);
      }
      {   // block output (+ LN2 residual, :316) -> X for the next block
        float v[32];
        ld_half(c, C_ACCK, v);
        if (a.residual_sa) {
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] += s2[j];
        }
        st_operand(c, C_XHI, C_XLO, v);
      }
      publish(c);
      dump_stage(a, c, 9 + 10 * b, C_XHI);
    }

    {   // final LayerNorm (:421) -> QN (operand of the decoder's K/V projections)
      float v[32];
      ld_half(c, C_XHI, v);
      layernorm_rows(c, v, a.fn_g, a.fn_b);
      st_operand(c, C_QNHI, C_QNLO, v);
      if (i == L - 1) {
#pragma unroll
        for (int j = 0; j < 32; ++j) s.plast[u][32 * c.half + j] = v[j];
      }
    }
    weight_wait_all();
    publish(c);
    dump_stage(a, c, 100, C_QNHI);

    if (a.decoder == 1) {   // keys / values of the encoded profile (:239-240 with p as key and value)
      if (c.tid == 0) {
        issue_proj(c.tmem, C_ACCK, C_QNHI, C_QNLO, s.w[0]);
        issue_proj(c.tmem, C_ACCV, C_QNHI, C_QNLO, s.w[1]);
      }
      commit_and_wait(c);
      weight_prefetch(c, 0, a.dwq);
      store_k_operand(c, C_ACCK);
      store_v_operand(c, C_ACCV);
      weight_wait_all();
      publish(c);
    }

    for (int du = 0; du < 2; ++du) {
      if (user0 + du >= a.B) break;
      for (int t0 = 0; t0 < a.T; t0 += 128) {
        const int nq = min(128, a.T - t0);
        __syncthreads();
        if (c.half == 0) {
          int id = 0;
          if (c.row < nq) id = a.o_x[(long long)(user0 + du) * a.T + t0 + c.row];
          s.tid_[c.row] = id;
          s.tmask[c.row] = id != 0 ? 1.f : 0.f;
        }
        __syncthreads();
        float e[32];   // target embedding half row (:426), also the residual of the decoder
        {
          const float* ctx = a.o_c + ((long long)(user0 + du) * a.T + t0 + min(c.row, nq - 1)) * a.C;
          embed_row(a, c, s.tid_[c.row], s.tmask[c.row], ctx, nullptr, e);
        }
        float acc = 0.f;
        if (a.decoder == 1) {
          st_operand(c, C_XHI, C_XLO, e);
          publish(c);
          if (c.tid == 0) issue_proj(c.tmem, C_ACCQ, C_XHI, C_XLO, s.w[0]);   // Q of the targets
          commit_and_wait(c);
          {
            float v[32], lo[32];
            ld_half(c, C_ACCQ, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) lo[j] = umma::tf32_lo(v[j]);
            tmem_st32(c.tmem + c.lane_base + C_QNLO + 32 * c.half, lo);
          }
          publish(c);
          for (int h = 0; h < a.H; ++h)
            attention_head_tc<true>(a, c, h, dh, du, C_ACCK, C_QNHI, C_XLO, C_ACCV + h * dh, sqrt_dh);
          float o[32];   // s = attention (+ o) (:340-343); y = sigmoid(<s, wf> + bf) (:345-347)
          ld_half(c, C_ACCV, o);
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float sv = o[j] + (a.residual_ca ? e[j] : 0.f);
            acc = fmaf(sv, __ldg(a.dwf + 32 * c.half + j), acc);
          }
          acc = row_combine<false>(c, acc) + __ldg(a.dbf);
        } else {   // dot product with the last profile position (:362)
#pragma unroll
          for (int j = 0; j < 32; ++j) acc = fmaf(e[j], s.plast[du][32 * c.half + j], acc);
          acc = row_combine<false>(c, acc);
        }
        if (c.half == 0 && c.row < nq)
          a.y[(long long)(user0 + du) * a.ldy + a.col0 + t0 + c.row] = 1.0f / (1.0f + expf(-acc));
      }
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (w == 0) umma::tmem_free(c.tmem, 512);
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
