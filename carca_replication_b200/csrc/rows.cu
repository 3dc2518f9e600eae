// extern "C" entry points of the bf16 packed-rows inference pipeline (rows_bf16.cuh); declared in
// include/carca_b200.h.  Reference path replaced: CARCA.forward in eval mode, src/carca.py:411-431.
#include "../../include/carca_b200.h"

#include <cmath>

#include "common.cuh"
#include "gemm.cuh"
#include "plan_layout.cuh"
#include "rows_bf16.cuh"

using namespace carca;

#define TRY(expr)             \
  do {                        \
    int _rc = (expr);         \
    if (_rc != 0) return _rc; \
  } while (0)

#ifdef CARCA_EMU
// the tcgen05 pipeline has no CPU emulation: the development emulator (tools/emu) only covers the plain-CUDA kernels
extern "C" {
int64_t carca_rows_plan_bytes(const carca_model_params*) { return 0; }
int carca_rows_prepare(void*, float*, const float*, const carca_model_params*, void*) {
  return fail(-5, "rows pipeline: not available under the CPU emulator");
}
int64_t carca_rows_scratch_bytes(const carca_model_params*, int, int) { return 0; }
int carca_rows_eval_forward(float*, int64_t, int, const void*, const carca_model_params*, const int32_t*, const float*,
                            const int32_t*, const float*, int, int, int, int, int, int32_t*, void*, void*) {
  return fail(-5, "rows pipeline: not available under the CPU emulator");
}
}
#else
namespace {

using rows::bf16;
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
inline long long align256(long long x) { return (x + 255) / 256 * 256; }

// byte offsets of the bf16 plan
struct RowsPlan {
  long long tb, mc, w, dw, tqb, tw, mcq, mcw, total;
};
RowsPlan rows_plan(const carca_model_params* m) {
  RowsPlan p;
  const long long d = m->embed.d, n = m->embed.n_items;
  const bool ca = m->decoder_kind == 1;
  p.tb = 0;
  p.mc = align256(p.tb + n * d * 2);
  p.w = align256(p.mc + d * 8 * 4);
  p.dw = align256(p.w + (long long)m->n_blocks * 5 * d * d * 2);
  p.tqb = align256(p.dw + (ca ? 2 * d * d * 2 : 0));
  p.tw = align256(p.tqb + (ca ? n * d * 2 : 0));
  p.mcq = align256(p.tw + (ca ? n * 4 : 0));
  p.mcw = align256(p.mcq + (ca ? d * 8 * 4 : 0));
  p.total = align256(p.mcw + (ca ? 8 * 4 : 0));
  return p;
}

// byte offsets of the per-call scratch
struct RowsScratch {
  long long counters, row_src, row_seg, useg, XA, QA, S2A, F1A, QN, S2, Qb, Kb, Vb, U, KM, total, Rp;
};
RowsScratch rows_scratch(const carca_model_params* m, int B, int L) {
  RowsScratch s;
  const long long d = m->embed.d, H = m->n_heads;
  const long long Rp = (ceil_div_ll((long long)B * L, rows::TILE) + 1) * rows::TILE;
  s.Rp = Rp;
  s.counters = 0;
  s.row_src = 256;
  s.row_seg = align256(s.row_src + Rp * 4);
  s.useg = align256(s.row_seg + Rp * 4);
  s.XA = align256(s.useg + (long long)B * 8);
  s.QA = align256(s.XA + Rp * d * 2);
  s.S2A = align256(s.QA + Rp * d * 2);
  s.F1A = align256(s.S2A + Rp * d * 2);
  s.QN = align256(s.F1A + Rp * d * 2);
  s.S2 = align256(s.QN + Rp * d * 4);
  s.Qb = align256(s.S2 + Rp * d * 4);
  s.Kb = align256(s.Qb + Rp * d * 2);
  s.Vb = align256(s.Kb + Rp * d * 2);
  s.U = align256(s.Vb + Rp * d * 2);
  s.KM = align256(s.U + Rp * H * 4);
  s.total = align256(s.KM + Rp * H * 8 * 4);
  return s;
}

bool rows_shape_ok(const carca_model_params* m) {
  const int d = m->embed.d, H = m->n_heads;
  return (d == 64 || d == 256) && (H == 1 || H == 2 || H == 4 || H == 8) && (d / H == 32 || d / H == 64) && m->embed.n_ctx <= 8 &&
         m->n_blocks >= 1 && m->n_blocks <= 8;
}

template <int D>
int launch_gemm_rows(const rows::GemmArgs& g, cudaStream_t st) {
  auto k = rows::rows_gemm_kernel<D>;
  const size_t smem = rows::GemmCfg<D>::SMEM;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "rows_gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  const int per_sm = D == 256 ? 1 : 3;
  const int slots = max(1, 148 * per_sm / g.n_jobs);
  CARCA_LAUNCH(k, dim3(slots * g.n_jobs), dim3(rows::GEMM_THREADS), smem, st, g);
  return check_launch("rows_gemm");
}

template <int D, int H>
int forward_t(float* y, int64_t ldy, int col0, const unsigned char* plan, const carca_model_params* m,
              const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L, int T,
              int ctx_per_user, int cat_lo, int32_t* status, unsigned char* scr, cudaStream_t st) {
  const RowsPlan pl = rows_plan(m);
  const RowsScratch sc = rows_scratch(m, B, L);
  const int C = m->embed.n_ctx;
  int* n_rows = reinterpret_cast<int*>(scr + sc.counters);
  int* row_src = reinterpret_cast<int*>(scr + sc.row_src);
  int* row_seg = reinterpret_cast<int*>(scr + sc.row_seg);
  int2* useg = reinterpret_cast<int2*>(scr + sc.useg);
  bf16* XA = reinterpret_cast<bf16*>(scr + sc.XA);
  bf16* QA = reinterpret_cast<bf16*>(scr + sc.QA);
  bf16* S2A = reinterpret_cast<bf16*>(scr + sc.S2A);
  bf16* F1A = reinterpret_cast<bf16*>(scr + sc.F1A);
  float* QN = reinterpret_cast<float*>(scr + sc.QN);
  float* S2 = reinterpret_cast<float*>(scr + sc.S2);
  bf16* Qb = reinterpret_cast<bf16*>(scr + sc.Qb);
  bf16* Kb = reinterpret_cast<bf16*>(scr + sc.Kb);
  bf16* Vb = reinterpret_cast<bf16*>(scr + sc.Vb);
  float* U = reinterpret_cast<float*>(scr + sc.U);
  float* KM = reinterpret_cast<float*>(scr + sc.KM);
  const bf16* Tb = reinterpret_cast<const bf16*>(plan + pl.tb);
  const float* Mc = reinterpret_cast<const float*>(plan + pl.mc);
  const bf16* W = reinterpret_cast<const bf16*>(plan + pl.w);
  const long long wsz = (long long)D * D;

  cudaMemsetAsync(n_rows, 0, 256, st);
  {
    auto k = rows::rows_pack_kernel;
    CARCA_LAUNCH(k, dim3(ceil_div(B, 8)), dim3(256), 0, st, row_src, row_seg, useg, n_rows, p_x, B, L);
    TRY(check_launch("rows_pack"));
  }
  const int row_grid = 148 * 4;
  {
    rows::EmbedArgs e;
    e.Tb = Tb; e.Mc = Mc; e.pos = m->embed.pos; e.p_x = p_x; e.p_c = p_c; e.row_src = row_src; e.n_rows = n_rows;
    e.ln_g = m->blocks[0].ln1_g; e.ln_b = m->blocks[0].ln1_b;
    e.XA = XA; e.QA = QA; e.QN = QN; e.L = L; e.C = C;
    auto k = rows::rows_embed_ln_kernel<D>;
    CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, e);
    TRY(check_launch("rows_embed_ln"));
  }
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    const bf16* wb = W + (long long)b * 5 * wsz;
    {
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 3; g.n_rows = n_rows; g.H = H; g.status = status;
      g.job[0].A = QA; g.job[0].W = wb;           g.job[0].bias = bp.bq; g.job[0].epi = rows::EPI_ROWS; g.job[0].out_rows = Qb;
      g.job[1].A = XA; g.job[1].W = wb + wsz;     g.job[1].bias = bp.bk; g.job[1].epi = rows::EPI_ROWS; g.job[1].out_rows = Kb;
      g.job[2].A = XA; g.job[2].W = wb + 2 * wsz; g.job[2].bias = bp.bv; g.job[2].epi = rows::EPI_ROWS; g.job[2].out_rows = Vb;
      TRY(launch_gemm_rows<D>(g, st));
    }
    {
      rows::AttnRowsArgs t;
      t.Q = Qb; t.K = Kb; t.V = Vb; t.QN = QN; t.row_src = row_src; t.row_seg = row_seg; t.n_rows = n_rows;
      t.ln_g = bp.ln2_g; t.ln_b = bp.ln2_b; t.S2 = S2; t.S2A = S2A; t.residual = m->residual_sa;
      auto k = rows::rows_attn_ln_kernel<D, H>;
      CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, t);
      TRY(check_launch("rows_attn_ln"));
    }
    {
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 1; g.n_rows = n_rows; g.H = H; g.status = status;
      g.job[0].A = S2A; g.job[0].W = wb + 3 * wsz; g.job[0].bias = bp.b1; g.job[0].epi = rows::EPI_LRELU_TILE;
      g.job[0].out_tile = F1A;
      TRY(launch_gemm_rows<D>(g, st));
    }
    {
      const bool last = b + 1 == m->n_blocks;
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 1; g.n_rows = n_rows; g.H = H; g.status = status;
      rows::GemmJob& j = g.job[0];
      j.A = F1A; j.W = wb + 4 * wsz; j.bias = bp.b2; j.epi = rows::EPI_LN;
      j.resid = m->residual_sa ? S2 : nullptr;
      j.out_tile = last ? nullptr : XA;
      j.ln_g = last ? m->norm_g : m->blocks[b + 1].ln1_g;
      j.ln_b = last ? m->norm_b : m->blocks[b + 1].ln1_b;
      j.out_f32 = QN; j.out_tile2 = QA;
      TRY(launch_gemm_rows<D>(g, st));
    }
  }
  rows::DecodeArgs d;
  memset(&d, 0, sizeof(d));
  d.useg = useg; d.row_src = row_src; d.o_x = o_x; d.o_c = o_c;
  d.oc_user = ctx_per_user ? C : (long long)T * C;
  d.oc_tgt = ctx_per_user ? 0 : C;
  d.y = y; d.ldy = ldy; d.col0 = col0; d.B = B; d.T = T; d.C = C; d.L = L; d.cat_lo = cat_lo;
  d.residual_ca = m->residual_ca;
  const long long items = (long long)B * ceil_div(T, 128);
  const int dgrid = (int)min(items, 148ll * 16);
  if (m->decoder_kind == 1) {
    const bf16* dw = reinterpret_cast<const bf16*>(plan + pl.dw);
    rows::GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.n_jobs = 2; g.n_rows = n_rows; g.H = H; g.status = status;
    g.job[0].A = QA; g.job[0].W = dw;       g.job[0].bias = m->cross.bk; g.job[0].epi = rows::EPI_KDEC;
    g.job[0].out_rows = Kb; g.job[0].McQ = reinterpret_cast<const float*>(plan + pl.mcq); g.job[0].KM = KM;
    g.job[1].A = QA; g.job[1].W = dw + wsz; g.job[1].bias = m->cross.bv; g.job[1].epi = rows::EPI_VDOT;
    g.job[1].wf = m->cross.wf; g.job[1].U = U;
    TRY(launch_gemm_rows<D>(g, st));
    d.Kd = Kb; d.U = U; d.KM = KM;
    d.TQb = reinterpret_cast<const bf16*>(plan + pl.tqb);
    d.tw = reinterpret_cast<const float*>(plan + pl.tw);
    d.mcw = reinterpret_cast<const float*>(plan + pl.mcw);
    d.dbf = m->cross.bf;
    auto k = rows::rows_decode_ca_kernel<D, H>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(128), 0, st, d);
    TRY(check_launch("rows_decode_ca"));
  } else {
    d.PE = QN; d.Tb = Tb; d.Mc = Mc;
    auto k = rows::rows_decode_dot_kernel<D>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(128), 0, st, d);
    TRY(check_launch("rows_decode_dot"));
  }
  return 0;
}

}  // namespace

extern "C" {

int64_t carca_rows_plan_bytes(const carca_model_params* m) { return rows_plan(m).total; }

int carca_rows_prepare(void* plan_v, float* tmp, const float* plan_f32, const carca_model_params* m, void* stream) {
  cudaStream_t st = S(stream);
  CARCA_REQUIRE(rows_shape_ok(m), "rows_prepare: needs d in {64, 256}, head width 32 or 64, C <= 8, 1..8 blocks "
                                  "(got d=%d H=%d C=%d blocks=%d)", m->embed.d, m->n_heads, m->embed.n_ctx, m->n_blocks);
  unsigned char* plan = reinterpret_cast<unsigned char*>(plan_v);
  const RowsPlan pl = rows_plan(m);
  const PlanLayout fl = plan_layout(m);
  const int d = m->embed.d;
  const long long n = m->embed.n_items;
  const float* T = plan_f32 + fl.tfold;
  const float* Mc = plan_f32 + fl.mc;
  auto cvt = rows::to_bf16_kernel;
  CARCA_LAUNCH(cvt, dim3((unsigned)ceil_div_ll(n * d, 1024)), dim3(256), 0, st, reinterpret_cast<bf16*>(plan + pl.tb), T,
               n * d);
  TRY(check_launch("to_bf16(T)"));
  cudaMemcpyAsync(plan + pl.mc, Mc, sizeof(float) * d * 8, cudaMemcpyDeviceToDevice, st);
  auto pk = rows::pack_weight_bf16_kernel;
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    const float* ws[5] = {bp.wq, bp.wk, bp.wv, bp.w1, bp.w2};
    for (int i = 0; i < 5; ++i) {
      CARCA_LAUNCH(pk, dim3(ceil_div(d * d, 256)), dim3(256), 0, st,
                   reinterpret_cast<bf16*>(plan + pl.w) + ((long long)b * 5 + i) * d * d, ws[i], d);
      TRY(check_launch("pack_weight_bf16"));
    }
  }
  if (m->decoder_kind == 1) {
    const float* ws[2] = {m->cross.wk, m->cross.wv};
    for (int i = 0; i < 2; ++i) {
      CARCA_LAUNCH(pk, dim3(ceil_div(d * d, 256)), dim3(256), 0, st,
                   reinterpret_cast<bf16*>(plan + pl.dw) + (long long)i * d * d, ws[i], d);
      TRY(check_launch("pack_weight_bf16"));
    }
    // candidate-side folds (exact re-associations of linear maps): TQ[i] = WQ T[i] + bq, tw[i] = <T[i], wf>,
    // McQ = WQ Mc, mcw = wf Mc   (src/carca.py:238 with :85-95 folded, :343-345)
    {
      GemmArgs g = gemm_defaults(T, m->cross.wq, tmp, (int)n, d, d);
      g.bias = m->cross.bq;
      TRY(launch_gemm(g, st));
      CARCA_LAUNCH(cvt, dim3((unsigned)ceil_div_ll(n * d, 1024)), dim3(256), 0, st,
                   reinterpret_cast<bf16*>(plan + pl.tqb), tmp, n * d);
      TRY(check_launch("to_bf16(TQ)"));
    }
    {
      GemmArgs g = gemm_defaults(T, m->cross.wf, reinterpret_cast<float*>(plan + pl.tw), (int)n, 1, d);
      TRY(launch_gemm(g, st));
    }
    {
      GemmArgs g = gemm_defaults(m->cross.wq, Mc, reinterpret_cast<float*>(plan + pl.mcq), d, 8, d);
      g.transB = 0; g.ldb = 8; g.ldc = 8;
      TRY(launch_gemm(g, st));
      GemmArgs gw = gemm_defaults(m->cross.wf, Mc, reinterpret_cast<float*>(plan + pl.mcw), 1, 8, d);
      gw.transB = 0; gw.ldb = 8; gw.ldc = 8;
      TRY(launch_gemm(gw, st));
    }
  }
  return 0;
}

int64_t carca_rows_scratch_bytes(const carca_model_params* m, int B, int L) { return rows_scratch(m, B, L).total; }

int carca_rows_eval_forward(float* y, int64_t ldy, int col0, const void* plan, const carca_model_params* m,
                            const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                            int T, int ctx_per_user, int cat_lo, int32_t* status, void* scratch, void* stream) {
  CARCA_REQUIRE(rows_shape_ok(m), "rows_eval_forward: needs d in {64, 256}, head width 32 or 64, C <= 8, 1..8 "
                                  "blocks (got d=%d H=%d C=%d blocks=%d)", m->embed.d, m->n_heads, m->embed.n_ctx,
                m->n_blocks);
  CARCA_REQUIRE(L >= 1 && L <= 256, "rows_eval_forward: L=%d outside 1..256", L);
  CARCA_REQUIRE(status != nullptr && scratch != nullptr, "rows_eval_forward: status and scratch are required");
  CARCA_REQUIRE((long long)B < (1ll << 23), "rows_eval_forward: at most 2^23 users per call");
  if (m->embed.pos) CARCA_REQUIRE(L <= m->embed.pos_len, "rows_eval_forward: sequence length %d > positional table %d", L,
                                  m->embed.pos_len);
  if (B <= 0 || T <= 0) return 0;
  const unsigned char* pl = reinterpret_cast<const unsigned char*>(plan);
  unsigned char* scr = reinterpret_cast<unsigned char*>(scratch);
  cudaStream_t st = S(stream);
  const int d = m->embed.d, H = m->n_heads;
#define ROWS_CASE(DD, HH)                                                                                          \
  if (d == DD && H == HH)                                                                                          \
    return forward_t<DD, HH>(y, ldy, col0, pl, m, p_x, p_c, o_x, o_c, B, L, T, ctx_per_user, cat_lo, status, scr, st);
  ROWS_CASE(64, 1) ROWS_CASE(64, 2) ROWS_CASE(256, 4) ROWS_CASE(256, 8)
#undef ROWS_CASE
  return fail(-4, "rows_eval_forward: unsupported (d, heads) = (%d, %d)", d, H);
}

}  // extern "C"
#endif
