// extern "C" entry points of the bf16 packed-rows inference pipeline (rows_bf16.cuh); declared in
// include/carca_b200.h.  Reference path replaced: CARCA.forward in eval mode, src/carca.py:411-431.
#include "../../include/carca_b200.h"

#include <cmath>
#include <cstdlib>
#include <utility>

#include "common.cuh"
#include "gemm.cuh"
#include "plan_layout.cuh"
#include "rows_bf16.cuh"
#include "rows_attn_tc.cuh"
#include "rows_ffn_chain.cuh"
#include "catalog_tc.cuh"

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
int carca_rows_prepare(void*, const float*, const carca_model_params*, void*) {
  return fail(-5, "rows pipeline: not available under the CPU emulator");
}
int64_t carca_rows_scratch_bytes(const carca_model_params*, int, int) { return 0; }
int carca_rows_set_stage_events(void* const*, int) { return 0; }
int carca_rows_stage_ids(int32_t*, int) { return 0; }
int64_t carca_rows_catalog_scratch_bytes(const carca_model_params*, int, int) { return 0; }
int carca_rows_catalog_counts(int32_t*, const void*, const float*, const carca_model_params*, const int32_t*, const float*,
                              const float*, const int32_t*, int, int, int, int, int32_t*, void*, void*) {
  return fail(-5, "rows pipeline: not available under the CPU emulator");
}
int carca_rows_eval_forward(float*, int64_t, int, const void*, const float*, const carca_model_params*, const int32_t*,
                            const float*, const int32_t*, const float*, int, int, int, int, int, int, int32_t*, void*, void*) {
  return fail(-5, "rows pipeline: not available under the CPU emulator");
}
}
#else
namespace {

using rows::bf16;
inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Launch with programmatic stream serialization (see pdl_wait in rows_bf16.cuh): the kernel may begin while the previous
// kernel of the stream drains.  Only for kernels that follow a kernel of this pipeline and call pdl_wait() before they
// touch anything a predecessor wrote.  OFF unless CARCA_ROWS_PDL is set: measured on a B200 (8192 Beauty users, bf16)
// it helps eager launches (0.251 -> 0.245 ms per step) but not graph replays (0.234 -> 0.239 ms; every position valid
// 0.976 -> 1.010 ms) — the early CTAs hold shared memory and TMEM while they wait — and replays are how the step runs.
template <typename... KArgs, typename... Args>
void launch_pdl(void (*k)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  static const bool off = getenv("CARCA_ROWS_PDL") == nullptr;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = off ? 0 : 1;
  cudaLaunchKernelEx(&cfg, k, std::forward<Args>(args)...);
}

// measurement aid (carca_rows_set_stage_events): the bf16 forward records the caller's CUDA events between its kernels
enum { ST_START = 0, ST_PACK, ST_EMBED, ST_QKV, ST_ATTN, ST_FFN1, ST_FFN2, ST_CHAIN, ST_DEC_KV, ST_DECODER };
cudaEvent_t g_stage_ev[64];
int g_stage_id[64];
int g_stage_n = 0, g_stage_used = 0;
inline void stage_mark(int id, cudaStream_t st) {
  if (g_stage_used < g_stage_n) {
    cudaEventRecord(g_stage_ev[g_stage_used], st);
    g_stage_id[g_stage_used++] = id;
  }
}
inline long long align256(long long x) { return (x + 255) / 256 * 256; }

// byte offsets of the bf16 plan
struct RowsPlan {
  long long mc, w, dw, tq, tw, mcq, mcw, tqb, wkv, bkv, total;   // wkv: fp32 [Wk; Wv] per block (+ decoder), bkv biases
};
RowsPlan rows_plan(const carca_model_params* m) {
  RowsPlan p;
  const long long d = m->embed.d, n = m->embed.n_items;
  const bool ca = m->decoder_kind == 1;
  p.mc = 0;
  p.w = align256(p.mc + d * 8 * 4);
  p.dw = align256(p.w + (long long)m->n_blocks * 5 * d * d * 2);
  p.tq = align256(p.dw + (ca ? 2 * d * d * 2 : 0));
  p.tw = align256(p.tq + (ca ? n * d * 4 : 0));
  p.mcq = align256(p.tw + (ca ? n * 4 : 0));
  p.mcw = align256(p.mcq + (ca ? d * 8 * 4 : 0));
  p.tqb = align256(p.mcw + (ca ? 8 * 4 : 0));
  p.wkv = align256(p.tqb + (ca ? n * d * 2 : 0));
  p.bkv = align256(p.wkv + (long long)(m->n_blocks + 1) * 2 * d * d * 4);
  p.total = align256(p.bkv + (long long)(m->n_blocks + 1) * 2 * d * 4);
  return p;
}

// byte offsets of the per-call scratch: seven [Rp, d] fp32-sized slots shared by both flavours
struct RowsScratch {
  long long counters, row_src, row_seg, row_id, useg, slot[7], U, KM, total, Rp;
};
RowsScratch rows_scratch(const carca_model_params* m, int B, int L) {
  RowsScratch s;
  const long long d = m->embed.d, H = m->n_heads;
  const long long Rp = (ceil_div_ll((long long)B * L, rows::TILE) + 1) * rows::TILE;
  s.Rp = Rp;
  s.counters = 0;
  s.row_src = 256;
  s.row_seg = align256(s.row_src + Rp * 4);
  s.row_id = align256(s.row_seg + Rp * 4);
  s.useg = align256(s.row_id + Rp * 4);
  long long off = align256(s.useg + (long long)B * 8);
  for (int i = 0; i < 7; ++i) {
    s.slot[i] = off;
    off = align256(off + Rp * d * 4);
  }
  s.U = off;
  s.KM = align256(s.U + Rp * H * 4);
  s.total = align256(s.KM + Rp * H * 8 * 4);
  return s;
}

// bf16 flavour: tcgen05 kind::f16 GEMMs with the weight matrix resident in shared memory
bool rows_shape_ok(const carca_model_params* m) {
  const int d = m->embed.d, H = m->n_heads;
  return (d == 64 || d == 256) && H >= 1 && d % H == 0 && (d / H == 32 || d / H == 64) && m->embed.n_ctx <= 8 &&
         m->n_blocks >= 1 && m->n_blocks <= 8;
}
// fp32 flavour: the general 3xTF32 tcgen05 GEMM (gemm_tc.cuh) over the packed rows
bool rows_shape_ok_f32(const carca_model_params* m) {
  const int d = m->embed.d, H = m->n_heads;
  return (d == 32 || d == 64 || d == 128 || d == 256) && H >= 1 && d % H == 0 &&
         (d / H == 16 || d / H == 32 || d / H == 64) && m->embed.n_ctx <= 8 && m->n_blocks >= 1 && m->n_blocks <= 8;
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
  const int per_sm = D == 256 ? 1 : 2;
  const int slots = max(1, 148 * per_sm / g.n_jobs);
  launch_pdl(k, dim3(slots * g.n_jobs), dim3(rows::GEMM_THREADS), smem, st, g);
  return check_launch("rows_gemm");
}

// tensor-core attention (rows_attn_tc.cuh): the key window of a 128-row tile is at most 127 + L rows
inline int attn_tc_window(int L) { return (127 + L + 15) / 16 * 16; }

template <int D, int H>
int launch_attn_tc(rows::AttnTcArgs& t, int L, cudaStream_t st) {
  using Cfg = rows::AttnTcCfg<D, H>;
  t.win_max = attn_tc_window(L);
  t.s_cols = (D == 64 && t.win_max <= 192) ? 192 : 256;
  t.tmem_cols = t.s_cols + D <= 256 ? 256 : 512;
  if (const char* e = getenv("CARCA_ATTN_VSWAP")) t.vswap = atoi(e);   // (development)
  const size_t smem = Cfg::smem_bytes(t.win_max);
  if (smem > 227 * 1024) return fail(-2, "rows_attn_tc: %zu bytes of shared memory (d=%d, L=%d)", smem, D, L);
  auto k = rows::rows_attn_tc_kernel<D, H>;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "rows_attn_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_smem = smem;
  }
  const int per_sm = (t.tmem_cols == 256 && 2 * (smem + 1024) <= 228 * 1024) ? 2 : 1;
  launch_pdl(k, dim3(148 * per_sm), dim3(rows::AT_THREADS), smem, st, t);
  return check_launch("rows_attn_tc");
}

int launch_ffn_chain(const rows::FfnChainArgs& f, cudaStream_t st) {
  auto k = rows::rows_ffn_chain_kernel;
  const size_t smem = sizeof(rows::FfnChainSmem);
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "rows_ffn_chain: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    attr_set = true;
  }
  launch_pdl(k, dim3(148 * 2), dim3(rows::FC_THREADS), smem, st, f);
  return check_launch("rows_ffn_chain");
}

template <int D, int H>
int forward_t(float* y, int64_t ldy, int col0, const unsigned char* plan, const float* Tf, const carca_model_params* m,
              const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L, int T,
              int ctx_per_user, int cat_lo, int32_t* status, unsigned char* scr, cudaStream_t st) {
  const RowsPlan pl = rows_plan(m);
  const RowsScratch sc = rows_scratch(m, B, L);
  const int C = m->embed.n_ctx;
  int* n_rows = reinterpret_cast<int*>(scr + sc.counters);
  int* row_src = reinterpret_cast<int*>(scr + sc.row_src);
  int* row_seg = reinterpret_cast<int*>(scr + sc.row_seg);
  int* row_id = reinterpret_cast<int*>(scr + sc.row_id);
  int2* useg = reinterpret_cast<int2*>(scr + sc.useg);
  const long long half = sc.Rp * D * 2;
  bf16* XA = reinterpret_cast<bf16*>(scr + sc.slot[0]);
  bf16* QA = reinterpret_cast<bf16*>(scr + sc.slot[0] + half);
  bf16* S2A = reinterpret_cast<bf16*>(scr + sc.slot[1]);
  bf16* F1A = reinterpret_cast<bf16*>(scr + sc.slot[1] + half);
  float* QN = reinterpret_cast<float*>(scr + sc.slot[2]);
  float* S2 = reinterpret_cast<float*>(scr + sc.slot[3]);
  bf16* Qb = reinterpret_cast<bf16*>(scr + sc.slot[4]);
  bf16* Kb = reinterpret_cast<bf16*>(scr + sc.slot[4] + half);
  bf16* Vb = reinterpret_cast<bf16*>(scr + sc.slot[5]);
  float* U = reinterpret_cast<float*>(scr + sc.U);
  float* KM = reinterpret_cast<float*>(scr + sc.KM);
  const float* Mc = reinterpret_cast<const float*>(plan + pl.mc);
  const bf16* W = reinterpret_cast<const bf16*>(plan + pl.w);
  const long long wsz = (long long)D * D;

  g_stage_used = 0;
  stage_mark(ST_START, st);
  cudaMemsetAsync(n_rows, 0, 256, st);
  {
    auto k = rows::rows_pack_kernel;
    CARCA_LAUNCH(k, dim3(ceil_div(B, 8)), dim3(256), 0, st, row_src, row_seg, useg, n_rows, p_x, B, L, row_id);
    TRY(check_launch("rows_pack"));
  }
  stage_mark(ST_PACK, st);
  const int row_grid = 148 * 4;
  {
    rows::EmbedArgs e;
    e.T = Tf; e.Mc = Mc; e.pos = m->embed.pos; e.p_x = p_x; e.p_c = p_c; e.row_src = row_src; e.n_rows = n_rows; e.row_id = row_id;
    e.ln_g = m->blocks[0].ln1_g; e.ln_b = m->blocks[0].ln1_b;
    e.XA = XA; e.QA = QA; e.QN = QN; e.L = L; e.C = C;
    auto k = rows::rows_embed_ln_kernel<D, false>;
    launch_pdl(k, dim3(row_grid), dim3(256), 0, st, e);
    TRY(check_launch("rows_embed_ln"));
  }
  stage_mark(ST_EMBED, st);
  const bool attn_tc = attn_tc_window(L) <= 256 && !getenv("CARCA_ROWS_ATTN_FFMA");
  const bool chain = D == 64 && attn_tc && !getenv("CARCA_ROWS_NO_CHAIN");
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    const bf16* wb = W + (long long)b * 5 * wsz;
    if (attn_tc) {
      // Q / K / V written as the attention's tensor-core operands, then the tcgen05 attention (rows_attn_tc.cuh)
      if (!(chain && b > 0)) {      // (chained: the previous block's tail kernel has produced them)
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 3; g.n_rows = n_rows; g.H = H; g.status = status;
      g.job[0].A = QA; g.job[0].W = wb;           g.job[0].bias = bp.bq; g.job[0].epi = rows::EPI_BIAS_TILE; g.job[0].out_tile = Qb;
      g.job[1].A = XA; g.job[1].W = wb + wsz;     g.job[1].bias = bp.bk; g.job[1].epi = rows::EPI_KMAJ; g.job[1].out_tile = Kb;
      g.job[2].A = XA; g.job[2].W = wb + 2 * wsz; g.job[2].bias = bp.bv; g.job[2].epi = rows::EPI_VMN; g.job[2].out_tile = Vb;
      g.job[1].ld_rows = g.job[2].ld_rows = sc.Rp;
      TRY(launch_gemm_rows<D>(g, st));
      stage_mark(ST_QKV, st);
      }
      rows::AttnTcArgs t;
      memset(&t, 0, sizeof(t));
      t.Qt = Qb; t.Kk = Kb; t.Vm = Vb; t.Rp = sc.Rp; t.QN = QN; t.row_src = row_src; t.row_seg = row_seg; t.n_rows = n_rows;
      t.ln_g = bp.ln2_g; t.ln_b = bp.ln2_b; t.S2 = S2; t.S2A = S2A; t.residual = m->residual_sa; t.status = status;
      TRY((launch_attn_tc<D, H>(t, L, st)));
      stage_mark(ST_ATTN, st);
    } else {
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 3; g.n_rows = n_rows; g.H = H; g.status = status;
      g.job[0].A = QA; g.job[0].W = wb;           g.job[0].bias = bp.bq; g.job[0].epi = rows::EPI_ROWS; g.job[0].out_rows = Qb;
      g.job[1].A = XA; g.job[1].W = wb + wsz;     g.job[1].bias = bp.bk; g.job[1].epi = rows::EPI_ROWS; g.job[1].out_rows = Kb;
      g.job[2].A = XA; g.job[2].W = wb + 2 * wsz; g.job[2].bias = bp.bv; g.job[2].epi = rows::EPI_ROWS; g.job[2].out_rows = Vb;
      TRY(launch_gemm_rows<D>(g, st));
      stage_mark(ST_QKV, st);
    }
    if (!attn_tc) {
      rows::AttnRowsArgs t;
      t.Q = Qb; t.K = Kb; t.V = Vb; t.ldkv = D; t.QN = QN; t.row_src = row_src; t.row_seg = row_seg; t.n_rows = n_rows;
      t.ln_g = bp.ln2_g; t.ln_b = bp.ln2_b; t.S2 = S2; t.S2A = S2A; t.residual = m->residual_sa;
      auto k = rows::rows_attn_ln_kernel<D, H, false>;
      CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, t);
      TRY(check_launch("rows_attn_ln"));
      stage_mark(ST_ATTN, st);
    }
    if (chain) {
      // FFN-1 -> FFN-2 + next LayerNorm -> next block's Q / K / V (or the decoder's K / V): one kernel, the intermediate
      // operands stay in shared memory (rows_ffn_chain.cuh)
      const bool last = b + 1 == m->n_blocks;
      rows::FfnChainArgs f;
      memset(&f, 0, sizeof(f));
      f.A = S2A; f.W12 = wb + 3 * wsz; f.b1 = bp.b1; f.b2 = bp.b2;
      f.resid = m->residual_sa ? S2 : nullptr;
      f.ln_g = last ? m->norm_g : m->blocks[b + 1].ln1_g;
      f.ln_b = last ? m->norm_b : m->blocks[b + 1].ln1_b;
      f.out_f32 = QN; f.n_rows = n_rows; f.H = H; f.status = status;
      if (!last) {
        const carca_block_params& nb = m->blocks[b + 1];
        f.tail = 1; f.W3 = wb + 5 * wsz; f.bq = nb.bq; f.bk = nb.bk; f.bv = nb.bv;
        f.Qt = Qb; f.Kk = Kb; f.Vm = Vb; f.ld_rows = sc.Rp;
      } else if (m->decoder_kind == 1) {
        f.tail = 2; f.W3 = reinterpret_cast<const bf16*>(plan + pl.dw); f.bk = m->cross.bk; f.bv = m->cross.bv;
        f.Kd = S2; f.McQ = reinterpret_cast<const float*>(plan + pl.mcq); f.KM = KM; f.wf = m->cross.wf; f.U = U;
      }
      TRY(launch_ffn_chain(f, st));
      stage_mark(ST_CHAIN, st);
      continue;
    }
    {
      rows::GemmArgs g;
      memset(&g, 0, sizeof(g));
      g.n_jobs = 1; g.n_rows = n_rows; g.H = H; g.status = status;
      g.job[0].A = S2A; g.job[0].W = wb + 3 * wsz; g.job[0].bias = bp.b1; g.job[0].epi = rows::EPI_LRELU_TILE;
      g.job[0].out_tile = F1A;
      TRY(launch_gemm_rows<D>(g, st));
      stage_mark(ST_FFN1, st);
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
      stage_mark(ST_FFN2, st);
    }
  }
  rows::DecodeArgs d;
  memset(&d, 0, sizeof(d));
  d.useg = useg; d.row_src = row_src; d.o_x = o_x; d.o_c = o_c;
  d.oc_user = ctx_per_user ? C : (long long)T * C;
  d.oc_tgt = ctx_per_user ? 0 : C;
  d.y = y; d.ldy = ldy; d.col0 = col0; d.B = B; d.T = T; d.C = C; d.L = L; d.cat_lo = cat_lo;
  d.residual_ca = m->residual_ca;
  const long long items = (long long)B * ceil_div(T, m->decoder_kind == 1 ? rows::DecCfg<D>::threads(H) / H : 128);
  const int dgrid = (int)min(items, 148ll * 16);
  if (m->decoder_kind == 1) {
    const bf16* dw = reinterpret_cast<const bf16*>(plan + pl.dw);
    rows::GemmArgs g;
    memset(&g, 0, sizeof(g));
    g.n_jobs = 2; g.n_rows = n_rows; g.H = H; g.status = status;
    g.job[0].A = QA; g.job[0].W = dw;       g.job[0].bias = m->cross.bk; g.job[0].epi = rows::EPI_KDEC;
    g.job[0].out_f32 = S2; g.job[0].McQ = reinterpret_cast<const float*>(plan + pl.mcq); g.job[0].KM = KM;
    g.job[1].A = QA; g.job[1].W = dw + wsz; g.job[1].bias = m->cross.bv; g.job[1].epi = rows::EPI_VDOT;
    g.job[1].wf = m->cross.wf; g.job[1].U = U;
    if (!chain) {      // (chained: the last block's tail kernel has produced them)
      TRY(launch_gemm_rows<D>(g, st));
      stage_mark(ST_DEC_KV, st);
    }
    (void)Kb;
    if (H <= 4 && !getenv("CARCA_ROWS_FFMA_DECODE")) {   // tensor-core decoder (rows_decode_tc_kernel)
      rows::DecTcArgs t;
      memset(&t, 0, sizeof(t));
      t.TQb = reinterpret_cast<const bf16*>(plan + pl.tqb);
      t.Kd = S2; t.U = U; t.KM = KM;
      t.tw = reinterpret_cast<const float*>(plan + pl.tw);
      t.mcw = reinterpret_cast<const float*>(plan + pl.mcw);
      t.dbf = m->cross.bf;
      t.useg = useg; t.row_src = row_src; t.o_x = o_x; t.o_c = o_c;
      t.oc_user = d.oc_user; t.oc_tgt = d.oc_tgt;
      t.y = y; t.ldy = ldy; t.col0 = col0; t.B = B; t.T = T; t.C = C; t.cat_lo = cat_lo; t.residual_ca = m->residual_ca;
      t.status = status;
      auto k = rows::rows_decode_tc_kernel<D, (H <= 4 ? H : 1)>;
      const size_t smem = sizeof(rows::DecTcSmem<D, (H <= 4 ? H : 1)>);
      static bool set = false;
      if (!set) {
        cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return fail(-3, "rows_decode_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        set = true;
      }
      const long long work = (long long)B * ceil_div(T, 128);
      const int per_sm = D == 256 ? 1 : 2;
      launch_pdl(k, dim3((unsigned)min(work, 148ll * per_sm)), dim3(rows::DT_THREADS), smem, st, t);
      TRY(check_launch("rows_decode_tc"));
      stage_mark(ST_DECODER, st);
      return 0;
    }
    d.Kd = S2; d.ldk = D; d.U = U; d.KM = KM;   // (the S2 buffer is free after the last block: it holds the fp32 keys)
    d.TQ = reinterpret_cast<const float*>(plan + pl.tq);
    d.tw = reinterpret_cast<const float*>(plan + pl.tw);
    d.mcw = reinterpret_cast<const float*>(plan + pl.mcw);
    d.dbf = m->cross.bf;
    auto k = rows::rows_decode_ca_kernel<D, H, false>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(rows::DecCfg<D>::threads(H)), 0, st, d);
    TRY(check_launch("rows_decode_ca"));
  } else {
    d.PE = QN; d.Tf = Tf; d.Mc = Mc;
    auto k = rows::rows_decode_dot_kernel<D>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(128), 0, st, d);
    TRY(check_launch("rows_decode_dot"));
  }
  stage_mark(ST_DECODER, st);
  return 0;
}

// fp32 flavour of the same pipeline: fp32 rows everywhere, every projection through the general 3xTF32 tcgen05 GEMM
// (gemm_tc.cuh) with the row count read on the device (GemmArgs::m_dev), LayerNorms as row kernels.  No per-user row
// limit: this is the fp32 path for windows longer than one 64-row bin and for widths the fused kernels do not cover.
int gemm_rows_f32(float* C, const float* A, const float* W, const float* bias, int Mcap, int d, const int* n_rows, int act,
                  const float* R, cudaStream_t st, int n_out = 0) {
  GemmArgs g = gemm_defaults(A, W, C, Mcap, n_out ? n_out : d, d);
  g.bias = bias;
  g.act = act;
  g.R = R;
  g.ldr = d;
  g.m_dev = n_rows;
  return launch_gemm(g, st);
}

// encoder half: ids -> final LayerNorm rows (QN) and, for the cross-attention decoder, its key / value rows (Kf, Vf)
template <int D, int H>
int encode_f32_t(const unsigned char* plan, const float* Tf, const carca_model_params* m, const int32_t* p_x,
                 const float* p_c, int B, int L, unsigned char* scr, cudaStream_t st) {
  const RowsPlan pl = rows_plan(m);
  const RowsScratch sc = rows_scratch(m, B, L);
  const int C = m->embed.n_ctx;
  int* n_rows = reinterpret_cast<int*>(scr + sc.counters);
  int* row_src = reinterpret_cast<int*>(scr + sc.row_src);
  int* row_seg = reinterpret_cast<int*>(scr + sc.row_seg);
  int* row_id = reinterpret_cast<int*>(scr + sc.row_id);
  int2* useg = reinterpret_cast<int2*>(scr + sc.useg);
  float* Xf = reinterpret_cast<float*>(scr + sc.slot[0]);
  float* QN = reinterpret_cast<float*>(scr + sc.slot[1]);
  float* S2 = reinterpret_cast<float*>(scr + sc.slot[2]);
  float* Qf = reinterpret_cast<float*>(scr + sc.slot[3]);
  float* Kf = reinterpret_cast<float*>(scr + sc.slot[4]);
  float* Vf = reinterpret_cast<float*>(scr + sc.slot[5]);
  float* F1 = reinterpret_cast<float*>(scr + sc.slot[6]);
  const float* Mc = reinterpret_cast<const float*>(plan + pl.mc);
  const float* wkv = reinterpret_cast<const float*>(plan + pl.wkv);
  const float* bkv = reinterpret_cast<const float*>(plan + pl.bkv);
  const int Mcap = (int)min((long long)B * L, (long long)INT32_MAX);

  cudaMemsetAsync(n_rows, 0, 256, st);
  {
    auto k = rows::rows_pack_kernel;
    CARCA_LAUNCH(k, dim3(ceil_div(B, 8)), dim3(256), 0, st, row_src, row_seg, useg, n_rows, p_x, B, L, row_id);
    TRY(check_launch("rows_pack"));
  }
  const int row_grid = 148 * 4;
  {
    rows::EmbedArgs e;
    memset(&e, 0, sizeof(e));
    e.T = Tf; e.Mc = Mc; e.pos = m->embed.pos; e.p_x = p_x; e.p_c = p_c; e.row_src = row_src; e.n_rows = n_rows; e.row_id = row_id;
    e.ln_g = m->blocks[0].ln1_g; e.ln_b = m->blocks[0].ln1_b;
    e.Xf = Xf; e.QN = QN; e.L = L; e.C = C;
    auto k = rows::rows_embed_ln_kernel<D, true>;
    CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, e);
    TRY(check_launch("rows_embed_ln"));
  }
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    TRY(gemm_rows_f32(Qf, QN, bp.wq, bp.bq, Mcap, D, n_rows, 0, nullptr, st));   // Q from LN1(x), K / V from x (:238-240)
    TRY(gemm_rows_f32(Kf, Xf, wkv + (long long)b * 2 * D * D, bkv + (long long)b * 2 * D, Mcap, D, n_rows, 0, nullptr, st,
                      2 * D));                                                     // [K | V] rows of 2 D floats
    {
      rows::AttnRowsArgs t;
      memset(&t, 0, sizeof(t));
      t.Q = Qf; t.K = Kf; t.V = Kf + D; t.ldkv = 2 * D; t.QN = QN; t.row_src = row_src; t.row_seg = row_seg;
      t.n_rows = n_rows;
      t.ln_g = bp.ln2_g; t.ln_b = bp.ln2_b; t.S2 = S2; t.residual = m->residual_sa;
      auto k = rows::rows_attn_ln_kernel<D, H, true>;
      CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, t);
      TRY(check_launch("rows_attn_ln"));
    }
    TRY(gemm_rows_f32(F1, S2, bp.w1, bp.b1, Mcap, D, n_rows, 1, nullptr, st));                              // :307-308
    TRY(gemm_rows_f32(Xf, F1, bp.w2, bp.b2, Mcap, D, n_rows, 0, m->residual_sa ? S2 : nullptr, st));       // :311, :316
    {
      const bool last = b + 1 == m->n_blocks;
      auto k = rows::rows_ln_kernel<D>;
      CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, QN, (const float*)Xf, last ? m->norm_g : m->blocks[b + 1].ln1_g,
                   last ? m->norm_b : m->blocks[b + 1].ln1_b, (const int*)n_rows);
      TRY(check_launch("rows_ln"));
    }
  }
  if (m->decoder_kind == 1) {   // decoder keys | values (:239-240 with the encoded profile) and their folds
    TRY(gemm_rows_f32(Kf, QN, wkv + (long long)m->n_blocks * 2 * D * D, bkv + (long long)m->n_blocks * 2 * D, Mcap, D, n_rows,
                      0, nullptr, st, 2 * D));
    auto k = rows::rows_fold_kv_kernel<D, H>;
    CARCA_LAUNCH(k, dim3(row_grid), dim3(256), 0, st, reinterpret_cast<float*>(scr + sc.U), reinterpret_cast<float*>(scr + sc.KM),
                 (const float*)Kf, (const float*)(Kf + D), 2 * D, m->cross.wf, reinterpret_cast<const float*>(plan + pl.mcq),
                 (const int*)n_rows);
    TRY(check_launch("rows_fold_kv"));
  }
  (void)useg; (void)Qf; (void)Vf;
  return 0;
}

// decoder half over the (user, candidate) rows
template <int D, int H>
int decode_f32_t(float* y, int64_t ldy, int col0, const unsigned char* plan, const float* Tf, const carca_model_params* m,
                 const int32_t* o_x, const float* o_c, int B, int L, int T, int ctx_per_user, int cat_lo, unsigned char* scr,
                 cudaStream_t st) {
  const RowsPlan pl = rows_plan(m);
  const RowsScratch sc = rows_scratch(m, B, L);
  const int C = m->embed.n_ctx;
  int* row_src = reinterpret_cast<int*>(scr + sc.row_src);
  int2* useg = reinterpret_cast<int2*>(scr + sc.useg);
  float* QN = reinterpret_cast<float*>(scr + sc.slot[1]);
  float* Kf = reinterpret_cast<float*>(scr + sc.slot[4]);       // [K | V] rows of 2 d floats (slots 4 and 5)
  const float* Mc = reinterpret_cast<const float*>(plan + pl.mc);
  rows::DecodeArgs d;
  memset(&d, 0, sizeof(d));
  d.useg = useg; d.row_src = row_src; d.o_x = o_x; d.o_c = o_c;
  d.oc_user = ctx_per_user ? C : (long long)T * C;
  d.oc_tgt = ctx_per_user ? 0 : C;
  d.y = y; d.ldy = ldy; d.col0 = col0; d.B = B; d.T = T; d.C = C; d.L = L; d.cat_lo = cat_lo;
  d.residual_ca = m->residual_ca;
  const long long items = (long long)B * ceil_div(T, m->decoder_kind == 1 ? rows::DecCfg<D>::threads(H) / H : 128);
  const int dgrid = (int)min(items, 148ll * 16);
  if (m->decoder_kind == 1) {
    d.Kd = Kf; d.ldk = 2 * D;
    d.U = reinterpret_cast<const float*>(scr + sc.U); d.KM = reinterpret_cast<const float*>(scr + sc.KM);
    d.TQ = reinterpret_cast<const float*>(plan + pl.tq);
    d.tw = reinterpret_cast<const float*>(plan + pl.tw);
    d.mcw = reinterpret_cast<const float*>(plan + pl.mcw);
    d.dbf = m->cross.bf;
    auto k = rows::rows_decode_ca_kernel<D, H, true>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(rows::DecCfg<D>::threads(H)), 0, st, d);
    TRY(check_launch("rows_decode_ca"));
  } else {
    d.PE = QN; d.Tf = Tf; d.Mc = Mc;
    auto k = rows::rows_decode_dot_kernel<D>;
    CARCA_LAUNCH(k, dim3(dgrid), dim3(128), 0, st, d);
    TRY(check_launch("rows_decode_dot"));
  }
  return 0;
}

template <int D, int H>
int forward_f32_t(float* y, int64_t ldy, int col0, const unsigned char* plan, const float* Tf, const carca_model_params* m,
                  const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L, int T,
                  int ctx_per_user, int cat_lo, unsigned char* scr, cudaStream_t st) {
  TRY((encode_f32_t<D, H>(plan, Tf, m, p_x, p_c, B, L, scr, st)));
  return decode_f32_t<D, H>(y, ldy, col0, plan, Tf, m, o_x, o_c, B, L, T, ctx_per_user, cat_lo, scr, st);
}

// ---- full-catalog rank counts on the tensor cores (catalog_tc.cuh), d = 64
struct CatScratch {
  long long ucol, cw, ypos, meta, bt, total;
  long long gmax;
};
CatScratch cat_scratch(const carca_model_params* m, int B, int L) {
  CatScratch c;
  const long long base = rows_scratch(m, B, L).total;
  c.gmax = m->decoder_kind == 1 ? 2ll * B * L / cat::CT + B / 64 + 2 : (B + cat::CT - 1) / cat::CT;
  c.ucol = base;
  c.cw = align256(c.ucol + 4ll * B);
  c.ypos = align256(c.cw + 4ll * B);
  c.meta = align256(c.ypos + 4ll * B);
  c.bt = align256(c.meta + c.gmax * cat::META_WORDS * 4ll);
  c.total = align256(c.bt + c.gmax * cat::GROUP_FLOATS * 4ll);
  return c;
}

template <int H>
int catalog_counts_t(int32_t* counts, const unsigned char* plan, const float* Tf, const carca_model_params* m,
                     const int32_t* p_x, const float* p_c, const float* ctx_user, const int32_t* pos_item, int item_lo,
                     int n_shard, int B, int L, int32_t* status, unsigned char* scr, cudaStream_t st) {
  constexpr int D = 64;
  const RowsPlan pl = rows_plan(m);
  const RowsScratch sc = rows_scratch(m, B, L);
  const CatScratch cs = cat_scratch(m, B, L);
  const int C = m->embed.n_ctx;
  const bool ca = m->decoder_kind == 1;
  TRY((encode_f32_t<D, H>(plan, Tf, m, p_x, p_c, B, L, scr, st)));
  int* counters = reinterpret_cast<int*>(scr + sc.counters);     // [0] rows, [1] column groups
  float* y_pos = reinterpret_cast<float*>(scr + cs.ypos);
  // the positive's own probability, by the same decoder arithmetic as the per-candidate path
  TRY((decode_f32_t<D, H>(y_pos, 1, 0, plan, Tf, m, pos_item, ctx_user, B, L, 1, 1, 0, scr, st)));
  int* ucol = reinterpret_cast<int*>(scr + cs.ucol);
  float* meta = reinterpret_cast<float*>(scr + cs.meta);
  float* Bt = reinterpret_cast<float*>(scr + cs.bt);
  cat::FillArgs f;
  memset(&f, 0, sizeof(f));
  f.Bt = Bt; f.meta = meta; f.ucol = ucol; f.useg = reinterpret_cast<const int2*>(scr + sc.useg);
  f.row_src = reinterpret_cast<const int*>(scr + sc.row_src); f.n_rows = counters;
  f.Kd = reinterpret_cast<const float*>(scr + sc.slot[4]); f.Vd = f.Kd + D; f.ld = 2 * D;   // [K | V] rows of the fused GEMM
  f.wf = m->cross.wf; f.McQ = reinterpret_cast<const float*>(plan + pl.mcq); f.ctx_user = ctx_user;
  f.PE = reinterpret_cast<const float*>(scr + sc.slot[1]); f.Mc = reinterpret_cast<const float*>(plan + pl.mc);
  f.mcw = reinterpret_cast<const float*>(plan + pl.mcw); f.y_pos = y_pos; f.pos_item = pos_item;
  f.B = B; f.C = C; f.sc = 1.4426950408889634f / sqrtf((float)(D / H));
  {
    auto k = cat::cat_clear_meta_kernel;
    const long long n = cs.gmax * cat::CT;
    CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, st, meta, n);
    TRY(check_launch("cat_clear_meta"));
  }
  if (ca) {
    auto ka = cat::cat_assign_kernel;
    CARCA_LAUNCH(ka, dim3(ceil_div(B, 64)), dim3(32), 0, st, ucol, counters + 1, f.useg, B);
    TRY(check_launch("cat_assign"));
    auto kf = cat::cat_fill_ca_kernel;
    const long long rows_cap = (long long)B * L;
    CARCA_LAUNCH(kf, dim3((unsigned)ceil_div_ll(rows_cap * 32, 256)), dim3(256), 0, st, f);
    TRY(check_launch("cat_fill_ca"));
  } else {
    f.n_groups_out = counters + 1;
    auto kf = cat::cat_fill_dot_kernel;
    CARCA_LAUNCH(kf, dim3((unsigned)ceil_div_ll((long long)B * 32, 256)), dim3(256), 0, st, f);
    TRY(check_launch("cat_fill_dot"));
  }
  cat::CatArgs a;
  memset(&a, 0, sizeof(a));
  a.TA = ca ? reinterpret_cast<const float*>(plan + pl.tq) : Tf;
  a.tw = reinterpret_cast<const float*>(plan + pl.tw);
  a.Bt = Bt; a.meta = meta; a.n_groups = counters + 1;
  a.counts = counts;
  a.dbf = m->cross.bf;
  a.item_lo = item_lo; a.n_items = n_shard; a.residual_ca = m->residual_ca; a.decoder = m->decoder_kind;
  a.status = status;
  const int n_tiles = ceil_div(n_shard, cat::CT);
  a.parts = max(1, min(16, ceil_div(148 * 6, n_tiles)));
  if (const char* e = getenv("CARCA_CAT_PARTS")) a.parts = max(1, atoi(e));   // (development: work split per item tile)
  const size_t smem = sizeof(cat::CatSmem);
  if (const char* e = getenv("CARCA_CAT_DBG")) a.dbg = reinterpret_cast<volatile int*>(strtoull(e, nullptr, 10));
  if (ca) {
    auto k = cat::catalog_tc_kernel<1>;
    static bool set1 = false;
    if (!set1) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set1 = true; }
    CARCA_LAUNCH(k, dim3(n_tiles * a.parts), dim3(cat::CAT_THREADS), smem, st, a);
  } else {
    auto k = cat::catalog_tc_kernel<0>;
    static bool set0 = false;
    if (!set0) { cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); set0 = true; }
    CARCA_LAUNCH(k, dim3(n_tiles * a.parts), dim3(cat::CAT_THREADS), smem, st, a);
  }
  return check_launch("catalog_tc");
}

}  // namespace

extern "C" {

int64_t carca_rows_plan_bytes(const carca_model_params* m) { return rows_plan(m).total; }

int carca_rows_prepare(void* plan_v, const float* plan_f32, const carca_model_params* m, void* stream) {
  cudaStream_t st = S(stream);
  CARCA_REQUIRE(rows_shape_ok_f32(m), "rows_prepare: needs d in {32, 64, 128, 256}, head width 16 / 32 / 64, C <= 8, "
                                      "1..8 blocks (got d=%d H=%d C=%d blocks=%d)", m->embed.d, m->n_heads, m->embed.n_ctx,
                m->n_blocks);
  unsigned char* plan = reinterpret_cast<unsigned char*>(plan_v);
  const RowsPlan pl = rows_plan(m);
  const PlanLayout fl = plan_layout(m);
  const int d = m->embed.d;
  const long long n = m->embed.n_items;
  const float* T = plan_f32 + fl.tfold;
  const float* Mc = plan_f32 + fl.mc;
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
  // fp32 flavour: K and V share their A operand (src/carca.py:239-240), so [Wk; Wv] runs as ONE [R, d] x [d, 2d] GEMM
  for (int b = 0; b <= m->n_blocks; ++b) {
    if (b == m->n_blocks && m->decoder_kind != 1) break;
    const float* wk = b < m->n_blocks ? m->blocks[b].wk : m->cross.wk;
    const float* wv = b < m->n_blocks ? m->blocks[b].wv : m->cross.wv;
    const float* bk = b < m->n_blocks ? m->blocks[b].bk : m->cross.bk;
    const float* bv = b < m->n_blocks ? m->blocks[b].bv : m->cross.bv;
    float* w2 = reinterpret_cast<float*>(plan + pl.wkv) + (long long)b * 2 * d * d;
    float* b2 = reinterpret_cast<float*>(plan + pl.bkv) + (long long)b * 2 * d;
    cudaMemcpyAsync(w2, wk, sizeof(float) * d * d, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(w2 + (long long)d * d, wv, sizeof(float) * d * d, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(b2, bk, sizeof(float) * d, cudaMemcpyDeviceToDevice, st);
    cudaMemcpyAsync(b2 + d, bv, sizeof(float) * d, cudaMemcpyDeviceToDevice, st);
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
      GemmArgs g = gemm_defaults(T, m->cross.wq, reinterpret_cast<float*>(plan + pl.tq), (int)n, d, d);
      g.bias = m->cross.bq;
      TRY(launch_gemm(g, st));
      auto cvt = rows::to_bf16_kernel;
      CARCA_LAUNCH(cvt, dim3((unsigned)ceil_div_ll(n * d, 1024)), dim3(256), 0, st, reinterpret_cast<bf16*>(plan + pl.tqb),
                   reinterpret_cast<const float*>(plan + pl.tq), n * d);
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

int carca_rows_set_stage_events(void* const* events, int n) {
  CARCA_REQUIRE(n >= 0 && n <= 64 && (n == 0 || events != nullptr), "rows_set_stage_events: 0..64 events");
  for (int i = 0; i < n; ++i) g_stage_ev[i] = reinterpret_cast<cudaEvent_t>(events[i]);
  g_stage_n = n;
  g_stage_used = 0;
  return 0;
}

int carca_rows_stage_ids(int32_t* ids, int cap) {
  const int n = g_stage_used < cap ? g_stage_used : cap;
  for (int i = 0; i < n; ++i) ids[i] = g_stage_id[i];
  return n;
}

int64_t carca_rows_catalog_scratch_bytes(const carca_model_params* m, int B, int L) { return cat_scratch(m, B, L).total; }

int carca_rows_catalog_counts(int32_t* counts, const void* plan, const float* plan_f32, const carca_model_params* m,
                              const int32_t* p_x, const float* p_c, const float* ctx_user, const int32_t* pos_item,
                              int item_lo, int n_shard, int B, int L, int32_t* status, void* scratch, void* stream) {
  const int d = m->embed.d, H = m->n_heads;
  CARCA_REQUIRE(d == 64 && rows_shape_ok_f32(m) && L >= 1 && L <= cat::CT,
                "rows_catalog_counts: needs d = 64, L <= 128, C <= 8 (got d=%d L=%d C=%d)", d, L, m->embed.n_ctx);
  CARCA_REQUIRE(m->decoder_kind == 0 || H == 2, "rows_catalog_counts: the cross-attention decoder needs 2 heads (got %d)", H);
  CARCA_REQUIRE(status != nullptr && scratch != nullptr, "rows_catalog_counts: status and scratch are required");
  if (m->embed.pos) CARCA_REQUIRE(L <= m->embed.pos_len, "rows_catalog_counts: sequence length %d > positional table %d", L,
                                  m->embed.pos_len);
  if (B <= 0 || n_shard <= 0) return 0;
  const unsigned char* pl = reinterpret_cast<const unsigned char*>(plan);
  unsigned char* scr = reinterpret_cast<unsigned char*>(scratch);
  const float* Tf = plan_f32 + plan_layout(m).tfold;
  cudaStream_t st = S(stream);
  if (H == 1) return catalog_counts_t<1>(counts, pl, Tf, m, p_x, p_c, ctx_user, pos_item, item_lo, n_shard, B, L, status, scr, st);
  if (H == 2) return catalog_counts_t<2>(counts, pl, Tf, m, p_x, p_c, ctx_user, pos_item, item_lo, n_shard, B, L, status, scr, st);
  if (H == 4) return catalog_counts_t<4>(counts, pl, Tf, m, p_x, p_c, ctx_user, pos_item, item_lo, n_shard, B, L, status, scr, st);
  return fail(-4, "rows_catalog_counts: unsupported head count %d", H);
}

int carca_rows_eval_forward(float* y, int64_t ldy, int col0, const void* plan, const float* plan_f32,
                            const carca_model_params* m, const int32_t* p_x, const float* p_c, const int32_t* o_x,
                            const float* o_c, int B, int L, int T, int ctx_per_user, int cat_lo, int precision,
                            int32_t* status, void* scratch, void* stream) {
  CARCA_REQUIRE(precision == 0 || precision == 1, "rows_eval_forward: precision %d (0 = bf16, 1 = fp32)", precision);
  if (precision == 0)
    CARCA_REQUIRE(rows_shape_ok(m), "rows_eval_forward(bf16): needs d in {64, 256}, head width 32 or 64, C <= 8, 1..8 "
                                    "blocks (got d=%d H=%d C=%d blocks=%d)", m->embed.d, m->n_heads, m->embed.n_ctx,
                  m->n_blocks);
  else
    CARCA_REQUIRE(rows_shape_ok_f32(m), "rows_eval_forward(fp32): needs d in {32, 64, 128, 256}, head width 16 / 32 / 64, "
                                        "C <= 8, 1..8 blocks (got d=%d H=%d C=%d blocks=%d)", m->embed.d, m->n_heads,
                  m->embed.n_ctx, m->n_blocks);
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
  const float* Tf = plan_f32 + plan_layout(m).tfold;
  if (precision == 0) {
#define ROWS_CASE(DD, HH)                                                                                            \
  if (d == DD && H == HH)                                                                                            \
    return forward_t<DD, HH>(y, ldy, col0, pl, Tf, m, p_x, p_c, o_x, o_c, B, L, T, ctx_per_user, cat_lo, status, scr, st);
    ROWS_CASE(64, 1) ROWS_CASE(64, 2) ROWS_CASE(256, 4) ROWS_CASE(256, 8)
#undef ROWS_CASE
  } else {
#define ROWS_CASE(DD, HH)                                                                                            \
  if (d == DD && H == HH)                                                                                            \
    return forward_f32_t<DD, HH>(y, ldy, col0, pl, Tf, m, p_x, p_c, o_x, o_c, B, L, T, ctx_per_user, cat_lo, scr, st);
    ROWS_CASE(32, 1) ROWS_CASE(32, 2) ROWS_CASE(64, 1) ROWS_CASE(64, 2) ROWS_CASE(64, 4) ROWS_CASE(128, 2)
    ROWS_CASE(128, 4) ROWS_CASE(128, 8) ROWS_CASE(256, 4) ROWS_CASE(256, 8)
#undef ROWS_CASE
  }
  return fail(-4, "rows_eval_forward: unsupported (d, heads) = (%d, %d)", d, H);
}

}  // extern "C"
#endif
