// extern "C" entry points of libcarca_b200.so (declared in include/carca_b200.h).
// Each composes the kernels of this directory on the caller's stream; nothing here allocates,
// synchronises or falls back to the host.
#include "../../include/carca_b200.h"

#include <cmath>

#include "attention.cuh"
#include "batch.cuh"
#include "common.cuh"
#include "embed.cuh"
#include "fused_eval.cuh"
#include "fused_eval_tc.cuh"
#include "fused_train.cuh"
#include "gemm.cuh"
#include "layernorm.cuh"
#include "optim.cuh"
#include "plan_layout.cuh"
#include "score.cuh"
#include "variants.cuh"

using namespace carca;

#define TRY(expr)            \
  do {                       \
    int _rc = (expr);        \
    if (_rc != 0) return _rc; \
  } while (0)

namespace carca {
void set_seed_source(const unsigned long long* p);
}

namespace {

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
// dropout configuration of a kernel launch: host seed (+ the optional device seed word)
inline DropCfg drop_cfg(float p, unsigned long long seed, unsigned site) {
  DropCfg c = make_drop(p, seed, site);
  c.seed_dev = seed_source();
  return c;
}
inline int warp_rows_grid(long long rows) { return (int)ceil_div_ll(rows, 8); }  // 8 warps / 256-thread CTA

// y = x w^T + b with the optional fused epilogue pieces
int linear(float* y, const float* x, const float* w, const float* b, int M, int N, int K, long long ldw,
           cudaStream_t st, int act = 0, const DropCfg* drop = nullptr, const float* R = nullptr,
           int accumulate = 0, const float* row_mask = nullptr, const int* a_rows = nullptr, long long lda = -1,
           float alpha = 1.0f, int r_mod = 0) {
  GemmArgs g = gemm_defaults(x, w, y, M, N, K);
  g.ldb = ldw;
  g.lda = lda < 0 ? K : lda;
  g.bias = b;
  g.act = act;
  if (drop) g.drop = *drop;
  g.R = R;
  g.ldr = N;
  g.r_mod = r_mod;
  g.accumulate = accumulate;
  g.row_mask = row_mask;
  g.a_rows = a_rows;
  g.alpha = alpha;
  return launch_gemm(g, st);
}

// dx[M,K] (=|+=) alpha * dy[M,N] w[N,K] (+ R)
int linear_dx(float* dx, const float* dy, const float* w, int M, int N, int K, long long ldw, cudaStream_t st,
              int accumulate = 0, float alpha = 1.0f, const float* R = nullptr) {
  GemmArgs g = gemm_defaults(dy, w, dx, M, K, N);
  g.lda = N;
  g.transB = 0;
  g.ldb = ldw;
  g.ldc = K;
  g.accumulate = accumulate;
  g.alpha = alpha;
  g.R = R;
  g.ldr = K;
  return launch_gemm(g, st);
}

// dw[N,K] (ld lddw) += alpha * dy[P,N]^T x[rows(P),K]
int linear_dw(float* dw, const float* dy, const float* x, int P, int N, int K, long long lddw, long long ldx,
              cudaStream_t st, const int* x_rows = nullptr, float alpha = 1.0f) {
  GemmArgs g = gemm_defaults(dy, x, dw, N, K, P);
  g.transA = 1;
  g.lda = N;
  g.transB = 0;
  g.ldb = ldx;
  g.b_rows = x_rows;
  g.ldc = lddw;
  g.accumulate = 1;
  g.alpha = alpha;
  return launch_gemm(g, st);
}

int colsum(float* out, const float* X, int M, int N, long long ldx, cudaStream_t st) {
  if (M <= 0 || N <= 0) return 0;
  const int rows_per_block = 256;
  dim3 grid(ceil_div(N, 32), ceil_div(M, rows_per_block)), block(32, 8);
  auto k = colsum_kernel;
  CARCA_LAUNCH(k, grid, block, 0, st, out, X, M, N, ldx, rows_per_block);
  return check_launch("colsum");
}

int transpose(float* dst, const float* src, int R, int Cc, long long ld_src, long long ld_dst, int accumulate,
              cudaStream_t st) {
  if (R <= 0 || Cc <= 0) return 0;
  dim3 grid(ceil_div(Cc, 32), ceil_div(R, 32)), block(32, 8);
  auto k = transpose_kernel;
  CARCA_LAUNCH(k, grid, block, 0, st, dst, src, R, Cc, ld_src, ld_dst, accumulate);
  return check_launch("transpose");
}

int scale_rows(float* Y, const float* X, const float* rs, DropCfg drop, long long rows, int N, cudaStream_t st) {
  const long long total = rows * N;
  if (total <= 0) return 0;
  auto k = scale_rows_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, st, Y, X, rs, drop, total, N);
  return check_launch("scale_rows");
}

int layernorm_fwd(float* y, float* mean, float* rstd, const float* x, const float* g, const float* b, int rows,
                  int d, cudaStream_t st) {
  if (rows <= 0) return 0;
  auto k = layernorm_fwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid(rows)), dim3(256), 0, st, y, mean, rstd, x, g, b, rows, d);
  return check_launch("layernorm_fwd");
}

int layernorm_bwd(float* dx, float* dg, float* db, const float* dy, const float* x, const float* mean,
                  const float* rstd, const float* g, int rows, int d, int acc, cudaStream_t st) {
  if (rows <= 0) return 0;
  const int warps = 8;
  const int grid = min(warp_rows_grid(rows), 148 * 4);
  const size_t smem = sizeof(float) * 2 * warps * (size_t)d;
  if (smem > 48 * 1024) return fail(-2, "layernorm_bwd: d=%d too large for the shared partials", d);
  auto k = layernorm_bwd_kernel;
  CARCA_LAUNCH(k, dim3(grid), dim3(256), smem, st, dx, dg, db, dy, x, mean, rstd, g, rows, d, acc);
  return check_launch("layernorm_bwd");
}

#ifndef CARCA_EMU
template <class K>
int allow_smem(K kern, size_t bytes) {
  if (bytes <= 48 * 1024) return 0;
  if (bytes > 227 * 1024) return fail(-2, "attention: %zu bytes of shared memory exceed 227 KB", bytes);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes);
  if (e != cudaSuccess) return fail(-3, "cudaFuncSetAttribute: %s", cudaGetErrorString(e));
  return 0;
}
#else
template <class K>
int allow_smem(K, size_t) { return 0; }
#endif

AttnArgs attn_args(const float* Q, const float* K, const float* V, const float* qm, const float* km, int B, int H,
                   int Lq, int Lk, int d, int causal_on, int diag, float p, uint64_t seed, uint32_t site) {
  AttnArgs a;
  memset(&a, 0, sizeof(a));
  a.Q = Q; a.K = K; a.V = V;
  a.q_mask = qm; a.k_mask = km;
  a.B = B; a.H = H; a.Lq = Lq; a.Lk = Lk; a.dh = d / H;
  a.ldq = d; a.ldk = d; a.ldo = d;
  a.causal = causal_on; a.diag = diag;
  a.sqrt_dh = (float)std::sqrt((double)(d / H));
  a.drop = drop_cfg(p, seed, site);
  return a;
}

int attention_fwd(AttnArgs a, cudaStream_t st) {
  if (a.B <= 0 || a.Lq <= 0) return 0;
  if (a.Lk > kAttnMaxChunks * 32) return fail(-2, "attention: Lk=%d > %d unsupported", a.Lk, kAttnMaxChunks * 32);
  const size_t smem = attention_fwd_smem(a.Lk, a.dh);
  auto k = attention_fwd_kernel;
  TRY(allow_smem(k, smem));
  // enough CTAs to fill the machine: split the query rows when B*H is small
  int chunks = 1;
  while ((long long)a.B * a.H * chunks < 2 * 148 && ceil_div(a.Lq, chunks) > kAttnWarps) chunks *= 2;
  a.rows_per_cta = ceil_div(a.Lq, chunks);
  dim3 grid(a.B * a.H, ceil_div(a.Lq, a.rows_per_cta));
  CARCA_LAUNCH(k, grid, dim3(kAttnWarps * 32), smem, st, a);
  return check_launch("attention_fwd");
}

int attention_bwd(AttnArgs a, cudaStream_t st) {
  if (a.B <= 0 || a.Lq <= 0) return 0;
  if (a.Lk > kAttnMaxChunks * 32) return fail(-2, "attention: Lk=%d > %d unsupported", a.Lk, kAttnMaxChunks * 32);
  const size_t smem = attention_bwd_smem(a.Lq, a.Lk, a.dh);
  auto k = attention_bwd_kernel;
  TRY(allow_smem(k, smem));
  CARCA_LAUNCH(k, dim3(a.B * a.H), dim3(kAttnWarps * 32), smem, st, a);
  return check_launch("attention_bwd");
}

__global__ void __launch_bounds__(256) lrelu_drop_bwd_kernel(float* __restrict__ g, const float* __restrict__ a1,
                                                             DropCfg drop, long long n) {
  // g <- g * dropfactor * LeakyReLU'(pre-activation); a1 = dropout(LeakyReLU(.)) keeps its sign
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float f = drop_factor(drop, (unsigned long long)i);
  g[i] = g[i] * f * (a1[i] > 0.f ? 1.0f : kLeakySlope);
}

}  // namespace

extern "C" {

const char* carca_last_error(void) { return err_buf(); }
int carca_abi_version(void) { return CARCA_B200_ABI_VERSION; }
int64_t carca_launch_count(void) { return (int64_t)launch_count(); }
void carca_set_seed_source(const uint64_t* device_seed) {
  set_seed_source(reinterpret_cast<const unsigned long long*>(device_seed));
}

int carca_transpose(float* dst, const float* src, int rows, int cols, int accumulate, void* stream) {
  return transpose(dst, src, rows, cols, cols, rows, accumulate, S(stream));
}

int carca_padding_mask(float* mask, const int32_t* ids, int64_t n, void* stream) {
  if (n <= 0) return 0;
  auto k = padding_mask_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, S(stream), mask, ids, (long long)n);
  return check_launch("padding_mask");
}

int carca_padding_mask_f32(float* mask, const float* x, int64_t n, void* stream) {
  if (n <= 0) return 0;
  auto k = padding_mask_f32_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, S(stream), mask, x, (long long)n);
  return check_launch("padding_mask_f32");
}

int carca_dropout(float* y, const float* x, int64_t n, float p, uint64_t seed, uint32_t site, void* stream) {
  CARCA_REQUIRE(p >= 0.f && p < 1.f, "dropout: p=%f outside [0,1)", p);
  return scale_rows(y, x, nullptr, drop_cfg(p, seed, site), n, 1, S(stream));
}

// ------------------------------------------------------------------------------------ embedding
// q = Wf [a | c] + bf for P positions (the first linear of every attribute embedding, src/carca.py:86,113,138)
static int feats_forward(float* q_out, const carca_embed_params* w, const carca_attr_source* at, const int32_t* x,
                         const float* ctx, const float* mask, int P, cudaStream_t st) {
  const int g = w->g, A = w->n_attrs, C = w->n_ctx;
  if (at->kind == CARCA_ATTR_CSR) {
    CARCA_REQUIRE(w->feats_wT != nullptr, "embed_fwd: CSR attributes need feats_wT");
    auto k = feat_csr_fwd_kernel;
    CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, st, q_out, x, at->csr_rowptr, at->csr_cols,
                 at->csr_vals, ctx, w->feats_wT, w->feats_b, mask, P, g, A, C);
    TRY(check_launch("feat_csr_fwd"));
  } else if (at->kind == CARCA_ATTR_DENSE && w->feats_wT != nullptr && A >= 1024 && g <= 256) {
    // dense rows of a large vocabulary (the reference API's multi-hot tensors): scan + gather-sum, bound by the read
    auto k = feat_dense_scan_fwd_kernel;
    CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, st, q_out, at->dense, ctx, w->feats_wT, w->feats_b, mask, P, g, A,
                 C);
    TRY(check_launch("feat_dense_scan_fwd"));
  } else if (at->kind == CARCA_ATTR_TABLE || at->kind == CARCA_ATTR_DENSE) {
    const int* rows = at->kind == CARCA_ATTR_TABLE ? x : nullptr;
    TRY(linear(q_out, at->dense, w->feats_w, w->feats_b, P, g, A, A + C, st, 0, nullptr, nullptr, 0, nullptr, rows,
               A));
    if (C > 0) TRY(linear(q_out, ctx, w->feats_w + A, nullptr, P, g, C, A + C, st, 0, nullptr, nullptr, 1));
  } else {
    return fail(-2, "embed_fwd: unknown attribute source kind %d", at->kind);
  }
  return 0;
}

// d feats_w, d feats_b from dq [P, g]
static int feats_backward(float* d_feats_w, float* d_feats_b, const float* dq, const carca_embed_params* w,
                          const carca_attr_source* at, const int32_t* x, const float* ctx, const float* mask, int P,
                          float* scratch_wT, cudaStream_t st) {
  const int g = w->g, A = w->n_attrs, C = w->n_ctx;
  TRY(colsum(d_feats_b, dq, P, g, g, st));
  if (C > 0) TRY(linear_dw(d_feats_w + A, dq, ctx, P, g, C, A + C, C, st));           // dWf[:, A:]
  if (at->kind == CARCA_ATTR_CSR) {
    CARCA_REQUIRE(scratch_wT != nullptr, "embed_bwd: CSR attributes need scratch_wT");
    auto k = feat_csr_bwd_kernel;
    CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, st, scratch_wT, dq, x, at->csr_rowptr, at->csr_cols,
                 at->csr_vals, mask, P, g);
    TRY(check_launch("feat_csr_bwd"));
    TRY(transpose(d_feats_w, scratch_wT, A, g, g, A + C, 1, st));                     // dWf[:, :A] += dWT^T
  } else if (at->kind == CARCA_ATTR_TABLE || at->kind == CARCA_ATTR_DENSE) {
    const int* rows = at->kind == CARCA_ATTR_TABLE ? x : nullptr;
    TRY(linear_dw(d_feats_w, dq, at->dense, P, g, A, A + C, A, st, rows));
  } else {
    return fail(-2, "embed_bwd: unknown attribute source kind %d", at->kind);
  }
  return 0;
}

int carca_feats_fwd(float* q, const carca_embed_params* w, const carca_attr_source* at, const int32_t* x,
                    const float* ctx, int P, void* stream) {
  if (P <= 0) return 0;
  return feats_forward(q, w, at, x, ctx, nullptr, P, S(stream));
}

int carca_feats_bwd(float* d_feats_w, float* d_feats_b, const float* dq, const carca_embed_params* w,
                    const carca_attr_source* at, const int32_t* x, const float* ctx, int P, float* scratch_wT,
                    void* stream) {
  if (P <= 0) return 0;
  return feats_backward(d_feats_w, d_feats_b, dq, w, at, x, ctx, nullptr, P, scratch_wT, S(stream));
}

int carca_embed_fwd(float* e, float* q_out, const carca_embed_params* w, const carca_attr_source* at,
                    const int32_t* x, const float* ctx, const float* mask, int n_rows, int n_cols, int is_target,
                    void* stream) {
  cudaStream_t st = S(stream);
  const int P = n_rows * n_cols;
  if (P <= 0) return 0;
  const int d = w->d, g = w->g;
  CARCA_REQUIRE(q_out != nullptr, "embed_fwd: q_out is required");
  TRY(feats_forward(q_out, w, at, x, ctx, mask, P, st));
  const float sqrt_d = (float)std::sqrt((double)d);
  // e = sqrt(d) * E[x] Wj[:, :d]^T
  TRY(linear(e, w->items_embed, w->joint_w, nullptr, P, d, d, d + g, st, 0, nullptr, nullptr, 0, nullptr, x, d,
             sqrt_d));
  // e = mask * (e + q Wj[:, d:]^T + bj (+ pos))
  const float* pos = (!is_target && w->pos) ? w->pos : nullptr;
  if (pos) CARCA_REQUIRE(n_cols <= w->pos_len, "embed_fwd: sequence length %d > positional table %d", n_cols,
                         w->pos_len);
  TRY(linear(e, q_out, w->joint_w + d, w->joint_b, P, d, g, d + g, st, 0, nullptr, pos, 1, mask, nullptr, g, 1.0f,
             pos ? n_cols : 0));
  return 0;
}

int carca_embed_bwd(const carca_embed_grads* gr, const float* de, const float* q_saved, const carca_embed_params* w,
                    const carca_attr_source* at, const int32_t* x, const float* ctx, const float* mask, int n_rows,
                    int n_cols, int is_target, float* scratch_pd, float* scratch_pg, float* scratch_wT,
                    void* stream) {
  cudaStream_t st = S(stream);
  const int P = n_rows * n_cols;
  if (P <= 0) return 0;
  const int d = w->d, g = w->g;
  const float sqrt_d = (float)std::sqrt((double)d);
  float* dem = scratch_pd;
  float* dz = scratch_pd + (long long)P * d;
  float* dq = scratch_pg;
  const DropCfg nodrop = make_drop(0.f, 0ull, 0u);
  TRY(scale_rows(dem, de, mask, nodrop, P, d, st));                               // de * mask  (:94)
  if (!is_target && w->pos && gr->pos) TRY(colsum(gr->pos, dem, n_rows, n_cols * d, (long long)n_cols * d, st));
  TRY(colsum(gr->joint_b, dem, P, d, d, st));
  TRY(linear_dw(gr->joint_w, dem, w->items_embed, P, d, d, d + g, d, st, x, sqrt_d));   // dWj[:, :d]
  TRY(linear_dw(gr->joint_w + d, dem, q_saved, P, d, g, d + g, g, st));                 // dWj[:, d:]
  TRY(linear_dx(dq, dem, w->joint_w + d, P, d, g, d + g, st));                          // dq [P,g]
  TRY(linear_dx(dz, dem, w->joint_w, P, d, d, d + g, st, 0, sqrt_d));                   // dz [P,d]
  {
    auto k = scatter_add_rows_kernel;
    CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, st, gr->items_embed, dz, x, P, d);
    TRY(check_launch("scatter_add_rows"));
  }
  return feats_backward(gr->feats_w, gr->feats_b, dq, w, at, x, ctx, mask, P, scratch_wT, st);
}

// ------------------------------------------------------------------------------------ module variants (N3)
int carca_gather_rows_fwd(float* out, const float* table, const int32_t* ids, float alpha, int P, int d,
                          void* stream) {
  if (P <= 0) return 0;
  auto k = gather_rows_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, S(stream), out, table, ids, alpha, P, d);
  return check_launch("gather_rows");
}

int carca_gather_rows_bwd(float* d_table, const float* d_out, const int32_t* ids, float alpha, int P, int d,
                          void* stream) {
  if (P <= 0) return 0;
  auto k = scatter_add_rows_scaled_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid(P)), dim3(256), 0, S(stream), d_table, d_out, ids, alpha, P, d);
  return check_launch("scatter_add_rows_scaled");
}

int carca_pos_mask_fwd(float* out, const float* in, const float* pos, const float* mask, int n_rows, int n_cols,
                       int d, void* stream) {
  const long long total = (long long)n_rows * n_cols * d;
  if (total <= 0) return 0;
  auto k = pos_mask_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(total, 256)), dim3(256), 0, S(stream), out, in, pos, mask, total, d,
               n_cols);
  return check_launch("pos_mask");
}

int carca_embed_folded_fwd(float* out, const float* T, const float* McT, const int32_t* x, const float* c,
                           const float* pos, const float* mask, int n_rows, int n_cols, int d, int n_ctx, void* stream) {
  CARCA_REQUIRE(d % 4 == 0, "embed_folded_fwd: d = %d must be a multiple of 4", d);
  const long long total4 = (long long)n_rows * n_cols * (d / 4);
  if (total4 <= 0) return 0;
  auto k = embed_folded_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(total4, 256)), dim3(256), 0, S(stream), out, T, McT, x, c, pos, mask, total4,
               d, n_ctx, n_cols);
  return check_launch("embed_folded");
}

int carca_pos_mask_bwd(float* d_in, float* d_pos, const float* d_out, const float* mask, int n_rows, int n_cols,
                       int d, void* stream) {
  const long long P = (long long)n_rows * n_cols;
  if (P <= 0) return 0;
  TRY(scale_rows(d_in, d_out, mask, make_drop(0.f, 0ull, 0u), P, d, S(stream)));
  if (d_pos) TRY(colsum(d_pos, d_in, n_rows, n_cols * d, (long long)n_cols * d, S(stream)));
  return 0;
}

int carca_wdot_score_fwd(float* y, const float* p, const float* o, int B, int T, int Lp, int d, int per_position,
                         float gamma, int normalize, int64_t ldy, int col0, void* stream) {
  if ((long long)B * T <= 0) return 0;
  if (per_position) CARCA_REQUIRE(T == Lp, "wdot_score: training mode needs T == L (%d vs %d)", T, Lp);
  auto k = wdot_score_fwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, S(stream), y, p, o, B, T, Lp, d, per_position,
               gamma, normalize, (long long)ldy, col0);
  return check_launch("wdot_score_fwd");
}

int carca_wdot_score_bwd(float* d_o, float* d_p, const float* dy, const float* y, const float* p, const float* o,
                         int B, int T, int Lp, int d, int per_position, float gamma, int normalize, int64_t ldy,
                         int col0, void* stream) {
  if ((long long)B * T <= 0) return 0;
  auto k = wdot_score_bwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, S(stream), d_o, d_p, dy, y, p, o, B, T, Lp, d,
               per_position, gamma, normalize, (long long)ldy, col0);
  return check_launch("wdot_score_bwd");
}

int carca_knn_score(float* y, const float* p_a, const float* o_a, int B, int T, int Lp, int A, int64_t ldy, int col0,
                    void* stream) {
  if ((long long)B * T <= 0) return 0;
  auto k = knn_score_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, S(stream), y, p_a, o_a, B, T, Lp, A,
               (long long)ldy, col0);
  return check_launch("knn_score");
}

// ------------------------------------------------------------------------------------ layernorm / linear
int carca_layernorm_fwd(float* y, float* mean, float* rstd, const float* x, const float* gamma, const float* beta,
                        int rows, int d, void* stream) {
  return layernorm_fwd(y, mean, rstd, x, gamma, beta, rows, d, S(stream));
}

int carca_layernorm_bwd(float* dx, float* dgamma, float* dbeta, const float* dy, const float* x, const float* mean,
                        const float* rstd, const float* gamma, int rows, int d, int accumulate_dx, void* stream) {
  return layernorm_bwd(dx, dgamma, dbeta, dy, x, mean, rstd, gamma, rows, d, accumulate_dx, S(stream));
}

int carca_linear_fwd(float* y, const float* x, const float* w, const float* bias, int M, int N, int K, int act_leaky,
                     void* stream) {
  return linear(y, x, w, bias, M, N, K, K, S(stream), act_leaky ? 1 : 0);
}

int carca_linear_bwd_input(float* dx, const float* dy, const float* w, int M, int N, int K, int accumulate,
                           void* stream) {
  return linear_dx(dx, dy, w, M, N, K, K, S(stream), accumulate);
}

int carca_linear_bwd_weight(float* dw, float* db, const float* dy, const float* x, int M, int N, int K,
                            void* stream) {
  TRY(linear_dw(dw, dy, x, M, N, K, K, K, S(stream)));
  if (db) TRY(colsum(db, dy, M, N, N, S(stream)));
  return 0;
}

// ------------------------------------------------------------------------------------ attention
int carca_attention_fwd(float* O, float* W_out, const float* Q, const float* K, const float* V, const float* q_mask,
                        const float* k_mask, int B, int H, int Lq, int Lk, int d, int causal_on, int diag,
                        float p_drop, uint64_t seed, uint32_t site, void* stream) {
  CARCA_REQUIRE(H > 0 && d % H == 0, "Embedding dim must be divisible by number of heads");  // src/carca.py:208
  AttnArgs a = attn_args(Q, K, V, q_mask, k_mask, B, H, Lq, Lk, d, causal_on, diag, p_drop, seed, site);
  a.O = O;
  a.W = W_out;
  return attention_fwd(a, S(stream));
}

int carca_attention_bwd(float* dQ, float* dK, float* dV, const float* dO, const float* Q, const float* K,
                        const float* V, const float* q_mask, const float* k_mask, int B, int H, int Lq, int Lk,
                        int d, int causal_on, int diag, float p_drop, uint64_t seed, uint32_t site, void* stream) {
  CARCA_REQUIRE(H > 0 && d % H == 0, "Embedding dim must be divisible by number of heads");
  AttnArgs a = attn_args(Q, K, V, q_mask, k_mask, B, H, Lq, Lk, d, causal_on, diag, p_drop, seed, site);
  a.dO = dO; a.dQ = dQ; a.dK = dK; a.dV = dV;
  return attention_bwd(a, S(stream));
}

// ------------------------------------------------------------------------------------ SelfAttentionBlock
int carca_sa_block_fwd(float* out, const carca_block_saved* sv, const float* x, const float* mask,
                       const carca_block_params* w, int B, int L, int d, int H, int residual, float p_drop,
                       uint64_t seed, int block_index, void* stream) {
  cudaStream_t st = S(stream);
  CARCA_REQUIRE(H > 0 && d % H == 0, "Embedding dim must be divisible by number of heads");
  const int P = B * L;
  if (P <= 0) return 0;
  const uint32_t s_attn = 1 + 3 * block_index, s_f1 = 2 + 3 * block_index, s_f2 = 3 + 3 * block_index;
  TRY(layernorm_fwd(sv->qn, sv->mean1, sv->rstd1, x, w->ln1_g, w->ln1_b, P, d, st));            // :298
  TRY(linear(sv->Q, sv->qn, w->wq, w->bq, P, d, d, d, st));                                      // :238
  TRY(linear(sv->K, x, w->wk, w->bk, P, d, d, d, st));                                           // :239 (raw x)
  TRY(linear(sv->V, x, w->wv, w->bv, P, d, d, d, st));                                           // :240
  AttnArgs a = attn_args(sv->Q, sv->K, sv->V, mask, mask, B, H, L, L, d, 1, 0, p_drop, seed, s_attn);  // :299
  a.O = sv->s;
  a.resid = residual ? sv->qn : nullptr;                                                         // :302
  TRY(attention_fwd(a, st));
  TRY(layernorm_fwd(sv->s2, sv->mean2, sv->rstd2, sv->s, w->ln2_g, w->ln2_b, P, d, st));         // :304
  const DropCfg d1 = drop_cfg(p_drop, seed, s_f1), d2 = drop_cfg(p_drop, seed, s_f2);
  TRY(linear(sv->a1, sv->s2, w->w1, w->b1, P, d, d, d, st, 1, &d1));                             // :307-309
  TRY(linear(out, sv->a1, w->w2, w->b2, P, d, d, d, st, 0, &d2, residual ? sv->s2 : nullptr));   // :311-316
  return 0;
}

int carca_sa_block_bwd(float* dx, const carca_block_grads* gr, const float* dout, const float* x, const float* mask,
                       const carca_block_params* w, const carca_block_saved* sv, int B, int L, int d, int H,
                       int residual, float p_drop, uint64_t seed, int block_index, float* scratch4, void* stream) {
  cudaStream_t st = S(stream);
  const int P = B * L;
  if (P <= 0) return 0;
  const long long n = (long long)P * d;
  float* t0 = scratch4;
  float* t1 = t0 + n;
  float* t2 = t1 + n;
  float* t3 = t2 + n;
  const uint32_t s_attn = 1 + 3 * block_index, s_f1 = 2 + 3 * block_index, s_f2 = 3 + 3 * block_index;
  const DropCfg d1 = drop_cfg(p_drop, seed, s_f1), d2 = drop_cfg(p_drop, seed, s_f2);
  // ---- FFN
  const float* df2 = dout;
  if (p_drop > 0.f) {
    TRY(scale_rows(t0, dout, nullptr, d2, n, 1, st));
    df2 = t0;
  }
  TRY(linear_dw(gr->w2, df2, sv->a1, P, d, d, d, d, st));
  TRY(colsum(gr->b2, df2, P, d, d, st));
  TRY(linear_dx(t1, df2, w->w2, P, d, d, d, st));                          // d a1
  {
    auto k = lrelu_drop_bwd_kernel;
    CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, st, t1, sv->a1, d1, n);
    TRY(check_launch("lrelu_drop_bwd"));                                   // t1 = d f1
  }
  TRY(linear_dw(gr->w1, t1, sv->s2, P, d, d, d, d, st));
  TRY(colsum(gr->b1, t1, P, d, d, st));
  TRY(linear_dx(t2, t1, w->w1, P, d, d, d, st, 0, 1.0f, residual ? dout : nullptr));   // d s2
  // ---- LN2
  TRY(layernorm_bwd(t0, gr->ln2_g, gr->ln2_b, t2, sv->s, sv->mean2, sv->rstd2, w->ln2_g, P, d, 0, st));  // t0 = d s
  // ---- attention
  AttnArgs a = attn_args(sv->Q, sv->K, sv->V, mask, mask, B, H, L, L, d, 1, 0, p_drop, seed, s_attn);
  a.dO = t0; a.dQ = t1; a.dK = t2; a.dV = t3;
  TRY(attention_bwd(a, st));
  TRY(linear_dw(gr->wq, t1, sv->qn, P, d, d, d, d, st));
  TRY(colsum(gr->bq, t1, P, d, d, st));
  TRY(linear_dw(gr->wk, t2, x, P, d, d, d, d, st));
  TRY(colsum(gr->bk, t2, P, d, d, st));
  TRY(linear_dw(gr->wv, t3, x, P, d, d, d, d, st));
  TRY(colsum(gr->bv, t3, P, d, d, st));
  // d qn = dQ WQ (+ d s through the residual) -> in place over t0
  TRY(linear_dx(t0, t1, w->wq, P, d, d, d, st, residual ? 1 : 0));
  TRY(linear_dx(dx, t2, w->wk, P, d, d, d, st, 0));
  TRY(linear_dx(dx, t3, w->wv, P, d, d, d, st, 1));
  // ---- LN1
  TRY(layernorm_bwd(dx, gr->ln1_g, gr->ln1_b, t0, x, sv->mean1, sv->rstd1, w->ln1_g, P, d, 1, st));
  return 0;
}

// ------------------------------------------------------------------------------------ decoders
int carca_dot_score_fwd(float* y, const float* p, const float* o, int B, int T, int Lp, int d, int per_position,
                        int64_t ldy, int col0, void* stream) {
  if ((long long)B * T <= 0) return 0;
  if (per_position) CARCA_REQUIRE(T == Lp, "dot_score: training mode needs T == L (%d vs %d)", T, Lp);
  auto k = dot_score_fwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, S(stream), y, p, o, B, T, Lp, d,
               per_position, (long long)ldy, col0);
  return check_launch("dot_score_fwd");
}

int carca_dot_score_bwd(float* d_o, float* d_p, const float* dy, const float* y, const float* p, const float* o,
                        int B, int T, int Lp, int d, int per_position, int64_t ldy, int col0, void* stream) {
  if ((long long)B * T <= 0) return 0;
  auto k = dot_score_bwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, S(stream), d_o, d_p, dy, y, p, o, B, T, Lp,
               d, per_position, (long long)ldy, col0);
  return check_launch("dot_score_bwd");
}

int carca_cross_score_fwd(float* y, const carca_cross_saved* sv, const float* o, const float* o_mask, const float* p,
                          const float* p_mask, const carca_cross_params* w, int B, int T, int Lp, int d, int H,
                          int residual, int training, float p_drop, uint64_t seed, uint32_t site, int64_t ldy,
                          int col0, void* stream) {
  cudaStream_t st = S(stream);
  CARCA_REQUIRE(H > 0 && d % H == 0, "Embedding dim must be divisible by number of heads");
  if ((long long)B * T <= 0) return 0;
  TRY(linear(sv->Q, o, w->wq, w->bq, B * T, d, d, d, st));
  TRY(linear(sv->K, p, w->wk, w->bk, B * Lp, d, d, d, st));
  TRY(linear(sv->V, p, w->wv, w->bv, B * Lp, d, d, d, st));
  AttnArgs a = attn_args(sv->Q, sv->K, sv->V, o_mask, p_mask, B, H, T, Lp, d, training ? 1 : 0, -1, p_drop, seed,
                         site);                                                          // :339-340
  a.O = sv->s;
  a.resid = residual ? o : nullptr;                                                      // :343
  TRY(attention_fwd(a, st));
  auto k = rowdot_sigmoid_fwd_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid((long long)B * T)), dim3(256), 0, st, y, sv->s, w->wf, w->bf, B, T, d,
               (long long)ldy, col0);                                                    // :345-347
  return check_launch("rowdot_sigmoid_fwd");
}

int carca_cross_score_bwd(float* d_o, float* d_p, const carca_cross_grads* gr, const float* dy, const float* y,
                          const carca_cross_saved* sv, const float* o, const float* o_mask, const float* p,
                          const float* p_mask, const carca_cross_params* w, int B, int T, int Lp, int d, int H,
                          int residual, int training, float p_drop, uint64_t seed, uint32_t site, int64_t ldy,
                          int col0, float* scratch4, void* stream) {
  cudaStream_t st = S(stream);
  if ((long long)B * T <= 0) return 0;
  const long long n = (long long)max(B * T, B * Lp) * d;
  float* ds = scratch4;
  float* dQ = ds + n;
  float* dK = dQ + n;
  float* dV = dK + n;
  {
    const int warps = 8;
    const int grid = min(warp_rows_grid((long long)B * T), 148 * 4);
    const size_t smem = sizeof(float) * ((size_t)warps * d + warps);
    CARCA_REQUIRE(smem <= 48 * 1024, "cross_score_bwd: d=%d too large", d);
    auto k = rowdot_sigmoid_bwd_kernel;
    CARCA_LAUNCH(k, dim3(grid), dim3(256), smem, st, ds, gr->wf, gr->bf, dy, y, sv->s, w->wf, B, T, d,
                 (long long)ldy, col0);
    TRY(check_launch("rowdot_sigmoid_bwd"));
  }
  AttnArgs a = attn_args(sv->Q, sv->K, sv->V, o_mask, p_mask, B, H, T, Lp, d, training ? 1 : 0, -1, p_drop, seed,
                         site);
  a.dO = ds; a.dQ = dQ; a.dK = dK; a.dV = dV;
  TRY(attention_bwd(a, st));
  TRY(linear_dw(gr->wq, dQ, o, B * T, d, d, d, d, st));
  TRY(colsum(gr->bq, dQ, B * T, d, d, st));
  TRY(linear_dw(gr->wk, dK, p, B * Lp, d, d, d, d, st));
  TRY(colsum(gr->bk, dK, B * Lp, d, d, st));
  TRY(linear_dw(gr->wv, dV, p, B * Lp, d, d, d, d, st));
  TRY(colsum(gr->bv, dV, B * Lp, d, d, st));
  TRY(linear_dx(d_o, dQ, w->wq, B * T, d, d, d, st, 0, 1.0f, residual ? ds : nullptr));
  TRY(linear_dx(d_p, dK, w->wk, B * Lp, d, d, d, st, 1));
  TRY(linear_dx(d_p, dV, w->wv, B * Lp, d, d, d, st, 1));
  return 0;
}

// ------------------------------------------------------------------------------------ fused training core
namespace {
long long* g_train_ticks = nullptr;
int g_train_ticks_cap = 0;
int train_args(TrainArgs& a, const carca_train_core* c) {
  CARCA_REQUIRE(c != nullptr, "train_core: null descriptor");
  if (c->L > TMAXL || c->L < 1) return fail(-4, "train_core: L=%d outside [1, %d]", c->L, TMAXL);
  if (!(c->n_heads == 1 || c->n_heads == 2 || c->n_heads == 4))
    return fail(-4, "train_core: n_heads=%d not in {1, 2, 4}", c->n_heads);
  if (c->n_blocks < 0 || c->n_blocks > TMAXB) return fail(-4, "train_core: n_blocks=%d > %d", c->n_blocks, TMAXB);
  if (c->n_tuples < 1 || c->n_tuples > TMAXT) return fail(-4, "train_core: %d target tuples (1 or 2)", c->n_tuples);
  CARCA_REQUIRE(c->decoder_kind == 0 || c->decoder_kind == 1, "train_core: unknown decoder kind %d", c->decoder_kind);
  CARCA_REQUIRE(c->p_drop >= 0.f && c->p_drop < 1.f, "train_core: p=%f outside [0,1)", c->p_drop);
  CARCA_REQUIRE(c->rows && c->saved && c->p_x, "train_core: missing workspace or inputs");
  memset(&a, 0, sizeof(a));
  a.B = c->B; a.L = c->L; a.H = c->n_heads; a.n_blocks = c->n_blocks; a.n_tuples = c->n_tuples;
  a.decoder = c->decoder_kind; a.residual_sa = c->residual_sa; a.residual_ca = c->residual_ca;
  a.drop = drop_cfg(c->p_drop, c->seed, 0u);
  a.n_sms = 148;
  a.ticks = g_train_ticks;
  a.ticks_cap = g_train_ticks_cap;
  a.p_x = c->p_x; a.p_e = c->p_e;
  for (int t = 0; t < c->n_tuples; ++t) {
    CARCA_REQUIRE(c->o_x[t] != nullptr, "train_core: target tuple %d missing", t);
    a.o_x[t] = c->o_x[t];
    a.o_e[t] = c->o_e[t];
  }
  a.n_bins = c->rows;
  a.row_src = c->rows + 4;
  a.row_info = a.row_src + (long long)c->B * TR;
  a.sv = c->saved;
  a.sv_stride = (long long)c->B * TR * TD;
  static_assert(sizeof(TrainBlockW) == sizeof(carca_block_params), "block parameter structs must match");
  static_assert(sizeof(TrainCrossW) == sizeof(carca_cross_params), "decoder parameter structs must match");
  for (int b = 0; b < c->n_blocks; ++b) memcpy(&a.blk[b], &c->blocks[b], sizeof(TrainBlockW));
  a.fn_g = c->norm_g; a.fn_b = c->norm_b;
  memcpy(&a.dec, &c->cross, sizeof(TrainCrossW));
  if (c->embed) {
    const carca_embed_params* w = c->embed;
    if (w->d != TD) return fail(-4, "train_core: d=%d, the fused kernels handle d=64", w->d);
    if (w->n_ctx > TMAXC) return fail(-4, "train_core: %d context features (at most %d)", w->n_ctx, TMAXC);
    CARCA_REQUIRE(c->attrs && c->attrs->kind == CARCA_ATTR_CSR, "train_core: the folded embedding needs CSR attributes");
    CARCA_REQUIRE(c->fold && c->p_c, "train_core: folded embedding without workspace / context");
    if (w->pos) CARCA_REQUIRE(c->L <= w->pos_len, "train_core: sequence length %d > positional table %d", c->L, w->pos_len);
    a.embed_mode = 1;
    a.C = w->n_ctx; a.A = w->n_attrs; a.ldj = w->d + w->g;
    a.sqrt_d = (float)std::sqrt((double)w->d);
    a.E = w->items_embed; a.Wj = w->joint_w; a.fold = c->fold; a.pos_table = w->pos;
    a.csr_rowptr = c->attrs->csr_rowptr; a.csr_cols = c->attrs->csr_cols; a.csr_vals = c->attrs->csr_vals;
    a.p_c = c->p_c;
    for (int t = 0; t < c->n_tuples; ++t) {
      CARCA_REQUIRE(c->o_c[t] != nullptr, "train_core: context of target tuple %d missing", t);
      a.o_c[t] = c->o_c[t];
    }
  } else {
    CARCA_REQUIRE(c->p_e != nullptr, "train_core: p_e missing");
    for (int t = 0; t < c->n_tuples; ++t) CARCA_REQUIRE(c->o_e[t] != nullptr, "train_core: o_e of tuple %d missing", t);
  }
  return 0;
}
int train_grid(int B) { return B < 148 ? (B < 1 ? 1 : B) : 148; }
}  // namespace

void carca_train_core_set_ticks(int64_t* device_buf, int capacity) {
  g_train_ticks = reinterpret_cast<long long*>(device_buf);
  g_train_ticks_cap = device_buf ? capacity : 0;
}

int64_t carca_train_core_rows_ints(int B) { return 4 + 2 * (int64_t)B * TR; }
int64_t carca_train_core_saved_floats(int B, int n_blocks, int n_tuples) {
  return (int64_t)sv_count(n_blocks, n_tuples) * B * TR * TD;
}

int64_t carca_train_core_fold_floats(const carca_embed_params* w) {
  return w ? ((int64_t)w->n_attrs + w->n_ctx + 1) * TD : 0;
}

int carca_train_core_fwd(float* y, int64_t ldy, const carca_train_core* c, void* stream) {
  cudaStream_t st = S(stream);
  TrainArgs a;
  TRY(train_args(a, c));
  if (a.B <= 0) return 0;
  a.y = y;
  a.ldy = ldy;
  if (a.embed_mode) {
    // GT[a][c] = sum_g Wf[g][a] Wj[c][d + g]   ([A + C, 64]);  cst = Wj[:, d:] bf + bj
    const carca_embed_params* w = c->embed;
    const int AC = w->n_attrs + w->n_ctx;
    GemmArgs g = gemm_defaults(w->feats_w, w->joint_w + w->d, c->fold, AC, TD, w->g);
    g.transA = 1; g.lda = AC;
    g.transB = 1; g.ldb = w->d + w->g;
    g.ldc = TD;
    TRY(launch_gemm(g, st));
    auto kc = fold_cst_kernel;
    CARCA_LAUNCH(kc, dim3(8), dim3(256), 0, st, c->fold + (long long)AC * TD, w->joint_w, w->feats_b, w->joint_b, w->d,
                 w->g);
    TRY(check_launch("fold_cst"));
  }
  cudaMemsetAsync(a.n_bins, 0, 4 * sizeof(int), st);
  cudaMemsetAsync(a.row_src, 0xFF, sizeof(int) * (size_t)a.B * TR, st);
  {
    auto k = train_pack_kernel;
    CARCA_LAUNCH(k, dim3(ceil_div(a.B, 128)), dim3(1024), 0, st, a);
    TRY(check_launch("train_pack"));
  }
  auto k = fused_train_fwd_kernel;
  const size_t smem = fused_train_smem();
  TRY(allow_smem(k, smem));
  CARCA_LAUNCH(k, dim3(train_grid(a.B)), dim3(TTHREADS), smem, st, a);
  return check_launch("fused_train_fwd");
}

int carca_train_core_bwd(float* d_pe, float* d_oe0, float* d_oe1, const carca_block_grads* g_blocks, float* g_norm_g,
                         float* g_norm_b, const carca_cross_grads* g_cross, const carca_embed_grads* g_embed,
                         float* d_fold, const float* dy, int64_t ldy, const carca_train_core* c, void* stream) {
  cudaStream_t st = S(stream);
  TrainArgs a;
  TRY(train_args(a, c));
  if (a.B <= 0) return 0;
  CARCA_REQUIRE(dy != nullptr, "train_core_bwd: dy missing");
  if (a.embed_mode) {
    CARCA_REQUIRE(g_embed && d_fold && g_embed->items_embed && g_embed->feats_w && g_embed->feats_b &&
                      g_embed->joint_w && g_embed->joint_b,
                  "train_core_bwd: embedding gradients / workspace missing");
    a.gE = g_embed->items_embed; a.gWj = g_embed->joint_w; a.dfold = d_fold;
    a.gpos = a.pos_table ? g_embed->pos : nullptr;
    cudaMemsetAsync(d_fold, 0, sizeof(float) * (size_t)carca_train_core_fold_floats(c->embed), st);
  } else {
    CARCA_REQUIRE(d_pe && d_oe0 && (c->n_tuples < 2 || d_oe1), "train_core_bwd: missing gradient buffers");
  }
  CARCA_REQUIRE(g_norm_g && g_norm_b && (c->n_blocks == 0 || g_blocks), "train_core_bwd: missing parameter gradients");
  a.dy = dy;
  a.ldy = ldy;
  a.d_pe = d_pe;
  a.d_oe[0] = d_oe0;
  a.d_oe[1] = d_oe1;
  for (int b = 0; b < c->n_blocks; ++b) memcpy(&a.gblk[b], &g_blocks[b], sizeof(TrainBlockG));
  a.g_fn_g = g_norm_g; a.g_fn_b = g_norm_b;
  if (c->decoder_kind == 1) {
    CARCA_REQUIRE(g_cross != nullptr, "train_core_bwd: decoder gradients missing");
    memcpy(&a.gdec, g_cross, sizeof(TrainCrossG));
  }
  auto k = fused_train_bwd_kernel;
  const size_t smem = fused_train_smem();
  TRY(allow_smem(k, smem));
  CARCA_LAUNCH(k, dim3(train_grid(a.B)), dim3(TTHREADS), smem, st, a);
  TRY(check_launch("fused_train_bwd"));
  if (a.embed_mode) {
    // un-fold: dG = d_fold rows [A + C, 64] (transposed d(Wj_q Wf)), dcst = its last row
    const carca_embed_params* w = c->embed;
    const int AC = w->n_attrs + w->n_ctx, d = w->d, gd = w->g;
    const float* dcst = d_fold + (long long)AC * TD;
    {  // d Wf[g][a] = sum_c Wj[c][d + g] dG[a][c]
      GemmArgs g = gemm_defaults(w->joint_w + d, d_fold, g_embed->feats_w, gd, AC, TD);
      g.transA = 1; g.lda = d + gd;
      g.transB = 1; g.ldb = TD;
      g.ldc = AC;
      TRY(launch_gemm(g, st));
    }
    {  // d Wj[c][d + g] += sum_a dG[a][c] Wf[g][a]
      GemmArgs g = gemm_defaults(d_fold, w->feats_w, g_embed->joint_w + d, TD, gd, AC);
      g.transA = 1; g.lda = TD;
      g.transB = 1; g.ldb = AC;
      g.ldc = d + gd;
      g.accumulate = 1;
      TRY(launch_gemm(g, st));
    }
    auto ku = unfold_cst_kernel;   // d Wj[:, d:] += dcst (x) bf;  d bf = Wj[:, d:]^T dcst;  d bj = dcst
    CARCA_LAUNCH(ku, dim3(d), dim3(256), 0, st, g_embed->joint_w, g_embed->feats_b, g_embed->joint_b, dcst, w->joint_w,
                 w->feats_b, d, gd);
    TRY(check_launch("unfold_cst"));
  }
  return 0;
}

// ------------------------------------------------------------------------------------ loss / metrics
int carca_bce_sums(float* sums, const float* y_pred, const int32_t* y_true, const float* mask, int64_t n, float eps,
                   void* stream) {
  if (n <= 0) return 0;
  const int grid = (int)min((long long)148 * 8, ceil_div_ll(n, 256));
  auto k = bce_sums_kernel;
  CARCA_LAUNCH(k, dim3(grid), dim3(256), 0, S(stream), sums, y_pred, y_true, mask, (long long)n, eps);
  return check_launch("bce_sums");
}

int carca_bce_finalize(float* loss, const float* sums, void* stream) {
  auto k = bce_finalize_kernel;
  CARCA_LAUNCH(k, dim3(1), dim3(32), 0, S(stream), loss, sums);
  return check_launch("bce_finalize");
}

int carca_bce_bwd(float* dy, const float* grad_out, const float* sums, const float* y_pred, const int32_t* y_true,
                  const float* mask, int64_t n, float eps, void* stream) {
  if (n <= 0) return 0;
  auto k = bce_bwd_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll(n, 256)), dim3(256), 0, S(stream), dy, grad_out, sums, y_pred, y_true,
               mask, (long long)n, eps);
  return check_launch("bce_bwd");
}

int carca_rank_metrics(double* acc, int32_t* first_rank, const float* y_pred, const int32_t* y_true, int B, int T,
                       int64_t ldy, int64_t ldt, int k, void* stream) {
  if (B <= 0) return 0;
  const int grid = min(148 * 4, ceil_div(B, 8));
  auto kern = rank_metrics_kernel;
  CARCA_LAUNCH(kern, dim3(grid), dim3(256), 0, S(stream), acc, first_rank, y_pred, y_true, B, T, (long long)ldy,
               (long long)ldt, k);
  return check_launch("rank_metrics");
}

int carca_eval_metrics(double* stats, double* work, const float* y_pred, const int32_t* y_true, const int32_t* o_x, int B,
                       int T, int64_t ldy, int64_t ldt, int64_t ldx, int k, float eps, void* stream) {
  if (B <= 0 || T <= 0) return 0;
  const int grid = min(148 * 4, ceil_div(B, 8));
  auto kern = eval_metrics_kernel;
  CARCA_LAUNCH(kern, dim3(grid), dim3(256), 0, S(stream), stats, work, y_pred, y_true, o_x, B, T, (long long)ldy,
               (long long)ldt, (long long)ldx, k, eps);
  return check_launch("eval_metrics");
}

// ------------------------------------------------------------------------------------ optimizer
int carca_adam_step(const carca_adam_tensor* tensors, int n_tensors, float lr, float beta1, float beta2, float eps,
                    float weight_decay, void* stream) {
  cudaStream_t st = S(stream);
  CARCA_REQUIRE(n_tensors >= 0 && (n_tensors == 0 || tensors), "adam_step: missing arguments");
  for (int first = 0; first < n_tensors; first += kAdamMaxTensors) {
    AdamArgs a;
    memset(&a, 0, sizeof(a));
    a.n_tensors = min(kAdamMaxTensors, n_tensors - first);
    int chunks = 0;
    for (int i = 0; i < a.n_tensors; ++i) {
      const carca_adam_tensor& t = tensors[first + i];
      CARCA_REQUIRE(t.param && t.grad && t.exp_avg && t.exp_avg_sq && t.step && t.numel >= 0,
                    "adam_step: tensor %d incomplete", first + i);
      a.p[i] = t.param; a.g[i] = t.grad; a.m[i] = t.exp_avg; a.v[i] = t.exp_avg_sq; a.step[i] = t.step;
      a.n[i] = t.numel;
      a.chunk0[i] = chunks;
      chunks += (int)ceil_div_ll(t.numel, kAdamChunk);
    }
    a.chunk0[a.n_tensors] = chunks;
    a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.weight_decay = weight_decay;
    {
      auto k = adam_tick_kernel;
      CARCA_LAUNCH(k, dim3(1), dim3(kAdamMaxTensors), 0, st, a);
      TRY(check_launch("adam_tick"));
    }
    if (chunks == 0) continue;
    auto k = adam_step_kernel;
    CARCA_LAUNCH(k, dim3(chunks), dim3(256), 0, st, a);
    TRY(check_launch("adam_step"));
  }
  return 0;
}

// ------------------------------------------------------------------------------------ batch construction
namespace {
int check_log(const carca_interactions* lg, int L, int n_neg) {
  CARCA_REQUIRE(lg && lg->rowptr && lg->items && lg->ctx, "build_batch: incomplete interaction log");
  CARCA_REQUIRE(L >= 1 && n_neg <= kMaxNeg, "build_batch: at most %d sampled negatives per user (got %d)", kMaxNeg, n_neg);
  return 0;
}
InteractionLog as_log(const carca_interactions* lg) {
  InteractionLog l;
  l.rowptr = lg->rowptr; l.items = lg->items; l.ctx = lg->ctx; l.n_users = lg->n_users; l.n_ctx = lg->n_ctx;
  return l;
}
}  // namespace

int carca_build_eval_batch(int32_t* p_x, float* p_c, int32_t* o_x, float* o_c_user, int32_t* y_true,
                           const carca_interactions* log, const int32_t* users, int B, int L, int T, int n_items,
                           int mode, int test, uint64_t seed, void* stream) {
  if (B <= 0) return 0;
  CARCA_REQUIRE(mode == 1 || mode == 2, "build_eval_batch: mode must be 1 (val) or 2 (test)");
  CARCA_REQUIRE(T >= 1, "build_eval_batch: T counts the positive, so T >= 1");
  TRY(check_log(log, L, T - 1));
  auto k = build_eval_batch_kernel;
  CARCA_LAUNCH(k, dim3(ceil_div(B, 4)), dim3(128), 0, S(stream), p_x, p_c, o_x, o_c_user, y_true, as_log(log), users, B,
               L, T, n_items, mode, test, (unsigned long long)seed);
  return check_launch("build_eval_batch");
}

int carca_build_train_batch(int32_t* p_x, float* p_c, int32_t* o_x, float* o_c, int32_t* y_true,
                            const carca_interactions* log, const int32_t* users, int B, int L, int n_items, int test,
                            uint64_t seed, void* stream) {
  if (B <= 0) return 0;
  TRY(check_log(log, L, L));
  auto k = build_train_batch_kernel;
  CARCA_LAUNCH(k, dim3(ceil_div(B, 4)), dim3(128), 0, S(stream), p_x, p_c, o_x, o_c, y_true, as_log(log), users, B, L,
               n_items, test, (unsigned long long)seed);
  return check_launch("build_train_batch");
}

int carca_unpack_windows(int32_t* p_x, float* p_c, const int32_t* offs, const float* rows, int B, int L, int C,
                         void* stream) {
  if (B <= 0 || L <= 0) return 0;
  CARCA_REQUIRE(C >= 0 && C <= 64, "unpack_windows: C=%d outside 0..64", C);
  auto k = unpack_windows_kernel;
  CARCA_LAUNCH(k, dim3((unsigned)ceil_div_ll((long long)B * L, 256)), dim3(256), 0, S(stream), p_x, p_c, offs, rows, B, L, C);
  return check_launch("unpack_windows");
}

// ------------------------------------------------------------------------------------ fused inference
int64_t carca_eval_plan_floats(const carca_model_params* m) { return plan_layout(m).total; }

int carca_eval_prepare(float* plan, float* scratch_q, const carca_model_params* m, const carca_attr_source* at,
                       void* stream) {
  cudaStream_t st = S(stream);
  const carca_embed_params* w = &m->embed;
  const int n = w->n_items, d = w->d, g = w->g, A = w->n_attrs, C = w->n_ctx;
  CARCA_REQUIRE(C <= 8, "eval_prepare: at most 8 context features (got %d)", C);
  const PlanLayout pl = plan_layout(m);
  float* T = plan + pl.tfold;
  // q[i] = Wf[:, :A] attrs[i] + bf   (no context: it is folded into Mc)
  if (at->kind == CARCA_ATTR_CSR) {
    CARCA_REQUIRE(w->feats_wT != nullptr, "eval_prepare: CSR attributes need feats_wT");
    auto k = feat_csr_fwd_kernel;
    CARCA_LAUNCH(k, dim3(warp_rows_grid(n)), dim3(256), 0, st, scratch_q, (const int*)nullptr, at->csr_rowptr,
                 at->csr_cols, at->csr_vals, (const float*)nullptr, w->feats_wT, w->feats_b, (const float*)nullptr, n,
                 g, A, 0);
    TRY(check_launch("feat_csr_fwd(fold)"));
  } else if (at->kind == CARCA_ATTR_TABLE) {
    TRY(linear(scratch_q, at->dense, w->feats_w, w->feats_b, n, g, A, A + C, st, 0, nullptr, nullptr, 0, nullptr,
               nullptr, A));
  } else {
    return fail(-2, "eval_prepare: needs a device-resident attribute table (CSR or TABLE)");
  }
  const float sqrt_d = (float)std::sqrt((double)d);
  TRY(linear(T, w->items_embed, w->joint_w, nullptr, n, d, d, d + g, st, 0, nullptr, nullptr, 0, nullptr, nullptr, d,
             sqrt_d));
  TRY(linear(T, scratch_q, w->joint_w + d, w->joint_b, n, d, g, d + g, st, 0, nullptr, nullptr, 1, nullptr, nullptr,
             g));
  // Mc[o][c] = sum_j Wj[o][d + j] Wf[j][A + c], stored with row stride 8
  cudaMemsetAsync(plan + pl.mc, 0, sizeof(float) * d * 8, st);
  if (C > 0) {
    GemmArgs gm = gemm_defaults(w->joint_w + d, w->feats_w + A, plan + pl.mc, d, C, g);
    gm.lda = d + g;
    gm.transB = 0;
    gm.ldb = A + C;
    gm.ldc = 8;
    TRY(launch_gemm(gm, st));
  }
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    float* dst = plan + pl.blocks + (long long)b * 5 * d * d;
    const float* src[5] = {bp.wq, bp.wk, bp.wv, bp.w1, bp.w2};
    for (int i = 0; i < 5; ++i) TRY(transpose(dst + (long long)i * d * d, src[i], d, d, d, d, 0, st));
  }
  if (m->decoder_kind == 1) {
    float* dst = plan + pl.cross;
    const float* src[3] = {m->cross.wq, m->cross.wk, m->cross.wv};
    for (int i = 0; i < 3; ++i) TRY(transpose(dst + (long long)i * d * d, src[i], d, d, d, d, 0, st));
  }
#ifndef CARCA_EMU
  if (d == 64) {   // K-major tf32 operands (hi | lo) with the bias folded in as an extra K step
    auto pk = pack_weight_tc_kernel;
    for (int b = 0; b < m->n_blocks; ++b) {
      const carca_block_params& bp = m->blocks[b];
      const float* ws[5] = {bp.wq, bp.wk, bp.wv, bp.w1, bp.w2};
      const float* bs[5] = {bp.bq, bp.bk, bp.bv, bp.b1, bp.b2};
      for (int i = 0; i < 5; ++i) {
        CARCA_LAUNCH(pk, dim3(5), dim3(256), 0, st, plan + pl.tc_blocks + ((long long)b * 5 + i) * kTcPacked, ws[i],
                     bs[i]);
        TRY(check_launch("pack_weight_tc"));
      }
    }
    if (m->decoder_kind == 1) {
      const float* ws[3] = {m->cross.wq, m->cross.wk, m->cross.wv};
      const float* bs[3] = {m->cross.bq, m->cross.bk, m->cross.bv};
      for (int i = 0; i < 3; ++i) {
        CARCA_LAUNCH(pk, dim3(5), dim3(256), 0, st, plan + pl.tc_cross + (long long)i * kTcPacked, ws[i], bs[i]);
        TRY(check_launch("pack_weight_tc"));
      }
      // candidate-side folds of the decoder (exact re-associations of linear maps, like the item table):
      //   TQ[i] = WQ T[i] + bq  (the query of candidate i before its context term),  tw[i] = <T[i], wf>,
      //   McQ = WQ Mc,  mcw = wf Mc,  so Q = TQ[x] + McQ c  and  <o, wf> = tw[x] + <mcw, c>.
      TRY(linear(plan + pl.tq, T, m->cross.wq, m->cross.bq, n, 64, 64, 64, st));
      TRY(linear(plan + pl.tw, T, m->cross.wf, nullptr, n, 1, 64, 64, st));
      {
        GemmArgs gm = gemm_defaults(m->cross.wq, plan + pl.mc, plan + pl.mcq, 64, 8, 64);
        gm.transB = 0;
        gm.ldb = 8;
        gm.ldc = 8;
        TRY(launch_gemm(gm, st));
        GemmArgs gw = gemm_defaults(m->cross.wf, plan + pl.mc, plan + pl.mcw, 1, 8, 64);
        gw.transB = 0;
        gw.ldb = 8;
        gw.ldc = 8;
        TRY(launch_gemm(gw, st));
      }
    }
  }
#endif
  return 0;
}

// row_src | row_seg | n_bins (packing pass), then — split decoder — Kg [rows, 64] | Ug [rows, 4] | useg [B] | cwsg [B]
static int64_t eval_scratch_pack_bytes(int B) {   // row_src | row_seg | n_bins [4] | bin_users [B + 1] | order [B / 2 + 2]
  return ((int64_t)sizeof(int) * (2ll * (B + 1) * 64 + 4 + (B + 1) + (B / 2 + 2)) + 255) / 256 * 256;
}
#ifndef CARCA_EMU
static int eval_forward_tc(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                           const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                           int T, int32_t* status, float* dbg, int dbg_stage, int cat_lo, int ctx_per_user, void* scratch,
                           void* stream, int dec) {
  const PlanLayout pl = plan_layout(m);
  TcArgs a;
  memset(&a, 0, sizeof(a));
  a.Tfold = plan + pl.tfold;
  a.Mc = plan + pl.mc;
  a.pos = m->embed.pos;
  a.p_x = p_x; a.p_c = p_c; a.o_x = o_x; a.o_c = o_c;
  a.y = y; a.ldy = ldy; a.col0 = col0;
  a.B = B; a.L = L; a.T = T; a.C = m->embed.n_ctx; a.H = m->n_heads;
  a.n_blocks = m->n_blocks;
  a.residual_sa = m->residual_sa; a.residual_ca = m->residual_ca; a.decoder = m->decoder_kind;
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    const float* wt = plan + pl.tc_blocks + (long long)b * 5 * kTcPacked;
    TcBlockW& f = a.blk[b];
    f.ln1_g = bp.ln1_g; f.ln1_b = bp.ln1_b; f.ln2_g = bp.ln2_g; f.ln2_b = bp.ln2_b;
    f.wq = wt; f.wk = wt + kTcPacked; f.wv = wt + 2 * kTcPacked; f.w1 = wt + 3 * kTcPacked; f.w2 = wt + 4 * kTcPacked;
  }
  a.fn_g = m->norm_g; a.fn_b = m->norm_b;
  if (m->decoder_kind == 1) {
    const float* wt = plan + pl.tc_cross;
    a.dwk = wt + kTcPacked; a.dwv = wt + 2 * kTcPacked;
    a.dwf = m->cross.wf; a.dbf = m->cross.bf;
    a.TQ = plan + pl.tq; a.tw = plan + pl.tw; a.McQ = plan + pl.mcq; a.mcw = plan + pl.mcw;
  }
  a.status = status;
  a.dbg = dbg;
  a.dbg_stage = dbg_stage;
  a.cat_lo = cat_lo;
  a.oc_user = ctx_per_user ? m->embed.n_ctx : (long long)T * m->embed.n_ctx;
  a.oc_tgt = ctx_per_user ? 0 : m->embed.n_ctx;
  // pack the valid profile rows into 64-row bins; scratch: row_src | row_seg (B + 1 bins of 64 rows each: a
  // tile is two bins, so an odd bin count reads one empty bin past the last) | n_bins
  const long long rows = ((long long)B + 1) * 64;
  int* row_src = reinterpret_cast<int*>(scratch);
  int* row_seg = row_src + rows;
  int* n_bins = row_seg + rows;
  cudaMemsetAsync(n_bins, 0, 2 * sizeof(int), S(stream));   // [0] bin count, [1] tile scheduler counter
  {
    auto pk = pack_rows_kernel;
    int* bin_users = n_bins + 4;
    int* order = bin_users + (B + 1);
    CARCA_LAUNCH(pk, dim3(ceil_div(B, 128)), dim3(128), 0, S(stream), row_src, row_seg, n_bins, status, p_x, B, L,
                 bin_users);
    TRY(check_launch("pack_rows"));
    a.order = nullptr;
    if (B > 2 * 148) {   // more tiles than CTAs can be: hand out the longest tiles first
      auto ok = tile_order_kernel;
      CARCA_LAUNCH(ok, dim3(1), dim3(256), 0, S(stream), order, bin_users, n_bins);
      TRY(check_launch("tile_order"));
      a.order = order;
    }
  }
  a.row_src = row_src; a.row_seg = row_seg; a.n_bins = n_bins;
  a.chunk_slices = max(1, ceil_div(ceil_div(T, 128), 16));   // <= 16 candidate chunks per work item
  const size_t smem = sizeof(TcSmem);
  // dec < 0: chosen here by mode and, for two-head cross-attention, on the device by profile density (TcArgs::auto_dec)
  const bool autosel = dec < 0;
  if (autosel) dec = cat_lo > 0 ? 3 : 1;
  a.auto_dec = (autosel && m->n_heads == 2 && m->decoder_kind == 1) ? 1 : 0;
  TcArgs a_dense = a;          // the tcgen05-decoder launch of the same call
  a_dense.auto_dec = 2;
  // split decoder: two heads, cross-attention, one context row per user, packed rows addressable in 24 bits
  if (dec == 3 && !(m->n_heads == 2 && m->decoder_kind == 1 && ctx_per_user && rows < (1ll << 24))) dec = cat_lo > 0 ? 2 : 1;
  if (dec == 3) {
    float* f = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(scratch) + eval_scratch_pack_bytes(B));
    a.Kg = f;
    a.Ug = f + rows * 64;
    a.useg = reinterpret_cast<int*>(f + rows * 68);
    a.cwsg = f + rows * 68 + B;
    a.chunk_slices = 1;   // the tile is only encoded here
  }
  // upper bound on the work items (a small batch's tiles are split into up to 4 items each: tail slicing in the kernel)
  const long long n_tiles = (long long)ceil_div(B, 2) * max(a.chunk_slices, 4);
  if (m->n_heads == 2 && dec == 3) {
    auto k = fused_eval_tc_kernel<2, 3>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles, 148ll)), dim3(TC_THREADS), smem, S(stream), a);
    TRY(check_launch("fused_eval_tc (encoder)"));
    DecPairsArgs d;
    d.Kg = a.Kg; d.Ug = a.Ug; d.cwsg = a.cwsg; d.TQ = a.TQ; d.tw = a.tw; d.dbf = a.dbf;
    d.useg = a.useg; d.o_x = o_x;
    d.y = y; d.ldy = ldy; d.col0 = col0; d.B = B; d.T = T; d.cat_lo = cat_lo; d.residual_ca = m->residual_ca;
    d.sc = 1.4426950408889634f / sqrtf(32.0f);
    d.n_bins = n_bins; d.auto_dec = a.auto_dec;
    const long long items = (long long)B * ((T + 1) / 2);
    auto dk = decode_pairs_kernel;
    CARCA_LAUNCH(dk, dim3((unsigned)min((items + 255) / 256, 148ll * 64)), dim3(256), 0, S(stream), d);
    TRY(check_launch("decode_pairs"));
  } else if (m->n_heads == 2 && dec == 2) {
    auto k = fused_eval_tc_kernel<2, 2>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles, 148ll)), dim3(TC_THREADS), smem, S(stream), a);
  } else if (m->n_heads == 2 && dec == 1) {
    auto k = fused_eval_tc_kernel<2, 1>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles, 148ll)), dim3(TC_THREADS), smem, S(stream), a);
  } else if (m->n_heads == 2) {
    auto k = fused_eval_tc_kernel<2, 0>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles, 148ll)), dim3(TC_THREADS), smem, S(stream), a);
  } else {
    auto k = fused_eval_tc_kernel<4, 0>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles, 148ll)), dim3(TC_THREADS), smem, S(stream), a);
  }
  TRY(check_launch("fused_eval_tc"));
  if (a.auto_dec) {   // dense profiles: the same forward with the tcgen05 decoder (returns at once otherwise)
    const long long n_tiles_d = (long long)ceil_div(B, 2) * max(a_dense.chunk_slices, 4);
    auto k = fused_eval_tc_kernel<2, 0>;
    TRY(allow_smem(k, smem));
    CARCA_LAUNCH(k, dim3((unsigned)min(n_tiles_d, 148ll)), dim3(TC_THREADS), smem, S(stream), a_dense);
    TRY(check_launch("fused_eval_tc (dense)"));
  }
  return 0;
}
#endif

static int eval_forward_any(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                            const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                            int T, int variant, int32_t* status, float* dbg, int dbg_stage, int cat_lo, int ctx_per_user,
                            void* scratch, void* stream) {
  const int d = m->embed.d, H = m->n_heads;
  const bool common_ok = d == FD && L >= 1 && m->embed.n_ctx <= 8 && m->n_blocks <= FMAXB && H >= 1 && FD % H == 0;
  const bool ffma_ok = common_ok && L <= FLP && (FD / H) % 4 == 0;
#ifndef CARCA_EMU
  const bool tc_ok = common_ok && L <= 256 && (H == 2 || H == 4) && status != nullptr && scratch != nullptr;
#else
  const bool tc_ok = false;
#endif
  if (variant >= 2 && variant <= 6 && !tc_ok)
    return fail(-4, "eval_forward: tensor-core kernel needs d=64, L<=256, H in {2,4}, status and scratch");
  if (variant == 1 && !ffma_ok) return fail(-4, "eval_forward: FFMA kernel needs d=64, L<=52, dh%%4==0");
  if (!ffma_ok && !tc_ok)
    return fail(-4, "eval_forward: fused kernels support d=64, L<=256, C<=8, <=8 blocks (got d=%d L=%d C=%d blocks=%d "
                    "H=%d)", d, L, m->embed.n_ctx, m->n_blocks, H);
  if (m->embed.pos) CARCA_REQUIRE(L <= m->embed.pos_len, "eval_forward: sequence length %d > positional table %d", L,
                                  m->embed.pos_len);
  if (B <= 0 || T <= 0) return 0;
#ifndef CARCA_EMU
  if ((variant >= 2 && variant <= 6) || (variant == 0 && tc_ok)) {
    // two-head cross-attention decoder: 3 -> tcgen05 score MMAs inside the kernel, 4 -> fp32 loop with one row per
    // thread, 5 -> fp32 loop over candidate pairs, 6 -> separate decoder kernel over all (user, pair) items; otherwise
    // (-1) the separate kernel for the long candidate lists of catalog mode (8.4 G scores/s; pairs 7.0, rows 5.7) and
    // one row per thread for sampled candidates (22.5 M users/s; separate kernel 21.9, pairs 20.9, tcgen05 19.8) — or,
    // decided on the device, the tcgen05 decoder when the batch's profiles are dense
    const int dec = variant == 3 ? 0 : variant == 4 ? 1 : variant == 5 ? 2 : variant == 6 ? 3 : -1;
    return eval_forward_tc(y, ldy, col0, plan, m, p_x, p_c, o_x, o_c, B, L, T, status, dbg, dbg_stage, cat_lo,
                           ctx_per_user, scratch, stream, dec);
  }
#endif
  const PlanLayout pl = plan_layout(m);
  FusedArgs a;
  memset(&a, 0, sizeof(a));
  a.Tfold = plan + pl.tfold;
  a.Mc = plan + pl.mc;
  a.pos = m->embed.pos;
  a.p_x = p_x; a.p_c = p_c; a.o_x = o_x; a.o_c = o_c;
  a.y = y; a.ldy = ldy; a.col0 = col0;
  a.B = B; a.L = L; a.T = T; a.C = 8; a.H = H;
  a.n_blocks = m->n_blocks;
  a.residual_sa = m->residual_sa; a.residual_ca = m->residual_ca; a.decoder = m->decoder_kind;
  for (int b = 0; b < m->n_blocks; ++b) {
    const carca_block_params& bp = m->blocks[b];
    const float* wt = plan + pl.blocks + (long long)b * 5 * d * d;
    FusedBlockW& f = a.blk[b];
    f.ln1_g = bp.ln1_g; f.ln1_b = bp.ln1_b; f.ln2_g = bp.ln2_g; f.ln2_b = bp.ln2_b;
    f.bq = bp.bq; f.bk = bp.bk; f.bv = bp.bv; f.b1 = bp.b1; f.b2 = bp.b2;
    f.wqT = wt; f.wkT = wt + d * d; f.wvT = wt + 2 * d * d; f.w1T = wt + 3 * d * d; f.w2T = wt + 4 * d * d;
  }
  a.fn_g = m->norm_g; a.fn_b = m->norm_b;
  if (m->decoder_kind == 1) {
    const float* wt = plan + pl.cross;
    a.dwqT = wt; a.dwkT = wt + d * d; a.dwvT = wt + 2 * d * d;
    a.dbq = m->cross.bq; a.dbk = m->cross.bk; a.dbv = m->cross.bv; a.dwf = m->cross.wf; a.dbf = m->cross.bf;
  }
  // the kernel indexes Mc with row stride 8 and reads C context values per position
  a.C = m->embed.n_ctx;
  a.cat_lo = cat_lo;
  a.oc_user = ctx_per_user ? m->embed.n_ctx : (long long)T * m->embed.n_ctx;
  a.oc_tgt = ctx_per_user ? 0 : m->embed.n_ctx;
  const size_t smem = sizeof(FusedSmem);
  auto k = fused_eval_kernel;
  TRY(allow_smem(k, smem));
  const int n_tiles = ceil_div(B, FU);
  CARCA_LAUNCH(k, dim3(min(n_tiles, 148)), dim3(FTHREADS), smem, S(stream), a);
  return check_launch("fused_eval");
}

int64_t carca_eval_scratch_bytes(int B) {
  const int64_t rows = ((int64_t)B + 1) * 64;
  return eval_scratch_pack_bytes(B) + 4 * (rows * 64 + rows * 4 + 2ll * B);
}

int carca_eval_forward_opts(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                            const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                            int T, int variant, int32_t* status, float* dbg, int dbg_stage, void* scratch,
                            void* stream) {
  // variant bit 8: o_c holds one context row per user ([B, C], e.g. the base of an expanded [B,T,C] view)
  return eval_forward_any(y, ldy, col0, plan, m, p_x, p_c, o_x, o_c, B, L, T, variant & 0xff, status, dbg, dbg_stage,
                          0, (variant >> 8) & 1, scratch, stream);
}

int carca_eval_forward_catalog(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                               const int32_t* p_x, const float* p_c, const float* ctx_user, int32_t item_lo,
                               int n_cand, int B, int L, int variant, int32_t* status, void* scratch, void* stream) {
  CARCA_REQUIRE(item_lo >= 1 && n_cand >= 0 && (long long)item_lo + n_cand <= m->embed.n_items,
                "eval_forward_catalog: item range [%d, %d) outside the item table [1, %d)", item_lo, item_lo + n_cand,
                m->embed.n_items);
  return eval_forward_any(y, ldy, col0, plan, m, p_x, p_c, nullptr, ctx_user, B, L, n_cand, variant, status, nullptr,
                          0, item_lo, 1, scratch, stream);
}

int carca_catalog_rank_count(int32_t* count, const float* y, int64_t ldy, const float* y_pos, const int32_t* pos_item,
                             int32_t item_lo, int B, int n_cand, void* stream) {
  if (B <= 0 || n_cand <= 0) return 0;
  auto k = catalog_rank_count_kernel;
  CARCA_LAUNCH(k, dim3(warp_rows_grid(B)), dim3(256), 0, S(stream), count, y, (long long)ldy, y_pos, pos_item, item_lo,
               B, n_cand);
  return check_launch("catalog_rank_count");
}

int carca_eval_forward(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                       const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                       int T, void* stream) {
  // FFMA kernel unless the caller supplies a status word (carca_eval_forward_opts) for the tensor-core one
  return carca_eval_forward_opts(y, ldy, col0, plan, m, p_x, p_c, o_x, o_c, B, L, T, 1, nullptr, nullptr, 0, nullptr,
                                 stream);
}

}  // extern "C"
