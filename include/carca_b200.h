/*
 * carca_b200.h — C ABI of libcarca_b200.so, the B200 (sm_100a) implementation of CARCA's
 * forward/backward + candidate-scoring hot path.
 *
 * Reference: r-papso/carca-replication.  Every entry point names the reference interface
 * (file:line, relative to the reference root) it replaces.  The reference is pure Python over
 * torch ATen ops; the binding a maintainer adds is a ctypes stub, and the
 * in-tree host mirror of the reference's classes is carca_replication_b200/ (Python).
 *
 * Conventions
 *  - All pointers are DEVICE pointers unless the name ends in `_host`.  Row-major, fp32,
 *    int32 ids (0 = padding), fp32 0/1 masks — the reference's tensor conventions
 *    (src/data.py:100-107, src/utils.py:6-7).
 *  - `stream` is a cudaStream_t passed as void*.  Calls only enqueue work; nothing synchronises
 *    except the *_host entry points, which say so.
 *  - Return value: 0 on success, negative on error; carca_last_error() returns the message.
 *    There is no CPU fallback: without a CUDA device every compute call fails.
 *  - Dropout uses a counter-based Philox4x32-10 stream keyed by (seed, site, element index)
 *    so the backward pass regenerates the forward mask (csrc/common.cuh, oracle/philox.py).
 */
#ifndef CARCA_B200_H
#define CARCA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARCA_B200_ABI_VERSION 2

const char* carca_last_error(void);
int carca_abi_version(void);
/* number of CUDA kernels this library has launched since it was loaded (bench.py's gpu_launches) */
int64_t carca_launch_count(void);

/* Optional device-resident seed word: while set (non-NULL), every dropout site XORs *device_seed into its
 * seed when the kernel RUNS, so a training step captured in a CUDA graph draws fresh masks on every replay
 * (increment the word inside the graph).  Process-wide; NULL restores host-only seeds.               */
void carca_set_seed_source(const uint64_t* device_seed);

/* ------------------------------------------------------------------ attribute source */
enum { CARCA_ATTR_CSR = 0, CARCA_ATTR_TABLE = 1, CARCA_ATTR_DENSE = 2 };

/* Where a position's attribute vector a[p, 0:A] comes from.
 *  CSR   : item -> sparse row (multi-hot / sparse attributes), device-resident, indexed by id
 *  TABLE : item -> dense row of `dense` [n_items, A]           (src/data.py:28-35 `attrs`)
 *  DENSE : `dense` is the per-position tensor [P, A] the reference API passes
 *          (src/abstract.py:22 `a`, materialised by src/data.py:119-131)                      */
typedef struct {
  int kind;
  const int32_t* csr_rowptr; /* [n_items + 1] */
  const int32_t* csr_cols;   /* [nnz] in [0, A) */
  const float* csr_vals;     /* [nnz] */
  const float* dense;
} carca_attr_source;

/* ------------------------------------------------------------------ AllEmbedding */
/* Parameters of AllEmbedding (src/carca.py:67-83), state_dict layout kept:
 *   items_embed.weight [n_items, d]; feats_embed.weight [g, A+C], bias [g];
 *   joint_embed.weight [d, d+g], bias [d].                                                   */
typedef struct {
  int n_items, d, g, n_attrs, n_ctx;
  const float* items_embed;
  const float* feats_w;
  const float* feats_wT; /* [A+C, g] transposed copy made by carca_transpose (CSR path only) */
  const float* feats_b;
  const float* joint_w;
  const float* joint_b;
  const float* pos;      /* optional positional table [pos_len, d] added to non-target rows */
  int pos_len;
} carca_embed_params;

typedef struct {
  float* items_embed;
  float* feats_w;
  float* feats_b;
  float* joint_w;
  float* joint_b;
  float* pos;            /* optional, [pos_len, d] */
} carca_embed_grads;

/* dst[c, r] (= | +=) src[r, c] */
int carca_transpose(float* dst, const float* src, int rows, int cols, int accumulate, void* stream);

/* mask[i] = ids[i] != 0           replaces get_mask, src/utils.py:6-7 */
int carca_padding_mask(float* mask, const int32_t* ids, int64_t n, void* stream);
/* the same for a float32 input (the reference's get_mask is dtype-agnostic: where(x == 0, 0, 1); -0.0 counts as 0) */
int carca_padding_mask_f32(float* mask, const float* x, int64_t n, void* stream);

/* e[p,:] = mask[p] * ( Wj [ sqrt(d) E[x_p] | Wf [a_p | c_p] + bf ] + bj (+ pos[p % n_cols]) )
 * replaces AllEmbedding.forward, src/carca.py:85-95.
 *   x [P] ids, ctx [P, C], mask [P]; P = n_rows * n_cols positions (n_cols = sequence length,
 *   used only for the positional table); is_target != 0 skips the positional add (:91-92).
 *   q_out [P, g] receives the attribute/context projection (saved for the backward; required). */
int carca_embed_fwd(float* e, float* q_out, const carca_embed_params* w, const carca_attr_source* attrs,
                    const int32_t* x, const float* ctx, const float* mask, int n_rows, int n_cols,
                    int is_target, void* stream);

/* Accumulates parameter gradients of carca_embed_fwd into `grads` (caller zero-initialises).
 * de [P, d] is the upstream gradient; scratch_pd [2P, d] and scratch_pg [P, g] are workspaces;
 * scratch_wT [A, g] (CSR only) must be zero-initialised.  Row 0 of items_embed gets no gradient
 * (padding_idx=0, src/carca.py:73).                                                         */
int carca_embed_bwd(const carca_embed_grads* grads, const float* de, const float* q_saved,
                    const carca_embed_params* w, const carca_attr_source* attrs, const int32_t* x,
                    const float* ctx, const float* mask, int n_rows, int n_cols, int is_target,
                    float* scratch_pd, float* scratch_pg, float* scratch_wT, void* stream);

/* ------------------------------------------------------------------ embedding / decoder variants */
/* q [P, g] = Wf [a | c] + bf — the attribute(/context) projection shared by AllEmbedding, AttrCtxEmbedding
 * and AttrEmbedding (src/carca.py:86, :113, :138; n_ctx = 0 and ctx = NULL for AttrEmbedding).
 * Uses feats_w / feats_wT / feats_b, g, n_attrs, n_ctx of `w`.                                         */
int carca_feats_fwd(float* q, const carca_embed_params* w, const carca_attr_source* attrs, const int32_t* x,
                    const float* ctx, int P, void* stream);
/* d feats_w [g, A+C], d feats_b [g] accumulated from dq [P, g]; scratch_wT [A, g] zero-initialised (CSR only) */
int carca_feats_bwd(float* d_feats_w, float* d_feats_b, const float* dq, const carca_embed_params* w,
                    const carca_attr_source* attrs, const int32_t* x, const float* ctx, int P, float* scratch_wT,
                    void* stream);

/* out[p,:] = alpha * table[ids[p],:] — nn.Embedding lookup with the sqrt(d) scale of IdEmbedding / MLPIdEmbedding
 * (src/carca.py:161-162, :187-188); backward scatter-adds alpha * d_out into d_table, skipping id 0 (padding_idx). */
int carca_gather_rows_fwd(float* out, const float* table, const int32_t* ids, float alpha, int P, int d, void* stream);
int carca_gather_rows_bwd(float* d_table, const float* d_out, const int32_t* ids, float alpha, int P, int d,
                          void* stream);

/* out = (in + pos[position]) * mask — positional encoding of non-target rows (pos may be NULL) and the final mask
 * of every embedding (e.g. src/carca.py:116-120).  in/out [n_rows, n_cols, d], mask [n_rows, n_cols].
 * Backward: d_in = d_out * mask; d_pos [n_cols, d] += column sums (d_pos may be NULL).                 */
/* AllEmbedding.forward (src/carca.py:85-95) in inference through a folded item table:
 * out[p,:] = mask[p] * (T[x[p],:] + sum_k c[p,k] McT[k,:] + pos[p % n_cols,:]) with T [n_items, d] =
 * the unfolded op applied to every item id with a zero context and McT [C, d] = (Wj[:, d:] Wf[:, A:])^T — exact
 * re-association of the two bias-only linears (:86-89).  x [n_rows, n_cols] ids (0 = padding -> zero row), c
 * [n_rows, n_cols, C], pos [n_cols, d] or NULL, mask [n_rows, n_cols].  d % 4 == 0.                        */
int carca_embed_folded_fwd(float* out, const float* T, const float* McT, const int32_t* x, const float* c,
                           const float* pos, const float* mask, int n_rows, int n_cols, int d, int n_ctx, void* stream);

int carca_pos_mask_fwd(float* out, const float* in, const float* pos, const float* mask, int n_rows, int n_cols, int d,
                       void* stream);
int carca_pos_mask_bwd(float* d_in, float* d_pos, const float* d_out, const float* mask, int n_rows, int n_cols, int d,
                       void* stream);

/* WeightedDotProduct.forward (src/carca.py:377-395): profile position i is scaled by sum_{j<=i} gamma^j, both sides
 * optionally L2-normalised, then sigmoid(<p, o>) or (<p, o> + 1) / 2.  Same layout rules as carca_dot_score_*.  */
int carca_wdot_score_fwd(float* y, const float* p, const float* o, int B, int T, int Lp, int d, int per_position,
                         float gamma, int normalize, int64_t ldy, int col0, void* stream);
int carca_wdot_score_bwd(float* d_o, float* d_p, const float* dy, const float* y, const float* p, const float* o, int B,
                         int T, int Lp, int d, int per_position, float gamma, int normalize, int64_t ldy, int col0,
                         void* stream);

/* KNN.forward (src/knn.py:14-21): y[b, col0 + t] = <p_a[b, Lp-1, :], o_a[b, t, :]> on the dense attribute tensors */
int carca_knn_score(float* y, const float* p_a, const float* o_a, int B, int T, int Lp, int A, int64_t ldy, int col0,
                    void* stream);

/* ------------------------------------------------------------------ dropout (nn.Dropout) */
/* y = x * keep/(1-p) with the Philox stream (site, seed); also its own backward (apply to dy).
 * replaces self.dropout(p_e), src/carca.py:416                                              */
int carca_dropout(float* y, const float* x, int64_t n, float p, uint64_t seed, uint32_t site, void* stream);

/* ------------------------------------------------------------------ LayerNorm */
/* replaces nn.LayerNorm(d).forward, src/carca.py:298,304,421 (eps 1e-5, affine) */
int carca_layernorm_fwd(float* y, float* mean, float* rstd, const float* x, const float* gamma,
                        const float* beta, int rows, int d, void* stream);
int carca_layernorm_bwd(float* dx, float* dgamma, float* dbeta, const float* dy, const float* x,
                        const float* mean, const float* rstd, const float* gamma, int rows, int d,
                        int accumulate_dx, void* stream);

/* ------------------------------------------------------------------ linear layers */
/* y[M,N] = act(alpha * x[M,K] w[N,K]^T + bias) — nn.Linear / Conv1d(k=1), src/carca.py:238-240,307,311 */
int carca_linear_fwd(float* y, const float* x, const float* w, const float* bias, int M, int N, int K,
                     int act_leaky, void* stream);
/* dx[M,K] (= | +=) dy[M,N] w[N,K] */
int carca_linear_bwd_input(float* dx, const float* dy, const float* w, int M, int N, int K, int accumulate,
                           void* stream);
/* dw[N,K] += dy[M,N]^T x[M,K];  db[N] += colsum(dy)  (db may be NULL) */
int carca_linear_bwd_weight(float* dw, float* db, const float* dy, const float* x, int M, int N, int K,
                            void* stream);

/* ------------------------------------------------------------------ attention core */
/* Masked multi-head attention after the projections; replaces src/carca.py:242-260.
 *   Q [B,Lq,d], K,V [B,Lk,d], heads = column slices of width d/H, q_mask [B,Lq], k_mask [B,Lk];
 *   causal_on/diag: tril(diagonal=diag) (0 self, -1 cross-train, off cross-eval; :299,:339);
 *   W_out optional [B,H,Lq,Lk] = weights after the mask, before dropout (return_w, :262-263).  */
int carca_attention_fwd(float* O, float* W_out, const float* Q, const float* K, const float* V,
                        const float* q_mask, const float* k_mask, int B, int H, int Lq, int Lk, int d,
                        int causal_on, int diag, float p_drop, uint64_t seed, uint32_t site, void* stream);
int carca_attention_bwd(float* dQ, float* dK, float* dV, const float* dO, const float* Q, const float* K,
                        const float* V, const float* q_mask, const float* k_mask, int B, int H, int Lq,
                        int Lk, int d, int causal_on, int diag, float p_drop, uint64_t seed, uint32_t site,
                        void* stream);

/* ------------------------------------------------------------------ SelfAttentionBlock */
/* state_dict layout kept (src/carca.py:279-289): norm1/norm2 {weight,bias}[d],
 * attn.WQ/WK/WV {weight [d,d], bias [d]}, ffn_1/ffn_2 {weight [d,d,1], bias [d]}.          */
typedef struct {
  const float *ln1_g, *ln1_b, *wq, *bq, *wk, *bk, *wv, *bv, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
} carca_block_params;
typedef struct {
  float *ln1_g, *ln1_b, *wq, *bq, *wk, *bk, *wv, *bv, *ln2_g, *ln2_b, *w1, *b1, *w2, *b2;
} carca_block_grads;
/* Activations kept for the backward; every buffer is [B*L, d] unless noted. */
typedef struct {
  float *qn;          /* LN1(x) */
  float *mean1, *rstd1; /* [B*L] */
  float *Q, *K, *V;
  float *s;           /* attention (+ residual), the LN2 input */
  float *mean2, *rstd2; /* [B*L] */
  float *s2;          /* LN2 output */
  float *a1;          /* dropout(LeakyReLU(ffn_1(s2))) */
} carca_block_saved;

/* out = SelfAttentionBlock(x, mask); replaces src/carca.py:297-318.
 * x,out [B,L,d]; mask [B,L]; dropout sites 1+3*block, 2+3*block, 3+3*block when p_drop>0.   */
int carca_sa_block_fwd(float* out, const carca_block_saved* saved, const float* x, const float* mask,
                       const carca_block_params* w, int B, int L, int d, int H, int residual, float p_drop,
                       uint64_t seed, int block_index, void* stream);
/* dx [B,L,d] (overwritten); parameter grads accumulated into `grads`; scratch: 4 buffers [B*L,d] */
int carca_sa_block_bwd(float* dx, const carca_block_grads* grads, const float* dout, const float* x,
                       const float* mask, const carca_block_params* w, const carca_block_saved* saved, int B,
                       int L, int d, int H, int residual, float p_drop, uint64_t seed, int block_index,
                       float* scratch4, void* stream);

/* ------------------------------------------------------------------ decoders */
/* y[b, col0+t] = sigmoid(<p[b, per_position ? t : Lp-1], o[b,t]>); replaces DotProduct.forward,
 * src/carca.py:358-365 (per_position = self.training).  y has row stride ldy (the concatenated
 * [B, sum T] output of CARCA.forward, src/carca.py:431).                                     */
int carca_dot_score_fwd(float* y, const float* p, const float* o, int B, int T, int Lp, int d,
                        int per_position, int64_t ldy, int col0, void* stream);
/* d_o overwritten; d_p accumulated (+=) */
int carca_dot_score_bwd(float* d_o, float* d_p, const float* dy, const float* y, const float* p,
                        const float* o, int B, int T, int Lp, int d, int per_position, int64_t ldy, int col0,
                        void* stream);

/* CrossAttentionBlock (src/carca.py:323-336): attn.WQ/WK/WV {weight,bias}, ffn {weight [1,d], bias [1]} */
typedef struct {
  const float *wq, *bq, *wk, *bk, *wv, *bv, *wf, *bf;
} carca_cross_params;
typedef struct {
  float *wq, *bq, *wk, *bk, *wv, *bv, *wf, *bf;
} carca_cross_grads;
typedef struct {
  float *Q;  /* [B*T, d] */
  float *K;  /* [B*Lp, d] */
  float *V;  /* [B*Lp, d] */
  float *s;  /* [B*T, d] attention (+ residual) */
} carca_cross_saved;

/* y[b, col0+t] = sigmoid(ffn(MHA(o, p, p) (+ o))); replaces CrossAttentionBlock.forward,
 * src/carca.py:338-349.  causal -1 when training (:339).  Output is always [B, T]
 * (the reference's squeeze() collapse for B == 1, :346, is deliberately not reproduced).     */
int carca_cross_score_fwd(float* y, const carca_cross_saved* saved, const float* o, const float* o_mask,
                          const float* p, const float* p_mask, const carca_cross_params* w, int B, int T,
                          int Lp, int d, int H, int residual, int training, float p_drop, uint64_t seed,
                          uint32_t site, int64_t ldy, int col0, void* stream);
/* d_o overwritten, d_p accumulated; scratch4: 4 buffers of max(B*T, B*Lp)*d floats */
int carca_cross_score_bwd(float* d_o, float* d_p, const carca_cross_grads* grads, const float* dy,
                          const float* y, const carca_cross_saved* saved, const float* o, const float* o_mask,
                          const float* p, const float* p_mask, const carca_cross_params* w, int B, int T,
                          int Lp, int d, int H, int residual, int training, float p_drop, uint64_t seed,
                          uint32_t site, int64_t ldy, int col0, float* scratch4, void* stream);

/* ------------------------------------------------------------------ fused training core */
/* Everything between the embeddings and the probabilities of CARCA.forward in TRAIN mode
 * (src/carca.py:416-431: post-embedding dropout, the SelfAttentionBlocks :297-318, the final LayerNorm :421,
 * the decoder per target tuple :424-429 with causal -1, :339) as ONE forward and ONE backward kernel that
 * run on the ACTIVE positions only (profile id != 0 or a target id != 0), packed into 64-row bins
 * (csrc/fused_train.cuh).  Padded positions reach no loss term, so every result equals the per-op entry
 * points' (carca_sa_block_*, carca_cross_score_*, carca_dot_score_*, carca_layernorm_*, carca_dropout) on the
 * same inputs, dropout masks included (same Philox sites and element indices).
 * Supported: d == 64, L <= 256 with at most 64 ACTIVE positions per user (always true for L <= 64; otherwise the
 * caller checks before calling — rows[1] counts the users that did not fit, whose rows are then truncated),
 * n_heads in {1, 2, 4}, n_blocks <= 8, 1 or 2 target tuples of L positions each; returns -4 otherwise so the
 * caller can use the per-op entry points.                                                                */
typedef struct {
  int B, L, n_heads, n_blocks, n_tuples;
  int decoder_kind;                 /* 0 = DotProduct (position-wise, :360), 1 = CrossAttentionBlock */
  int residual_sa, residual_ca;
  float p_drop;
  uint64_t seed;
  const int32_t* p_x;               /* [B, L] */
  const int32_t* o_x[2];            /* [B, L] per tuple (src/train.py:86-88) */
  const float* p_e;                 /* [B, L, 64] AllEmbedding output of the profile (masked, before :416) */
  const float* o_e[2];              /* [B, L, 64] AllEmbedding output of each target tuple */
  const carca_block_params* blocks; /* HOST array of n_blocks entries */
  const float *norm_g, *norm_b;
  carca_cross_params cross;         /* decoder_kind == 1 */
  int32_t* rows;                    /* workspace, carca_train_core_rows_ints(B) int32 */
  float* saved;                     /* workspace, carca_train_core_saved_floats(...) floats: activations kept by
                                       the forward for the backward */
  /* Optional in-kernel AllEmbedding (src/carca.py:85-95) — embed != NULL: p_e / o_e are not read; the kernels
   * embed every row set themselves from ids + context through tables folded once per call,
   *   GT = (Wj[:, d:] Wf)^T [A + C, 64],  cst = Wj[:, d:] bf + bj,
   *   e = mask * (Wj[:, :d] sqrt(d) E[x] + sum_attr val GT[attr] + sum_k c_k GT[A + k] + cst (+ pos)),
   * an exact re-association of the two bias-only linears (as in carca_eval_prepare); their gradients are
   * un-folded by three GEMMs after the backward kernel.  attrs must be CARCA_ATTR_CSR.                      */
  const carca_embed_params* embed;
  const carca_attr_source* attrs;
  const float* p_c;                 /* [B, L, C] */
  const float* o_c[2];              /* [B, L, C] per tuple */
  float* fold;                      /* workspace, carca_train_core_fold_floats(embed) floats */
} carca_train_core;

/* Development aid (tools/train_phase_times.py): while set, thread 0 of CTA 0 of the fused training kernels writes
 * (label, clock64) pairs at its phase boundaries into device_buf [capacity, 2].  NULL switches it off.        */
void carca_train_core_set_ticks(int64_t* device_buf, int capacity);

int64_t carca_train_core_rows_ints(int B);
int64_t carca_train_core_saved_floats(int B, int n_blocks, int n_tuples);
int64_t carca_train_core_fold_floats(const carca_embed_params* embed);

/* y [B, ldy]: tuple t's probabilities at columns t*L .. t*L + L - 1 (the cat of src/carca.py:431) */
int carca_train_core_fwd(float* y, int64_t ldy, const carca_train_core* c, void* stream);

/* Backward of carca_train_core_fwd for the upstream gradient dy [B, ldy] (call after the forward with the same
 * `c`, whose workspaces still hold the forward's row maps, activations and folded tables).  Parameter gradients
 * are ACCUMULATED (caller zero-initialises; g_blocks is a HOST array); d_pe / d_oe[t] [B, L, 64] are written at
 * the active positions only (caller zero-initialises) and unused when c->embed != NULL.  With c->embed:
 * g_embed receives the AllEmbedding gradients (items_embed, joint_w and pos accumulated — zero-initialise;
 * feats_w, feats_b, joint_b overwritten) and d_fold is a workspace of carca_train_core_fold_floats floats.  */
int carca_train_core_bwd(float* d_pe, float* d_oe0, float* d_oe1, const carca_block_grads* g_blocks, float* g_norm_g,
                         float* g_norm_b, const carca_cross_grads* g_cross, const carca_embed_grads* g_embed,
                         float* d_fold, const float* dy, int64_t ldy, const carca_train_core* c, void* stream);

/* ------------------------------------------------------------------ loss */
/* sums[0] += sum(ell * mask), sums[1] += sum(mask); replaces src/carca.py:442-443 numerators.
 * y_true int32.  Under data parallelism `sums` is all-reduced before finalize.   */
int carca_bce_sums(float* sums, const float* y_pred, const int32_t* y_true, const float* mask, int64_t n,
                   float eps, void* stream);
int carca_bce_finalize(float* loss, const float* sums, void* stream);
/* dy = grad_out * mask / sums[1] * (-t/(y+eps) + (1-t)/(1-y+eps)) */
int carca_bce_bwd(float* dy, const float* grad_out, const float* sums, const float* y_pred,
                  const int32_t* y_true, const float* mask, int64_t n, float eps, void* stream);

/* ------------------------------------------------------------------ optimizer */
/* One Adam step over a list of parameter tensors in one launch (plus a one-thread counter tick):
 * torch.optim.Adam as scripts/training.py:174 builds it and src/train.py:96 steps it (weight_decay is added to
 * the gradient; no amsgrad):  g += wd p;  m += (g - m)(1 - b1);  v = b2 v + (1 - b2) g^2;
 * p -= lr / (1 - b1^t) * m / (sqrt(v) / sqrt(1 - b2^t) + eps).  `tensors` is a HOST array (device pointers
 * inside); `step` is the tensor's own device float counter t (torch keeps one per parameter), incremented by
 * this call before the update, which makes the call CUDA-graph capturable.                                 */
typedef struct {
  float* param;
  const float* grad;
  float* exp_avg;
  float* exp_avg_sq;
  float* step;
  int64_t numel;
} carca_adam_tensor;
int carca_adam_step(const carca_adam_tensor* tensors, int n_tensors, float lr, float beta1, float beta2, float eps,
                    float weight_decay, void* stream);

/* ------------------------------------------------------------------ ranking metrics */
/* acc[0] += hits@k, acc[1] += sum 1/log2(rank+2) over labelled candidates with rank < k,
 * acc[2] += B (fp64 device accumulators); replaces compute_HR / compute_NDCG,
 * src/train.py:15-32, incl. torch's stable tie order.  first_rank [B] optional (rank of the
 * first labelled candidate of each row).                                                     */
int carca_rank_metrics(double* acc, int32_t* first_rank, const float* y_pred, const int32_t* y_true, int B,
                       int T, int64_t ldy, int64_t ldt, int k, void* stream);

/* The per-batch reductions of evaluate() (src/train.py:44-50) in one launch: stats[0] += hits@k, stats[1] += sum
 * 1/log2(rank+2) (as carca_rank_metrics), stats[2] += B, stats[3] += the batch's masked BCE
 * sum(l * mask) / sum(mask) with mask = (o_x != 0) (src/carca.py:441-444, src/utils.py:6-7, eps on probabilities).
 * stats: fp64[4] device accumulators; work: fp64[3] device scratch, zero before the first call (the kernel leaves it
 * zero again, so the call can be replayed inside a CUDA graph).                                              */
int carca_eval_metrics(double* stats, double* work, const float* y_pred, const int32_t* y_true, const int32_t* o_x, int B,
                       int T, int64_t ldy, int64_t ldt, int64_t ldx, int k, float eps, void* stream);

/* ------------------------------------------------------------------ batch construction */
/* The users' interaction log on the device, CSR over users: items of user u (chronological) are
 * items[rowptr[u] .. rowptr[u+1]) and ctx[j, :] is the context of interaction j — what
 * load_profiles / load_ctx return (src/data.py:17-50), flattened once.                          */
typedef struct {
  const int32_t* rowptr; /* [n_users + 1] */
  const int32_t* items;  /* [nnz] */
  const float* ctx;      /* [nnz, n_ctx] */
  int n_users, n_ctx;
} carca_interactions;

/* One evaluation batch for `users` [B]: replaces CARCADataset.__getitem__ -> get_test_sequences
 * (src/data.py:140-192, pad_profile :53-74, sample_negatives :77-87) + default_collate.
 *   p_x [B,L] left-padded window, p_c [B,L,C], o_x [B,T] = positive | T-1 sampled negatives,
 *   o_c_user [B,C] = the positive's context (every candidate carries it, :185; pass it to
 *   CARCA.forward as an expanded [B,T,C] view), y_true [B,T] = [1, 0, ...].
 * mode 1 = val, 2 = test; test = the `test` flag of CARCADataset.  Windows, padding, positives,
 * contexts and labels equal the reference's; negatives are uniform over [1, n_items-1], distinct and
 * outside the user's whole profile like the reference's, drawn from Philox(seed, user) instead of
 * Python's `random`.  Users whose profile is too short for the mode get all-zero rows.          */
int carca_build_eval_batch(int32_t* p_x, float* p_c, int32_t* o_x, float* o_c_user, int32_t* y_true,
                           const carca_interactions* log, const int32_t* users, int B, int L, int T, int n_items,
                           int mode, int test, uint64_t seed, void* stream);

/* One training batch: replaces get_train_sequences (src/data.py:90-137).  o_x [B,2L] = next items |
 * negatives, o_c [B,2L,C] (negatives take the positive's context, :130), y_true [B,2L]. */
int carca_build_train_batch(int32_t* p_x, float* p_c, int32_t* o_x, float* o_c, int32_t* y_true,
                            const carca_interactions* log, const int32_t* users, int B, int L, int n_items, int test,
                            uint64_t seed, void* stream);

/* Host -> device transfer diet for eval / train batches: the loader's windows are LEFT-padded (src/data.py:53-74,
 * :112-113), so a batch of windows is fully described by the users' valid lengths and the valid positions' (id,
 * context) records; Beauty-shaped windows are ~86 % padding.  offs [B + 1] = exclusive scan of the lengths,
 * rows [R, 1 + C] = per valid position the id (int32 bits in a float slot) then its C context values, oldest first.
 * Writes p_x [B, L] and p_c [B, L, C] exactly as the dense tensors the reference loader would have sent.            */
int carca_unpack_windows(int32_t* p_x, float* p_c, const int32_t* offs, const float* rows, int B, int L, int C,
                         void* stream);

/* ------------------------------------------------------------------ whole-model inference */
/* Every parameter of a CARCA model (src/carca.py:401-409) in its state_dict layout.
 * `blocks` is a HOST array of n_blocks entries; decoder_kind 0 = DotProduct, 1 = CrossAttentionBlock. */
typedef struct {
  carca_embed_params embed;
  int n_blocks, n_heads, residual_sa, residual_ca, decoder_kind;
  const carca_block_params* blocks;
  const float *norm_g, *norm_b;
  carca_cross_params cross;
} carca_model_params;

/* Size in floats of the inference plan built by carca_eval_prepare. */
int64_t carca_eval_plan_floats(const carca_model_params* m);

/* Builds the inference plan for the current weights (call again whenever they change):
 *   - folded item table  T[i] = Wj [ sqrt(d) E[i] | Wf_a attrs[i] + bf ] + bj        [n_items, d]
 *   - folded context map Mc   = Wj[:, d:] Wf[:, A:]                                   [d, 8]
 *     so that AllEmbedding.forward (src/carca.py:85-95) becomes e = mask * (T[x] + Mc c (+ pos)),
 *     an exact re-association of the two bias-only linears (no nonlinearity between them);
 *   - the projection weights of every block / the decoder transposed to [in, out].
 * attrs must be CARCA_ATTR_CSR or CARCA_ATTR_TABLE (a device-resident item table).
 * scratch_q: [n_items, g] floats.                                                            */
int carca_eval_prepare(float* plan, float* scratch_q, const carca_model_params* m, const carca_attr_source* attrs,
                       void* stream);

/* y[b, col0 + t] = CARCA.forward(profile, [targets]) in eval mode (src/carca.py:411-431), one fused
 * kernel: embedding gather -> encoder blocks -> final LayerNorm -> decoder, activations never
 * leave the SM.  p_x [B,L], p_c [B,L,C], o_x [B,T], o_c [B,T,C].  Supported: d == 64, L <= 52,
 * C <= 8, n_blocks <= 8, n_heads in {1,2,4,8,16}; returns -4 (and sets the message) otherwise so
 * the caller can use the per-op entry points instead.                                         */
int carca_eval_forward(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                       const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                       int T, void* stream);

/* Same call with the kernel variant chosen explicitly: 0 = best available, 1 = fp32 FFMA kernel
 * (L <= 52), 2 = tcgen05 tensor-core kernel (n_heads in {2,4}; L <= 256 with at most 64 non-padding
 * positions per user — always true for L <= 64; activations in TMEM, 3xTF32 fp32-grade MMAs; with
 * two heads the cross-attention decoder runs in fp32 against the user's own keys: one thread per
 * candidate row inside the kernel, or — catalog mode — a second kernel over all (user, candidate
 * pair) items reading the keys the first one exported to the scratch buffer), 3 = variant 2 with
 * that decoder on tcgen05 score MMAs over 128-row tiles, 4 / 5 = the in-kernel row / pair loop
 * forced, 6 = the separate decoder kernel forced (one context row per user; otherwise as 2).  With
 * variants 0 / 2 and two-head cross-attention the call also launches the tcgen05-decoder kernel, and
 * the two decide ON THE DEVICE which of them runs: the tcgen05 decoder when the batch averages more
 * than 24 non-padding positions per user (the packing pass knows; the host would need a sync), the
 * fp32 decoder otherwise; the other launch returns at once.
 * status (device int32[1], required for variants 2 - 6): bit 0 is set if an MMA completion wait timed
 * out, bit 1 if some user had more than 64 non-padding positions (that user's scores are then
 * computed from its last 64 positions only: route such batches to the per-op entry points).  dbg (optional device [128,64]) receives the intermediate
 * activation `dbg_stage` of the first tile (10*block + {1: LN1, 2: Q, 3: K, 4: V, 5: attention +
 * residual, 9: block output}, 100: final LayerNorm) for stage-by-stage validation; dbg_stage -1: phase
 * clock ticks of CTA 0's first tile, -2 - k: start clocks of every tile CTA k processed (tools/tc_*.py).
 * variant | 0x100: o_c is [B, C], one context row per user shared by all of the user's candidates
 * (what src/data.py:185 builds; lets the host pass the base of an expanded [B,T,C] view).      */
int carca_eval_forward_opts(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                            const int32_t* p_x, const float* p_c, const int32_t* o_x, const float* o_c, int B, int L,
                            int T, int variant, int32_t* status, float* dbg, int dbg_stage, void* scratch,
                            void* stream);

/* Bytes of device scratch the tensor-core kernel needs for a batch of B users (variants 0, 2 - 6;
 * packing tables, plus per packed row the exported decoder keys of the split decoder): it
 * first packs the VALID profile positions of every user into 64-row bins — Beauty-shaped profiles
 * are mostly left padding (src/data.py:112-113) and padded rows influence nothing the decoder reads
 * (src/carca.py:246-251) — so the encoder runs on valid rows only.                              */
int64_t carca_eval_scratch_bytes(int B);

/* ------------------------------------------------------------------ inference over packed rows */
/* CARCA.forward in eval mode (src/carca.py:411-431) as a pipeline of kernels over the batch's PACKED valid profile
 * rows (csrc/rows_bf16.cuh): the valid positions of every user (plus position L-1) form one flat row array, every
 * stage of the model is one kernel over all rows, and a user may have ANY number of valid positions (L <= 256) —
 * there is no bin or tile limit and nothing is decided on the host.  Both decoders.  Two arithmetic flavours:
 *   precision 0 (bf16; BASELINE configs[1] "bf16/fp32", configs[2] Men-shaped d = 256): tcgen05 kind::f16 GEMMs with
 *     the weight matrix resident in shared memory, A tiles streamed by cp.async.bulk, double-buffered TMEM
 *     accumulators and fused bias / LeakyReLU / residual / LayerNorm epilogues; fp32 accumulation, softmax, LayerNorm,
 *     embedding gather and decoder.  d in {64, 256}, head width 32 or 64.  Scores within 1e-2 of the fp32 reference.
 *   precision 1 (fp32): every projection through the 3xTF32 tcgen05 GEMM of the training path with the row count
 *     read on the device; d in {32, 64, 128, 256}, head width 16 / 32 / 64.  Scores within 1e-4 (the fp32 contract).
 *
 * carca_rows_plan_bytes / carca_rows_prepare: plan derived from the fp32 plan of carca_eval_prepare (folded item
 * table T and context map Mc) and the model parameters: packed bf16 projection weights and, for the cross-attention
 * decoder, the folded candidate tables TQ = WQ T + bq, tw = <T, wf>, McQ = WQ Mc, mcw = wf Mc (fp32).  Call again
 * whenever the weights change; the forward call takes both plans.                                                   */
int64_t carca_rows_plan_bytes(const carca_model_params* m);
int carca_rows_prepare(void* plan, const float* plan_f32, const carca_model_params* m, void* stream);
/* Measurement aid (bench.py's per-kernel roofline): the next carca_rows_eval_forward calls in bf16 record the given
 * CUDA events (cudaEvent_t handles, at most 64) on their stream between the kernels of the pipeline, in launch order;
 * carca_rows_stage_ids returns how many the last call recorded and which stage each one closes (0 start, 1 pack,
 * 2 embedding + LN, 3 Q/K/V GEMM, 4 attention, 5 FFN-1, 6 FFN-2 + LN, 7 FFN chain (+ next Q/K/V or decoder K/V),
 * 8 decoder K/V GEMM, 9 decoder).  n = 0 switches the recording off.  Not thread safe; not for use under capture.  */
int carca_rows_set_stage_events(void* const* events, int n);
int carca_rows_stage_ids(int32_t* ids, int cap);
/* Bytes of device scratch for a batch of B users with windows of L positions (worst case: every position valid). */
int64_t carca_rows_scratch_bytes(const carca_model_params* m, int B, int L);
/* y[b, col0 + t] = CARCA.forward(profile, [targets]) in eval mode.  p_x [B,L], p_c [B,L,C], o_x [B,T],
 * o_c [B,T,C] (ctx_per_user != 0: [B,C], one context row per user as src/data.py:185 builds).  cat_lo > 0: catalog
 * mode, candidate t is item cat_lo + t and o_x is not read.  status (device int32[1]): bit 1 is set if an mbarrier
 * wait of the bf16 GEMM pipeline timed out (results invalid).                                                        */
int carca_rows_eval_forward(float* y, int64_t ldy, int col0, const void* plan, const float* plan_f32,
                            const carca_model_params* m, const int32_t* p_x, const float* p_c, const int32_t* o_x,
                            const float* o_c, int B, int L, int T, int ctx_per_user, int cat_lo, int precision,
                            int32_t* status, void* scratch, void* stream);

/* Full-catalog rank counts on the tensor cores (csrc/catalog_tc.cuh; d = 64, L <= 128, dot decoder or two-head
 * cross-attention): counts[b] += number of items of the shard [item_lo, item_lo + n_shard) that a stable descending
 * sort of the user's scores places before the positive pos_item[b] (same rule as carca_catalog_rank_count), computed
 * as a 3xTF32 tcgen05 GEMM between the folded item table and the users' packed key / profile vectors with the softmax,
 * sigmoid and comparison in its epilogue — the [B, n_shard] score matrix is never written.  ctx_user [B, C] is the
 * context every candidate of the user carries (src/data.py:185).  counts must be zeroed by the caller.              */
int64_t carca_rows_catalog_scratch_bytes(const carca_model_params* m, int B, int L);
int carca_rows_catalog_counts(int32_t* counts, const void* plan, const float* plan_f32, const carca_model_params* m,
                              const int32_t* p_x, const float* p_c, const float* ctx_user, const int32_t* pos_item,
                              int item_lo, int n_shard, int B, int L, int32_t* status, void* scratch, void* stream);

/* ------------------------------------------------------------------ full-catalog scoring */
/* Scores every item of the contiguous id range [item_lo, item_lo + n_cand) (an item-table shard)
 * for every user: y[b, col0 + j] = CARCA.forward(profile_b, [(item_lo + j, ., ctx_user_b)]) in eval
 * mode.  ctx_user [B, C] is one context row per user, used for all of that user's candidates (the
 * positive's context, as src/data.py:185 gives the sampled negatives).  The reference has no such
 * mode; by construction it equals model.forward(profile, targets=[chunk_1, ..]) over candidate
 * chunks (src/carca.py:424-431).  Same kernels, plan, variants and limits as carca_eval_forward_opts;
 * no candidate id / context tensors are read.                                                   */
int carca_eval_forward_catalog(float* y, int64_t ldy, int col0, const float* plan, const carca_model_params* m,
                               const int32_t* p_x, const float* p_c, const float* ctx_user, int32_t item_lo,
                               int n_cand, int B, int L, int variant, int32_t* status, void* scratch, void* stream);

/* count[b] += number of candidates j of this shard that a stable descending sort (src/train.py:16)
 * ranks before the positive: y[b,j] > y_pos[b], or equal with item_lo + j < pos_item[b].  Summed
 * over shards (all-reduce under item-table sharding) it is the positive's rank: hit = rank < k,
 * ndcg = 1 / log2(rank + 2) (src/train.py:18-21, :27-32).                                        */
int carca_catalog_rank_count(int32_t* count, const float* y, int64_t ldy, const float* y_pos, const int32_t* pos_item,
                             int32_t item_lo, int B, int n_cand, void* stream);

/* ------------------------------------------------------------------ tensor-core self test */
/* C[128,N] = A[128,K] B[N,K]^T on the tcgen05 tensor cores (tf32), accumulator in TMEM.
 * mode bits: 1 = 3xTF32 fp32-grade split; 2 = A operand read from TMEM (written by tcgen05.st)
 * instead of shared memory; 4 = B operand stored MN-major instead of K-major.  status[0] = 1 if an
 * mbarrier wait timed out (the kernel never hangs).  N % 16 == 0, K % 8 == 0.  Known-answer test of
 * the primitives in csrc/umma.cuh that the fused kernels build on. */
int carca_umma_selftest(float* C, const float* A, const float* B, int N, int K, int mode, int32_t* status,
                        void* stream);

/* Layout probe for the same primitives: raw shared-memory images of both operands and every
 * descriptor field come from the host (tools/probe_umma_layout.py).  Development aid. */
int carca_umma_probe(float* C, const float* a_img, int a_floats, const float* b_img, int b_floats, int N,
                     int ksteps, uint32_t a_lbo, uint32_t a_sbo, uint32_t a_step, uint32_t b_lbo, uint32_t b_sbo,
                     uint32_t b_step, uint32_t idesc, int32_t* status, void* stream);

/* ------------------------------------------------------------------------------------------------------------------
 * Peer all-reduce (csrc/peer.cu): sum of a flat fp32 tensor over the GPUs of one NVLink / NVSwitch box in ONE kernel
 * on peer-mapped memory — the gradient exchange of the data-parallel training step (the reference is
 * single-device: src/train.py:90-97 has no counterpart for it).  Every rank allocates a communication buffer of
 * carca_peer_buffer_bytes(n) bytes, exports its 64-byte CUDA IPC handle, and opens the handles of its peers (the
 * handles travel through whatever the host uses for bootstrap, e.g. torch.distributed.all_gather_object).
 * carca_peer_allreduce: `local` (n floats, 16-byte aligned) is summed IN PLACE over all ranks; `bases` is a HOST array
 * of `world` device pointers (bases[rank] = this rank's own buffer, the others as returned by carca_peer_open); all
 * ranks must call it with the same n, in the same order.  Every rank receives bit-identical sums (each element is
 * reduced by one owner rank in rank order).  status (device int32[1]): bit 8 is set if a peer did not arrive within
 * ~20 s (results invalid) — a wait never hangs the GPU.  world <= 8.  carca_peer_data_offset: byte offset of the data
 * region inside a communication buffer; a tensor that lives there (local == own base + offset, its tail up to the
 * next multiple of 4 floats readable and writable) is reduced in place without staging copies.                    */
int64_t carca_peer_data_offset(void);
int64_t carca_peer_buffer_bytes(int64_t n_floats);
int carca_peer_alloc(void** base, int64_t bytes);
int carca_peer_free(void* base);
int carca_peer_export(void* base, void* handle64);
int carca_peer_open(const void* handle64, void** base);
int carca_peer_close(void* base);
int carca_peer_allreduce(float* local, int64_t n, void* const* bases, int rank, int world, int32_t* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CARCA_B200_H */
