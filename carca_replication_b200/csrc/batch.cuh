// Device-side batch construction: the windows, labels and sampled negatives the reference builds per
// user in Python (src/data.py:53-87 pad_profile / sample_negatives, :90-137 get_train_sequences,
// :140-192 get_test_sequences), straight from a device-resident interaction log (CSR over users).
// Output = exactly the tensors CARCA.forward consumes in table mode: ids + context, no attribute
// tensors (those stay in the ItemAttrTable).  One warp per user.
//
// Deterministic parts (window, left padding, positive, contexts, labels) equal the reference bit for
// bit.  Negatives are drawn like sample_negatives: uniform over [1, n_items-1], distinct, never an
// item of the user's WHOLE profile — from a Philox stream keyed by (seed, user, draw), 32 draws per
// round with in-warp duplicate rejection, so they are reproducible but not Python's `random` stream.
#pragma once
#include "common.cuh"

namespace carca {

constexpr int kMaxNeg = 512;   // negatives per user held in shared memory while sampling

struct InteractionLog {
  const int* rowptr;    // [n_users + 1]
  const int* items;     // [nnz], chronological per user
  const float* ctx;     // [nnz, C]: context of the (user, item) interaction (src/data.py:17-26)
  int n_users, n_ctx;
};

// pad_profile (src/data.py:53-74): index range [start, end) of the profile, empty when too short
__host__ __device__ __forceinline__ void window_of(int len, int max_len, int mode, int test, int& start, int& end) {
  start = 0;
  end = 0;
  int n_ex, min_end, min_len;
  if (mode == 0) { n_ex = test ? 2 : 1; min_end = 1; min_len = 1; }        // train
  else if (mode == 1) { n_ex = test ? 1 : 0; min_end = 2; min_len = 2; }   // val
  else { n_ex = 0; min_end = 3; min_len = 3; }                             // test
  if (len > min_len) {
    start = max(0, len - n_ex - max_len - 1);
    end = max(min_end, len - n_ex);
  }
}

// One warp samples `need` distinct items in [1, n_items-1] outside the user's profile into out[] (shared).
__device__ __forceinline__ void sample_negatives_warp(int* out, int need, const int* __restrict__ prof, int len,
                                                      int n_items, unsigned long long seed, unsigned user) {
  const int lane = threadIdx.x % kWarp;
  int count = 0;
  const unsigned span = (unsigned)(n_items - 1);
  // fewer free items than requested (tiny catalogs): take what exists — the reference would loop forever
  for (unsigned round = 0; count < need && round < 4096u; ++round) {
    const Philox4 r = philox4x32_10(round, user, lane, 0x6e656773u, (unsigned)seed, (unsigned)(seed >> 32));
    const int cand = 1 + (int)__umulhi(r.x, span);
    bool ok = true;
    for (int j = 0; j < len && ok; ++j) ok = prof[j] != cand;
    for (int j = 0; j < count && ok; ++j) ok = out[j] != cand;
    for (int l = 0; l < kWarp; ++l) {   // duplicates inside the round: the lower lane keeps the item
      const int other = __shfl_sync(kFull, cand, l);
      const int other_ok = __shfl_sync(kFull, (int)ok, l);
      if (l < lane && other_ok && other == cand) ok = false;
    }
    const unsigned acc = __ballot_sync(kFull, ok);
    const int slot = count + __popc(acc & ((1u << lane) - 1u));
    if (ok && slot < need) out[slot] = cand;
    count = min(need, count + __popc(acc));
    __syncwarp();
  }
  for (int j = count + lane; j < need; j += kWarp) out[j] = 0;
  __syncwarp();
}

// get_test_sequences (src/data.py:140-192).  p_x [B,L], p_c [B,L,C], o_x [B,T] (T = 1 + negatives, positive in
// column 0), o_c_user [B,C] (the positive's context, which every candidate carries, :185), y_true [B,T].
__global__ void __launch_bounds__(128) build_eval_batch_kernel(int* __restrict__ p_x, float* __restrict__ p_c,
                                                               int* __restrict__ o_x, float* __restrict__ o_c_user,
                                                               int* __restrict__ y_true, InteractionLog log,
                                                               const int* __restrict__ users, int B, int L, int T,
                                                               int n_items, int mode, int test,
                                                               unsigned long long seed) {
  __shared__ int negs[4][kMaxNeg];
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int b = blockIdx.x * 4 + w;
  if (b >= B) return;
  const int C = log.n_ctx;
  const int usr = users[b];
  const int base = log.rowptr[usr], len = log.rowptr[usr + 1] - base;
  int start, end;
  window_of(len, L, mode, test, start, end);
  const int wlen = max(0, end - start - 1);        // profile positions: padded[:-1], right-aligned (:172-173)
  for (int idx = lane; idx < L; idx += kWarp) {
    const int j = idx - (L - wlen);
    const int pi = base + start + j;
    p_x[(long long)b * L + idx] = j >= 0 ? log.items[pi] : 0;
    for (int c = 0; c < C; ++c) p_c[((long long)b * L + idx) * C + c] = j >= 0 ? log.ctx[(long long)pi * C + c] : 0.f;
  }
  const bool any = end > start;
  const int one_out = base + end - 1;              // :162-165
  if (lane == 0) {
    o_x[(long long)b * T] = any ? log.items[one_out] : 0;
    y_true[(long long)b * T] = 1;                  // :189-190
  }
  for (int c = lane; c < C; c += kWarp) o_c_user[(long long)b * C + c] = any ? log.ctx[(long long)one_out * C + c] : 0.f;
  sample_negatives_warp(negs[w], T - 1, log.items + base, len, n_items, seed, (unsigned)usr);   // :160
  for (int t = lane; t < T - 1; t += kWarp) {
    o_x[(long long)b * T + 1 + t] = any ? negs[w][t] : 0;
    y_true[(long long)b * T + 1 + t] = 0;
  }
}

// get_train_sequences (src/data.py:90-137).  p_x [B,L], p_c [B,L,C], o_x [B,2L] = next items | negatives,
// o_c [B,2L,C] (negatives take the positive's context, :130), y_true [B,2L] = [p_x > 0 | 0] (:134-135).
__global__ void __launch_bounds__(128) build_train_batch_kernel(int* __restrict__ p_x, float* __restrict__ p_c,
                                                                int* __restrict__ o_x, float* __restrict__ o_c,
                                                                int* __restrict__ y_true, InteractionLog log,
                                                                const int* __restrict__ users, int B, int L,
                                                                int n_items, int test, unsigned long long seed) {
  __shared__ int negs[4][kMaxNeg];
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int b = blockIdx.x * 4 + w;
  if (b >= B) return;
  const int C = log.n_ctx;
  const int usr = users[b];
  const int base = log.rowptr[usr], len = log.rowptr[usr + 1] - base;
  int start, end;
  window_of(len, L, 0, test, start, end);
  const int wlen = max(0, end - start - 1);
  sample_negatives_warp(negs[w], wlen, log.items + base, len, n_items, seed, (unsigned)usr);   // :108
  for (int idx = lane; idx < L; idx += kWarp) {
    const int j = idx - (L - wlen);                // j-th position of the window, oldest first
    const bool on = j >= 0;
    const int pi = base + start + j;
    const long long r = (long long)b * L + idx, ro = (long long)b * 2 * L + idx;
    p_x[r] = on ? log.items[pi] : 0;
    o_x[ro] = on ? log.items[pi + 1] : 0;          // the next item is the positive (:116)
    o_x[ro + L] = on ? negs[w][wlen - 1 - j] : 0;  // i = wlen-1-j enumerates from the newest (:111-117)
    y_true[ro] = on && log.items[pi] != 0;
    y_true[ro + L] = 0;
    for (int c = 0; c < C; ++c) {
      const float pc = on ? log.ctx[(long long)pi * C + c] : 0.f;
      const float nc = on ? log.ctx[(long long)(pi + 1) * C + c] : 0.f;
      p_c[r * C + c] = pc;
      o_c[ro * C + c] = nc;
      o_c[(ro + L) * C + c] = nc;
    }
  }
}

// Left-padded windows rebuilt on the device from their packed form (host -> device transfer diet): offs [B + 1] is
// the exclusive scan of the users' valid lengths, rows [R, 1 + C] holds per valid position the item id (as int bits)
// followed by its C context values, oldest first.  p_x [B, L] / p_c [B, L, C] get the window layout of
// src/data.py:53-74,112-113 (padding on the left, zeros).
__global__ void __launch_bounds__(256) unpack_windows_kernel(int* __restrict__ p_x, float* __restrict__ p_c,
                                                             const int* __restrict__ offs, const float* __restrict__ rows,
                                                             int B, int L, int C) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)B * L) return;
  const int u = (int)(i / L), pos = (int)(i % L);
  const int o0 = offs[u], n = min(offs[u + 1] - o0, L);
  const int j = pos - (L - n);
  float* c = p_c + i * C;
  if (j < 0) {
    p_x[i] = 0;
    for (int k = 0; k < C; ++k) c[k] = 0.f;
    return;
  }
  const float* r = rows + (long long)(o0 + j) * (1 + C);
  p_x[i] = __float_as_int(r[0]);
  for (int k = 0; k < C; ++k) c[k] = r[1 + k];
}

}  // namespace carca
