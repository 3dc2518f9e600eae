// Candidate scoring heads, masked BCE and the HR@k / NDCG@k ranking reduction.
//   DotProduct.forward            src/carca.py:358-365
//   CrossAttentionBlock ffn+sig   src/carca.py:345-347
//   BinaryCrossEntropy.forward    src/carca.py:441-444
//   compute_HR / compute_NDCG     src/train.py:15-32
#pragma once
#include "common.cuh"

namespace carca {

__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

// y[b, col0 + t] = sigmoid( <p[b, pi(t), :], o[b, t, :]> ),  pi(t) = per_position ? t : Lp-1
__global__ void __launch_bounds__(256) dot_score_fwd_kernel(float* __restrict__ y, const float* __restrict__ p,
                                                            const float* __restrict__ o, int B, int T, int Lp,
                                                            int d, int per_position, long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const int pi = per_position ? t : Lp - 1;
  const float* pr = p + ((long long)b * Lp + pi) * d;
  const float* orow = o + ((long long)b * T + t) * d;
  float acc = 0.f;
  for (int j = lane; j < d; j += kWarp) acc = fmaf(pr[j], orow[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[(long long)b * ldy + col0 + t] = sigmoidf_(acc);
}

// dlogit = dy*y*(1-y);  do[b,t,:] = dlogit * p[b,pi,:];  dp[b,pi,:] += dlogit * o[b,t,:]
__global__ void __launch_bounds__(256) dot_score_bwd_kernel(float* __restrict__ d_o, float* __restrict__ d_p,
                                                            const float* __restrict__ dy,
                                                            const float* __restrict__ y,
                                                            const float* __restrict__ p,
                                                            const float* __restrict__ o, int B, int T, int Lp, int d,
                                                            int per_position, long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const int pi = per_position ? t : Lp - 1;
  const float yv = y[(long long)b * ldy + col0 + t];
  const float dl = dy[(long long)b * ldy + col0 + t] * yv * (1.0f - yv);
  const float* pr = p + ((long long)b * Lp + pi) * d;
  const float* orow = o + ((long long)b * T + t) * d;
  float* dor = d_o + ((long long)b * T + t) * d;
  float* dpr = d_p + ((long long)b * Lp + pi) * d;
  for (int j = lane; j < d; j += kWarp) {
    dor[j] = dl * pr[j];
    if (per_position) dpr[j] += dl * orow[j];        // one writer per (b, t)
    else atomicAdd(dpr + j, dl * orow[j]);            // every candidate hits row Lp-1
  }
}

// y[r_out] = sigmoid(<s[r,:], w> + bias)
__global__ void __launch_bounds__(256) rowdot_sigmoid_fwd_kernel(float* __restrict__ y, const float* __restrict__ s,
                                                                 const float* __restrict__ w,
                                                                 const float* __restrict__ bias, int B, int T, int d,
                                                                 long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const float* sr = s + warp * d;
  float acc = 0.f;
  for (int j = lane; j < d; j += kWarp) acc = fmaf(sr[j], w[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[(long long)b * ldy + col0 + t] = sigmoidf_(acc + bias[0]);
}

// ds[r,:] = dlogit_r * w;  dw += sum_r dlogit_r * s[r,:];  db += sum_r dlogit_r
// Dynamic smem: warps_per_block * d floats (+ warps for db).
__global__ void __launch_bounds__(256) rowdot_sigmoid_bwd_kernel(float* __restrict__ ds, float* __restrict__ dw,
                                                                 float* __restrict__ db,
                                                                 const float* __restrict__ dy,
                                                                 const float* __restrict__ y,
                                                                 const float* __restrict__ s,
                                                                 const float* __restrict__ w, int B, int T, int d,
                                                                 long long ldy, int col0) {
  CARCA_DYN_SMEM(float, sm);
  const int warps = blockDim.x / kWarp;
  const int wi = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* sw = sm + (long long)wi * d;
  float* sb = sm + (long long)warps * d;
  for (int j = lane; j < d; j += kWarp) sw[j] = 0.f;
  float bacc = 0.f;
  const long long rows = (long long)B * T;
  for (long long r = (long long)blockIdx.x * warps + wi; r < rows; r += (long long)gridDim.x * warps) {
    const int b = (int)(r / T), t = (int)(r % T);
    const float yv = y[(long long)b * ldy + col0 + t];
    const float dl = dy[(long long)b * ldy + col0 + t] * yv * (1.0f - yv);
    const float* sr = s + r * d;
    float* dsr = ds + r * d;
    for (int j = lane; j < d; j += kWarp) {
      dsr[j] = dl * w[j];
      sw[j] = fmaf(dl, sr[j], sw[j]);
    }
    bacc += dl;
  }
  if (lane == 0) sb[wi] = bacc;
  __syncthreads();
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float acc = 0.f;
    for (int ww = 0; ww < warps; ++ww) acc += sm[(long long)ww * d + j];
    atomicAdd(dw + j, acc);
  }
  if (threadIdx.x == 0) {
    float acc = 0.f;
    for (int ww = 0; ww < warps; ++ww) acc += sb[ww];
    atomicAdd(db, acc);
  }
}

// sums[0] += sum ell*mask, sums[1] += sum mask,  ell = -(t log(y+eps) + (1-t) log(1-y+eps))
__global__ void __launch_bounds__(256) bce_sums_kernel(float* __restrict__ sums, const float* __restrict__ y,
                                                       const int* __restrict__ t, const float* __restrict__ mask,
                                                       long long n, float eps) {
  __shared__ float part[2][8];
  float a = 0.f, m = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const float tv = (float)t[i], yv = y[i], mv = mask[i];
    const float ell = -(tv * logf(yv + eps) + (1.0f - tv) * logf(1.0f - yv + eps));
    a = fmaf(ell, mv, a);
    m += mv;
  }
  a = warp_sum(a);
  m = warp_sum(m);
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  if (lane == 0) { part[0][w] = a; part[1][w] = m; }
  __syncthreads();
  if (threadIdx.x == 0) {
    float sa = 0.f, smk = 0.f;
    for (int i = 0; i < (int)(blockDim.x / kWarp); ++i) { sa += part[0][i]; smk += part[1][i]; }
    atomicAdd(sums + 0, sa);
    atomicAdd(sums + 1, smk);
  }
}

// loss[0] = sums[0] / sums[1]
__global__ void bce_finalize_kernel(float* __restrict__ loss, const float* __restrict__ sums) {
  if (threadIdx.x == 0 && blockIdx.x == 0) loss[0] = sums[0] / sums[1];
}

// dy[i] = gout * mask/sum_mask * (-t/(y+eps) + (1-t)/(1-y+eps))
__global__ void __launch_bounds__(256) bce_bwd_kernel(float* __restrict__ dy, const float* __restrict__ gout,
                                                      const float* __restrict__ sums, const float* __restrict__ y,
                                                      const int* __restrict__ t, const float* __restrict__ mask,
                                                      long long n, float eps) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float tv = (float)t[i], yv = y[i];
  const float g = gout[0] / sums[1];
  dy[i] = g * mask[i] * (-tv / (yv + eps) + (1.0f - tv) / (1.0f - yv + eps));
}

// Stable descending rank of every labelled candidate:
//   rank_j = #{i : s_i > s_j} + #{i < j : s_i == s_j}      (torch CPU sort keeps ties in index order)
//   hits += y_true_j * [rank_j < k];   ndcg += [y_true_j != 0][rank_j < k] / log2(rank_j + 2)
// acc[0] += hits, acc[1] += ndcg, acc[2] += rows (double).  One warp per row.
__global__ void __launch_bounds__(256) rank_metrics_kernel(double* __restrict__ acc, int* __restrict__ first_rank,
                                                           const float* __restrict__ y, const int* __restrict__ yt,
                                                           int B, int T, long long ldy, long long ldt, int k) {
  __shared__ double part[2][8];
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int warps = blockDim.x / kWarp;
  double hits = 0.0, ndcg = 0.0;
  for (int b = blockIdx.x * warps + w; b < B; b += gridDim.x * warps) {
    const float* yr = y + (long long)b * ldy;
    const int* tr = yt + (long long)b * ldt;
    int first = -1;
    for (int j0 = 0; j0 < T; j0 += kWarp) {
      const int jj = j0 + lane;
      const int lab = jj < T ? tr[jj] : 0;
      unsigned pos = __ballot_sync(kFull, lab != 0);
      while (pos) {
        const int l = __ffs((int)pos) - 1;
        pos &= pos - 1;
        const int j = j0 + l;
        const float sj = yr[j];
        const int labj = __shfl_sync(kFull, lab, l);
        int cnt = 0;
        for (int i = lane; i < T; i += kWarp) {
          const float si = yr[i];
          cnt += (si > sj) || (si == sj && i < j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
        if (first < 0) first = cnt;
        if (cnt < k) {
          hits += (double)labj;
          ndcg += (double)(1.0f / log2f((float)(cnt + 2)));
        }
      }
    }
    if (first_rank && lane == 0) first_rank[b] = first;
  }
  if (lane == 0) { part[0][w] = hits; part[1][w] = ndcg; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double h = 0.0, n = 0.0;
    for (int i = 0; i < warps; ++i) { h += part[0][i]; n += part[1][i]; }
    atomicAdd(acc + 0, h);
    atomicAdd(acc + 1, n);
    if (blockIdx.x == 0) atomicAdd(acc + 2, (double)B);
  }
}

// Full-catalog ranking: count[b] += #{j : y[b,j] > y_pos[b]  or  (y[b,j] == y_pos[b] and item_lo + j < pos_item[b])},
// i.e. how many items of this shard a stable descending sort (src/train.py:16) places before the
// positive.  One warp per user, coalesced row reads.
__global__ void __launch_bounds__(256) catalog_rank_count_kernel(int* __restrict__ count, const float* __restrict__ y,
                                                                 long long ldy, const float* __restrict__ y_pos,
                                                                 const int* __restrict__ pos_item, int item_lo, int B,
                                                                 int n) {
  const int b = blockIdx.x * (blockDim.x / kWarp) + threadIdx.x / kWarp;
  if (b >= B) return;
  const int lane = threadIdx.x % kWarp;
  const float yp = y_pos[b];
  const int pi = pos_item[b];
  const float* row = y + (long long)b * ldy;
  int c = 0;
  for (int j = lane; j < n; j += kWarp) {
    const float v = row[j];
    const int item = item_lo + j;
    c += item != pi && ((v > yp) || (v == yp && item < pi));   // the positive never outranks itself
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(kFull, c, o);
  if (lane == 0 && c) atomicAdd(count + b, c);
}

// evaluate()'s per-batch reductions in ONE launch (src/train.py:44-50): masked BCE of the batch
// (src/carca.py:441-444 with mask = get_mask(o_x), src/utils.py:6-7) and HR@k / NDCG@k (src/train.py:15-32, torch's
// stable tie order) — y is read once, five small launches (mask, BCE sums, finalize, rank metrics, accumulator add)
// become one.  One warp per row.  stats += [hits, sum 1/log2(rank+2), B, sum(l mask) / sum(mask)];  work (3 doubles,
// zero before the first launch) holds the two BCE partial sums and a block counter; the last block to finish forms
// the batch loss and clears it again, so replays (CUDA graphs) need no memset.
__global__ void __launch_bounds__(256) eval_metrics_kernel(double* __restrict__ stats, double* __restrict__ work,
                                                           const float* __restrict__ y, const int* __restrict__ yt,
                                                           const int* __restrict__ ox, int B, int T, long long ldy,
                                                           long long ldt, long long ldx, int k, float eps) {
  __shared__ double part[4][8];
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  const int warps = blockDim.x / kWarp;
  double hits = 0.0, ndcg = 0.0;
  float ls = 0.f, ms = 0.f;
  for (int b = blockIdx.x * warps + w; b < B; b += gridDim.x * warps) {
    const float* yr = y + (long long)b * ldy;
    const int* tr = yt + (long long)b * ldt;
    const int* xr = ox + (long long)b * ldx;
    for (int j0 = 0; j0 < T; j0 += kWarp) {
      const int jj = j0 + lane;
      const int lab = jj < T ? tr[jj] : 0;
      if (jj < T && xr[jj] != 0) {
        const float tv = (float)lab, yv = yr[jj];
        ls += -(tv * logf(yv + eps) + (1.0f - tv) * logf(1.0f - yv + eps));
        ms += 1.0f;
      }
      unsigned pos = __ballot_sync(kFull, lab != 0);
      while (pos) {
        const int l = __ffs((int)pos) - 1;
        pos &= pos - 1;
        const int j = j0 + l;
        const float sj = yr[j];
        const int labj = __shfl_sync(kFull, lab, l);
        int cnt = 0;
        for (int i = lane; i < T; i += kWarp) {
          const float si = yr[i];
          cnt += (si > sj) || (si == sj && i < j);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(kFull, cnt, o);
        if (cnt < k) {
          hits += (double)labj;
          ndcg += (double)(1.0f / log2f((float)(cnt + 2)));
        }
      }
    }
  }
  ls = warp_sum(ls);
  ms = warp_sum(ms);
  if (lane == 0) { part[0][w] = hits; part[1][w] = ndcg; part[2][w] = (double)ls; part[3][w] = (double)ms; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double h = 0.0, n = 0.0, a = 0.0, m = 0.0;
    for (int i = 0; i < warps; ++i) { h += part[0][i]; n += part[1][i]; a += part[2][i]; m += part[3][i]; }
    atomicAdd(stats + 0, h);
    atomicAdd(stats + 1, n);
    atomicAdd(work + 0, a);
    atomicAdd(work + 1, m);
    __threadfence();
    const double done = atomicAdd(work + 2, 1.0) + 1.0;
    if (done == (double)gridDim.x) {   // last block: every partial sum above is visible (fence + atomics at L2)
      __threadfence();
      const double num = atomicAdd(work + 0, 0.0), den = atomicAdd(work + 1, 0.0);
      atomicAdd(stats + 2, (double)B);
      atomicAdd(stats + 3, (double)((float)num / (float)den));
      work[0] = 0.0; work[1] = 0.0; work[2] = 0.0;
    }
  }
}

}  // namespace carca
