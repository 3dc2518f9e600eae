// Kernels around AllEmbedding (src/carca.py:66-95): sparse attribute projection,
// row scatter-add for the item table gradient, column sums, transposes, row scaling.
#pragma once
#include "common.cuh"

namespace carca {

// q[p] = Wf[:, :A] a[p] + Wf[:, A:] ctx[p] + bf for DENSE attribute rows a[P, A] of a large, mostly-zero vocabulary
// (the reference API's multi-hot tensors, src/carca.py:86 with src/data.py:119-131: ~8 of 6,507 set).  One warp scans
// its row with coalesced loads (the read of `a` is the bound: 26 KB per position) and adds value x WT[j] for the
// non-zeros it finds — the feature projection as a gather-sum instead of a [P, A] x [A, g] GEMM over zeros.  Exact for
// any input (a fully dense row just costs A rank-1 updates); g <= 256.
__global__ void __launch_bounds__(256) feat_dense_scan_fwd_kernel(
    float* __restrict__ q, const float* __restrict__ a, const float* __restrict__ ctx, const float* __restrict__ WT,
    const float* __restrict__ bias, const float* __restrict__ row_mask, int P, int g, int A, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const int p = warp;
  float* out = q + (long long)p * g;
  if (row_mask && row_mask[p] == 0.f) {
    for (int j = lane; j < g; j += kWarp) out[j] = 0.f;
    return;
  }
  float acc[8];
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = u * kWarp + lane;
    acc[u] = (bias && j < g) ? bias[j] : 0.f;
  }
  for (int c = 0; c < C; ++c) {
    const float cv = ctx[(long long)p * C + c];
    const float* w = WT + (long long)(A + c) * g;
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const int j = u * kWarp + lane;
      if (j < g) acc[u] = fmaf(cv, w[j], acc[u]);
    }
  }
  const float* row = a + (long long)p * A;
  constexpr int UN = 8;
  for (int j0 = 0; j0 < A; j0 += kWarp * UN) {
    float v[UN];
#pragma unroll
    for (int t = 0; t < UN; ++t) {
      const int jj = j0 + kWarp * t + lane;
      v[t] = jj < A ? __ldg(row + jj) : 0.f;
    }
#pragma unroll
    for (int t = 0; t < UN; ++t) {
      unsigned m = __ballot_sync(kFull, v[t] != 0.f);
      while (m) {
        const int src = __ffs(m) - 1;
        m &= m - 1;
        const float av = __shfl_sync(kFull, v[t], src);
        const float* w = WT + (long long)(j0 + kWarp * t + src) * g;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          const int j = u * kWarp + lane;
          if (j < g) acc[u] = fmaf(av, __ldg(w + j), acc[u]);
        }
      }
    }
  }
#pragma unroll
  for (int u = 0; u < 8; ++u) {
    const int j = u * kWarp + lane;
    if (j < g) out[j] = acc[u];
  }
}

// q[p,:] = (bias) + sum_c ctx[p,c] * WT[A+c,:] + sum_{e in row(item_p)} vals[e] * WT[cols[e],:]
// One warp per position; lanes sweep g so every WT row read is a coalesced burst.
// Rows with row_mask[p]==0 are written as zeros (their q is never consumed: the embedding is
// multiplied by the mask, src/carca.py:94).
__global__ void __launch_bounds__(256) feat_csr_fwd_kernel(
    float* __restrict__ q, const int* __restrict__ items, const int* __restrict__ rowptr,
    const int* __restrict__ cols, const float* __restrict__ vals, const float* __restrict__ ctx,
    const float* __restrict__ WT, const float* __restrict__ bias, const float* __restrict__ row_mask,
    int P, int g, int A, int C) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const int p = warp;
  float* out = q + (long long)p * g;
  if (row_mask && row_mask[p] == 0.f) {
    for (int j = lane; j < g; j += kWarp) out[j] = 0.f;
    return;
  }
  const int item = items ? items[p] : p;   // items == null: position p IS item p (table folding)
  const int e0 = rowptr[item], e1 = rowptr[item + 1];
  for (int j0 = 0; j0 < g; j0 += 4 * kWarp) {
    float acc[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * kWarp + lane;
      acc[u] = (bias && j < g) ? bias[j] : 0.f;
    }
    for (int c = 0; c < C; ++c) {
      const float cv = ctx[(long long)p * C + c];
      const float* w = WT + (long long)(A + c) * g;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kWarp + lane;
        if (j < g) acc[u] = fmaf(cv, w[j], acc[u]);
      }
    }
    for (int e = e0; e < e1; ++e) {
      const float av = vals[e];
      const float* w = WT + (long long)cols[e] * g;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int j = j0 + u * kWarp + lane;
        if (j < g) acc[u] = fmaf(av, w[j], acc[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int j = j0 + u * kWarp + lane;
      if (j < g) out[j] = acc[u];
    }
  }
}

// dWT[cols[e],:] += vals[e] * dq[p,:] for every stored attribute of the item at position p.
__global__ void __launch_bounds__(256) feat_csr_bwd_kernel(
    float* __restrict__ dWT, const float* __restrict__ dq, const int* __restrict__ items,
    const int* __restrict__ rowptr, const int* __restrict__ cols, const float* __restrict__ vals,
    const float* __restrict__ row_mask, int P, int g) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const int p = warp;
  if (row_mask && row_mask[p] == 0.f) return;
  const int item = items[p];
  const int e0 = rowptr[item], e1 = rowptr[item + 1];
  const float* src = dq + (long long)p * g;
  for (int e = e0; e < e1; ++e) {
    const float av = vals[e];
    float* dst = dWT + (long long)cols[e] * g;
    for (int j = lane; j < g; j += kWarp) atomicAdd(dst + j, av * src[j]);
  }
}

// table[idx[p],:] += src[p,:]   (skips idx 0: padding_idx row keeps a zero gradient, src/carca.py:73)
__global__ void __launch_bounds__(256) scatter_add_rows_kernel(float* __restrict__ table,
                                                               const float* __restrict__ src,
                                                               const int* __restrict__ idx, int P, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const int row = idx[warp];
  if (row == 0) return;
  const float* s = src + (long long)warp * d;
  float* t = table + (long long)row * d;
  for (int j = lane; j < d; j += kWarp) atomicAdd(t + j, s[j]);
}

// out[n] += sum_m X[m*ldx + n]
__global__ void __launch_bounds__(256) colsum_kernel(float* __restrict__ out, const float* __restrict__ X, int M,
                                                     int N, long long ldx, int rows_per_block) {
  // blockDim = (32, 8): x over columns, y over rows
  __shared__ float part[8][33];
  const int n = blockIdx.x * 32 + threadIdx.x;
  const int m_begin = blockIdx.y * rows_per_block;
  const int m_end = min(M, m_begin + rows_per_block);
  float acc = 0.f;
  if (n < N)
    for (int m = m_begin + threadIdx.y; m < m_end; m += 8) acc += X[(long long)m * ldx + n];
  part[threadIdx.y][threadIdx.x] = acc;
  __syncthreads();
  if (threadIdx.y == 0 && n < N) {
    float s = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) s += part[r][threadIdx.x];
    atomicAdd(out + n, s);
  }
}

// dst[c*ld_dst + r] (=|+=) src[r*ld_src + c]
__global__ void __launch_bounds__(256) transpose_kernel(float* __restrict__ dst, const float* __restrict__ src,
                                                        int R, int Cc, long long ld_src, long long ld_dst,
                                                        int accumulate) {
  __shared__ float tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + threadIdx.x;
    tile[i][threadIdx.x] = (r < R && c < Cc) ? src[(long long)r * ld_src + c] : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + threadIdx.x;
    if (r < R && c < Cc) {
      float* p = dst + (long long)c * ld_dst + r;
      const float v = tile[threadIdx.x][i];
      *p = accumulate ? (*p + v) : v;
    }
  }
}

// Y[m,n] = X[m,n] * rs[m] * dropfactor(m*N+n)     (rs / drop optional)
__global__ void __launch_bounds__(256) scale_rows_kernel(float* __restrict__ Y, const float* __restrict__ X,
                                                         const float* __restrict__ rs, DropCfg drop,
                                                         long long total, int N) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  float v = X[i];
  if (rs) v *= rs[i / N];
  if (drop.p > 0.f) v *= drop_factor(drop, (unsigned long long)i);
  Y[i] = v;
}

// mask[i] = ids[i] != 0 ? 1 : 0     (get_mask, src/utils.py:6-7)
__global__ void __launch_bounds__(256) padding_mask_kernel(float* __restrict__ mask, const int* __restrict__ ids,
                                                           long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = ids[i] != 0 ? 1.0f : 0.0f;
}
__global__ void __launch_bounds__(256) padding_mask_f32_kernel(float* __restrict__ mask, const float* __restrict__ x,
                                                               long long n) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) mask[i] = x[i] == 0.0f ? 0.0f : 1.0f;
}

}  // namespace carca
