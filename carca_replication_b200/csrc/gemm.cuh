// General fp32 GEMM used by the projection / weight-gradient paths of the CARCA kernels.
//
//   C[m,n] (=|+=)  rowmask[m] * ( drop( act( alpha * sum_k A'(m,k) B'(k,n) + bias[n] (+ C_old) ) ) + R[m % r_mod, n] )
//
//   A'(m,k) = transA ? A[k*lda + m]            : A[a_rows(m)*lda + k]     (a_rows: optional row gather)
//   B'(k,n) = transB ? B[n*ldb + k]            : B[b_rows(k)*ldb + n]     (b_rows: optional row gather)
//
// Three uses (reference op in brackets):
//   y = x W^T + b      transA=0 transB=1   [nn.Linear / Conv1d(k=1) forward, src/carca.py:86,89,238-240,307,311]
//   dx = dy W          transA=0 transB=0   [autograd of the above]
//   dW = dy^T x        transA=1 transB=0   split over k (= positions) with atomic accumulation
// Row gathers let the item-embedding / attribute tables be read in place (no [P, d] copy of E[x]).
#pragma once
#include "common.cuh"

namespace carca {

struct GemmArgs {
  const float* A;
  const float* B;
  float* C;
  int M, N, K;
  long long lda, ldb, ldc;
  int transA, transB;
  const int* a_rows;      // !transA: A row of output row m is a_rows[m]
  const int* b_rows;      // !transB: B row of reduction index k is b_rows[k]
  float alpha;
  const float* bias;      // [N] or null
  int accumulate;         // 1: add to what C holds (atomically when split-K)
  int act;                // 0 none, 1 LeakyReLU(0.01)
  DropCfg drop;           // element index m*N + n
  const float* R;         // residual, row (m % r_mod if r_mod else m), ldr
  long long ldr;
  int r_mod;
  const float* row_mask;  // [M] or null; multiplies the final value
  int k_per_split;        // split-K chunk (multiple of BK); gridDim.z chunks
  const int* m_dev;       // optional device word: only the first min(M, *m_dev) rows exist (row counts known on the
                          //   device only, e.g. the packed rows of a batch); the grid is sized for M, tiles past it exit
};

template <int BM, int BN, int TM, int TN>
__global__ void __launch_bounds__(256) gemm_kernel(const GemmArgs g) {
  constexpr int BK = 16;
  constexpr int PAD = 4;
  static_assert((BM / TM) * (BN / TN) == 256, "256 threads");
  static_assert(TM % 4 == 0 && TN % 4 == 0, "float4 smem reads");
  __shared__ __align__(16) float As[BK][BM + PAD];
  __shared__ __align__(16) float Bs[BK][BN + PAD];
  constexpr int A_PER = BM * BK / 256;
  constexpr int B_PER = BN * BK / 256;

  const int tid = threadIdx.x;
  const int m0 = blockIdx.y * BM;
  const int gM = g.m_dev ? min(g.M, *g.m_dev) : g.M;
  if (m0 >= gM) return;
  const int n0 = blockIdx.x * BN;
  const int k_begin = blockIdx.z * g.k_per_split;
  const int k_end = min(g.K, k_begin + g.k_per_split);
  const int tx = tid % (BN / TN);
  const int ty = tid / (BN / TN);

  float acc[TM][TN];
#pragma unroll
  for (int i = 0; i < TM; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  float ra[A_PER], rb[B_PER];

  auto fetch = [&](int k0) {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (g.transA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      const int gm = m0 + m, gk = k0 + k;
      float v = 0.f;
      if (gm < gM && gk < k_end) {
        if (g.transA) {
          v = g.A[(long long)gk * g.lda + gm];
        } else {
          const long long row = g.a_rows ? (long long)g.a_rows[gm] : (long long)gm;
          v = g.A[row * g.lda + gk];
        }
      }
      ra[i] = v;
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (g.transB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      const int gn = n0 + n, gk = k0 + k;
      float v = 0.f;
      if (gn < g.N && gk < k_end) {
        if (g.transB) {
          v = g.B[(long long)gn * g.ldb + gk];
        } else {
          const long long row = g.b_rows ? (long long)g.b_rows[gk] : (long long)gk;
          v = g.B[row * g.ldb + gn];
        }
      }
      rb[i] = v;
    }
  };
  auto stash = [&]() {
#pragma unroll
    for (int i = 0; i < A_PER; ++i) {
      const int e = tid + i * 256;
      int m, k;
      if (g.transA) { k = e / BM; m = e % BM; } else { m = e / BK; k = e % BK; }
      As[k][m] = ra[i];
    }
#pragma unroll
    for (int i = 0; i < B_PER; ++i) {
      const int e = tid + i * 256;
      int n, k;
      if (g.transB) { n = e / BK; k = e % BK; } else { k = e / BN; n = e % BN; }
      Bs[k][n] = rb[i];
    }
  };

  if (k_begin < k_end) {
    fetch(k_begin);
    stash();
    __syncthreads();
    for (int k0 = k_begin; k0 < k_end; k0 += BK) {
      const bool more = (k0 + BK) < k_end;
      if (more) fetch(k0 + BK);
#pragma unroll
      for (int kk = 0; kk < BK; ++kk) {
        float a[TM], b[TN];
#pragma unroll
        for (int i = 0; i < TM / 4; ++i) {
          const float4 v = *reinterpret_cast<const float4*>(&As[kk][ty * TM + i * 4]);
          a[i * 4 + 0] = v.x; a[i * 4 + 1] = v.y; a[i * 4 + 2] = v.z; a[i * 4 + 3] = v.w;
        }
#pragma unroll
        for (int j = 0; j < TN / 4; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(&Bs[kk][tx * TN + j * 4]);
          b[j * 4 + 0] = v.x; b[j * 4 + 1] = v.y; b[j * 4 + 2] = v.z; b[j * 4 + 3] = v.w;
        }
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
          for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
      }
      __syncthreads();
      if (more) {
        stash();
        __syncthreads();
      }
    }
  }

  const bool split = gridDim.z > 1;
#pragma unroll
  for (int i = 0; i < TM; ++i) {
    const int gm = m0 + ty * TM + i;
    if (gm >= gM) continue;
    const float rm = g.row_mask ? g.row_mask[gm] : 1.0f;
#pragma unroll
    for (int j = 0; j < TN; ++j) {
      const int gn = n0 + tx * TN + j;
      if (gn >= g.N) continue;
      float* dst = g.C + (long long)gm * g.ldc + gn;
      float v = g.alpha * acc[i][j];
      if (split) {
        atomicAdd(dst, v);
        continue;
      }
      if (g.bias) v += g.bias[gn];
      if (g.accumulate) v += *dst;
      if (g.act == 1) v = v > 0.f ? v : kLeakySlope * v;
      if (g.drop.p > 0.f) v *= drop_factor(g.drop, (unsigned long long)gm * (unsigned long long)g.N + gn);
      if (g.R) v += g.R[(long long)(g.r_mod ? gm % g.r_mod : gm) * g.ldr + gn];
      *dst = v * rm;
    }
  }
}

// Host-side launcher (defined in gemm.cu)
GemmArgs gemm_defaults(const float* A, const float* B, float* C, int M, int N, int K);
int launch_gemm(GemmArgs g, cudaStream_t stream);

}  // namespace carca
