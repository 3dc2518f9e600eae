// Kernels of the non-default module variants: the id-only / attribute-only embeddings
// (src/carca.py:98-198), WeightedDotProduct (:368-395) and the KNN baseline (src/knn.py:8-21).
#pragma once
#include "common.cuh"
#include "score.cuh"

namespace carca {

// out[p,:] = alpha * table[ids[p],:]            (nn.Embedding lookup + the sqrt(d) scale, :161-162, :187-188)
__global__ void __launch_bounds__(256) gather_rows_kernel(float* __restrict__ out, const float* __restrict__ table,
                                                          const int* __restrict__ ids, float alpha, int P, int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const float* src = table + (long long)ids[warp] * d;
  float* dst = out + (long long)warp * d;
  for (int j = lane; j < d; j += kWarp) dst[j] = alpha * src[j];
}

// table[ids[p],:] += alpha * src[p,:]           (skips id 0: padding_idx keeps a zero gradient)
__global__ void __launch_bounds__(256) scatter_add_rows_scaled_kernel(float* __restrict__ table,
                                                                      const float* __restrict__ src,
                                                                      const int* __restrict__ ids, float alpha, int P,
                                                                      int d) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp, lane = threadIdx.x % kWarp;
  if (warp >= P) return;
  const int row = ids[warp];
  if (row == 0) return;
  const float* s = src + (long long)warp * d;
  float* t = table + (long long)row * d;
  for (int j = lane; j < d; j += kWarp) atomicAdd(t + j, alpha * s[j]);
}

// out[p,:] = (in[p,:] + pos[p % n_cols,:]) * mask[p]      (enc.forward + the final mask, e.g. :116-120)
__global__ void __launch_bounds__(256) pos_mask_kernel(float* __restrict__ out, const float* __restrict__ in,
                                                       const float* __restrict__ pos, const float* __restrict__ mask,
                                                       long long total, int d, int n_cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i / d;
  float v = in[i];
  if (pos) v += pos[(p % n_cols) * d + (i % d)];
  out[i] = v * mask[p];
}

// AllEmbedding.forward (src/carca.py:85-95) in inference through a folded item table: the two linears have no
// nonlinearity between them, so  e = Wj [sqrt(d) E[x] | Wf [a | c] + bf] + bj = T[x] + Mc c  with
// T[i] = Wj [sqrt(d) E[i] | Wf_a attrs[i] + bf] + bj  (one row per item, built once per weight version by running
// the unfolded op over all item ids with a zero context) and Mc = Wj[:, d:] Wf[:, A:]  ([d, C], passed transposed).
//   out[p, :] = mask[p] * (T[x[p], :] + sum_k c[p, k] McT[k, :] + pos[p % n_cols, :])
// One thread per 4 features; HBM-bound gather (d * 4 bytes per position instead of a K = A + C and a K = g + d product).
__global__ void __launch_bounds__(256) embed_folded_kernel(float* __restrict__ out, const float* __restrict__ T,
                                                           const float* __restrict__ McT, const int* __restrict__ x,
                                                           const float* __restrict__ c, const float* __restrict__ pos,
                                                           const float* __restrict__ mask, long long total4, int d,
                                                           int n_ctx, int n_cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total4) return;
  const int d4 = d / 4;
  const long long p = i / d4;
  const int f = (int)(i % d4) * 4;
  const float m = mask[p];
  const int id = x[p];
  float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
  if (m != 0.f && id != 0) {
    v = *reinterpret_cast<const float4*>(T + (long long)id * d + f);
    for (int k = 0; k < n_ctx; ++k) {
      const float ck = c[p * n_ctx + k];
      const float4 w = *reinterpret_cast<const float4*>(McT + (long long)k * d + f);
      v.x = fmaf(ck, w.x, v.x); v.y = fmaf(ck, w.y, v.y); v.z = fmaf(ck, w.z, v.z); v.w = fmaf(ck, w.w, v.w);
    }
    if (pos) {
      const float4 q = *reinterpret_cast<const float4*>(pos + (p % n_cols) * d + f);
      v.x += q.x; v.y += q.y; v.z += q.z; v.w += q.w;
    }
    v.x *= m; v.y *= m; v.z *= m; v.w *= m;
  }
  *reinterpret_cast<float4*>(out + p * d + f) = v;
}

__device__ __forceinline__ float geometric_sum(float gamma, int n_terms) {   // sum_{j<n} gamma^j as the reference's W row
  float s = 0.f;
  for (int j = 0; j < n_terms; ++j) s += powf(gamma, (float)j);
  return s;
}

// WeightedDotProduct.forward (src/carca.py:377-395): the [B,L,L,d] product there reduces to scaling profile
// position i by s_i = sum_{j<=i} gamma^j; optional L2 normalisation of both sides; sigmoid or (y+1)/2.
__global__ void __launch_bounds__(256) wdot_score_fwd_kernel(float* __restrict__ y, const float* __restrict__ p,
                                                             const float* __restrict__ o, int B, int T, int Lp, int d,
                                                             int per_position, float gamma, int normalize,
                                                             long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const int pi = per_position ? t : Lp - 1;
  const float s = geometric_sum(gamma, pi + 1);
  const float* pr = p + ((long long)b * Lp + pi) * d;
  const float* orow = o + ((long long)b * T + t) * d;
  float acc = 0.f, pp = 0.f, oo = 0.f;
  for (int j = lane; j < d; j += kWarp) {
    const float pv = s * pr[j], ov = orow[j];
    acc = fmaf(pv, ov, acc);
    pp = fmaf(pv, pv, pp);
    oo = fmaf(ov, ov, oo);
  }
  acc = warp_sum(acc);
  pp = warp_sum(pp);
  oo = warp_sum(oo);
  if (lane == 0) {
    float out;
    if (normalize) out = (acc / (fmaxf(sqrtf(pp), 1e-12f) * fmaxf(sqrtf(oo), 1e-12f)) + 1.0f) * 0.5f;
    else out = sigmoidf_(acc);
    y[(long long)b * ldy + col0 + t] = out;
  }
}

__global__ void __launch_bounds__(256) wdot_score_bwd_kernel(float* __restrict__ d_o, float* __restrict__ d_p,
                                                             const float* __restrict__ dy, const float* __restrict__ y,
                                                             const float* __restrict__ p, const float* __restrict__ o,
                                                             int B, int T, int Lp, int d, int per_position, float gamma,
                                                             int normalize, long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const int pi = per_position ? t : Lp - 1;
  const float s = geometric_sum(gamma, pi + 1);
  const float* pr = p + ((long long)b * Lp + pi) * d;
  const float* orow = o + ((long long)b * T + t) * d;
  float* dor = d_o + ((long long)b * T + t) * d;
  float* dpr = d_p + ((long long)b * Lp + pi) * d;
  const float yv = y[(long long)b * ldy + col0 + t], g = dy[(long long)b * ldy + col0 + t];
  if (!normalize) {
    const float dl = g * yv * (1.0f - yv);
    for (int j = lane; j < d; j += kWarp) {
      dor[j] = dl * s * pr[j];
      if (per_position) dpr[j] += dl * s * orow[j];
      else atomicAdd(dpr + j, dl * s * orow[j]);
    }
    return;
  }
  float pp = 0.f, oo = 0.f;
  for (int j = lane; j < d; j += kWarp) {
    const float pv = s * pr[j], ov = orow[j];
    pp = fmaf(pv, pv, pp);
    oo = fmaf(ov, ov, oo);
  }
  const float np = fmaxf(sqrtf(warp_sum(pp)), 1e-12f), no = fmaxf(sqrtf(warp_sum(oo)), 1e-12f);
  const float cosv = 2.0f * yv - 1.0f;   // <p_hat, o_hat>
  const float dl = 0.5f * g;
  for (int j = lane; j < d; j += kWarp) {
    const float ph = s * pr[j] / np, oh = orow[j] / no;
    dor[j] = dl * (ph - oh * cosv) / no;
    const float dpv = dl * (oh - ph * cosv) / np * s;
    if (per_position) dpr[j] += dpv;
    else atomicAdd(dpr + j, dpv);
  }
}

// KNN.forward (src/knn.py:14-21): y[b, col0 + t] = <p_a[b, Lp-1, :], o_a[b, t, :]>  (no sigmoid)
__global__ void __launch_bounds__(256) knn_score_kernel(float* __restrict__ y, const float* __restrict__ p_a,
                                                        const float* __restrict__ o_a, int B, int T, int Lp, int A,
                                                        long long ldy, int col0) {
  const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (warp >= (long long)B * T) return;
  const int b = (int)(warp / T), t = (int)(warp % T);
  const float* pr = p_a + ((long long)b * Lp + (Lp - 1)) * A;
  const float* orow = o_a + ((long long)b * T + t) * A;
  float acc = 0.f;
  for (int j = lane; j < A; j += kWarp) acc = fmaf(pr[j], orow[j], acc);
  acc = warp_sum(acc);
  if (lane == 0) y[(long long)b * ldy + col0 + t] = acc;
}

}  // namespace carca
