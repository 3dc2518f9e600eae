// Masked multi-head attention core of MultiHeadAttention.forward (src/carca.py:242-260),
// i.e. everything after the Q/K/V projections:
//   allow[i,j] = q_mask[i] && k_mask[j] && (no causal || j - i <= diag)
//   S = (allow ? 0 : -2^32+1  +  Q K^T) / sqrt(dh);  W = softmax_j(S) * allow;  O = dropout(W) V
// Heads are column slices [h*dh, (h+1)*dh) of the [B, L, d] projections (the reference's
// split/cat copies, :242-244/:260, are index arithmetic here).  Because the fill value dwarfs any
// logit, softmax over a row with at least one allowed key equals softmax over the allowed keys
// only, and a row with none is exactly zero after the `* allow` (:256): that is what is computed.
//
// One CTA per (user, head) [x query chunk]; K_h and V_h live in shared memory; one warp per
// query row; logits/softmax by warp shuffle.  The backward recomputes the logits from Q/K
// (flash-style, no [L,L] tensor saved) in two sweeps: rows -> dQ, keys -> dK, dV.
#pragma once
#include "common.cuh"

namespace carca {

struct AttnArgs {
  const float* Q;       // [B, Lq, ldq] (head h at column h*dh)
  const float* K;       // [B, Lk, ldk]
  const float* V;       // [B, Lk, ldk]
  float* O;             // [B, Lq, ldo]
  const float* resid;   // optional [B, Lq, ldo]: O = attention + resid (src/carca.py:302,343)
  float* W;             // optional [B, H, Lq, Lk] post-mask, pre-dropout weights (return_w)
  const float* q_mask;  // [B, Lq]
  const float* k_mask;  // [B, Lk]
  int B, H, Lq, Lk, dh;
  long long ldq, ldk, ldo;
  int causal;           // 0: none, 1: tril(diagonal=diag)
  int diag;
  float inv_div;        // unused (division by sqrt(dh) is done as a division, like the reference)
  float sqrt_dh;
  DropCfg drop;
  // backward only
  const float* dO;      // [B, Lq, ldo]
  float* dQ;            // [B, Lq, ldq]
  float* dK;            // [B, Lk, ldk]
  float* dV;            // [B, Lk, ldk]
  int rows_per_cta;
};

constexpr int kAttnWarps = 8;
constexpr int kAttnMaxChunks = 8;  // Lk <= 256

__device__ __forceinline__ bool attn_allowed(const AttnArgs& a, float qm, float km, int i, int j) {
  return qm != 0.f && km != 0.f && (!a.causal || (j - i) <= a.diag);
}

// smem: Ks[Lk][dh+1], Vs[Lk][dh+1], km[Lk], per warp: qrow[dh], prow[Lk]
__global__ void __launch_bounds__(kAttnWarps * 32) attention_fwd_kernel(const AttnArgs a) {
  CARCA_DYN_SMEM(float, sm);
  const int dh = a.dh, Lk = a.Lk, Lq = a.Lq, ldh = dh + 1;
  float* Ks = sm;
  float* Vs = Ks + (long long)Lk * ldh;
  float* kms = Vs + (long long)Lk * ldh;
  float* wbuf = kms + Lk;
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* qrow = wbuf + (long long)w * (dh + Lk);
  float* prow = qrow + dh;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;
  const int row_begin = blockIdx.y * a.rows_per_cta;
  const int row_end = min(Lq, row_begin + a.rows_per_cta);

  const float* Kg = a.K + (long long)b * Lk * a.ldk + h * dh;
  const float* Vg = a.V + (long long)b * Lk * a.ldk + h * dh;
  for (int e = threadIdx.x; e < Lk * dh; e += blockDim.x) {
    const int j = e / dh, c = e % dh;
    Ks[j * ldh + c] = Kg[(long long)j * a.ldk + c];
    Vs[j * ldh + c] = Vg[(long long)j * a.ldk + c];
  }
  for (int j = threadIdx.x; j < Lk; j += blockDim.x) kms[j] = a.k_mask[(long long)b * Lk + j];
  __syncthreads();

  for (int i = row_begin + w; i < row_end; i += kAttnWarps) {
    const float qm = a.q_mask[(long long)b * Lq + i];
    const float* Qg = a.Q + ((long long)b * Lq + i) * a.ldq + h * dh;
    __syncwarp();
    for (int c = lane; c < dh; c += kWarp) qrow[c] = Qg[c];
    __syncwarp();
    float s[kAttnMaxChunks];
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const int j = ch * kWarp + lane;
      s[ch] = -INFINITY;
      if (ch * kWarp < Lk && j < Lk && attn_allowed(a, qm, kms[j], i, j)) {
        float acc = 0.f;
        for (int c = 0; c < dh; ++c) acc = fmaf(qrow[c], Ks[j * ldh + c], acc);
        s[ch] = acc / a.sqrt_dh;
        mx = fmaxf(mx, s[ch]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const float e = (s[ch] == -INFINITY) ? 0.f : expf(s[ch] - mx);
      s[ch] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const int j = ch * kWarp + lane;
      if (ch * kWarp < Lk && j < Lk) {
        const float p = s[ch] * inv;
        if (a.W) a.W[(((long long)b * a.H + h) * Lq + i) * Lk + j] = p;
        const unsigned long long elem = (((unsigned long long)b * a.H + h) * Lq + i) * Lk + j;
        prow[j] = (p != 0.f) ? p * drop_factor(a.drop, elem) : 0.f;
      }
    }
    __syncwarp();
    float* Og = a.O + ((long long)b * Lq + i) * a.ldo + h * dh;
    for (int c = lane; c < dh; c += kWarp) {
      float acc = 0.f;
      for (int j = 0; j < Lk; ++j) acc = fmaf(prow[j], Vs[j * ldh + c], acc);
      if (a.resid) acc += a.resid[((long long)b * Lq + i) * a.ldo + h * dh + c];
      Og[c] = acc;
    }
  }
}

// smem: Qs[Lq][dh+1], dOs[Lq][dh+1], Ks[Lk][dh+1], Vs[Lk][dh+1], qm[Lq], km[Lk],
//       rmax[Lq], rinv[Lq], rD[Lq], per warp: buf1[max(Lq,Lk)], buf2[max(Lq,Lk)]
__global__ void __launch_bounds__(kAttnWarps * 32) attention_bwd_kernel(const AttnArgs a) {
  CARCA_DYN_SMEM(float, sm);
  const int dh = a.dh, Lk = a.Lk, Lq = a.Lq, ldh = dh + 1;
  const int Lm = max(Lq, Lk);
  float* Qs = sm;
  float* dOs = Qs + (long long)Lq * ldh;
  float* Ks = dOs + (long long)Lq * ldh;
  float* Vs = Ks + (long long)Lk * ldh;
  float* qms = Vs + (long long)Lk * ldh;
  float* kms = qms + Lq;
  float* rmax = kms + Lk;
  float* rinv = rmax + Lq;
  float* rD = rinv + Lq;
  float* wbuf = rD + Lq;
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* buf1 = wbuf + (long long)w * 2 * Lm;
  float* buf2 = buf1 + Lm;
  const int b = blockIdx.x / a.H, h = blockIdx.x % a.H;

  const float* Qg = a.Q + (long long)b * Lq * a.ldq + h * dh;
  const float* dOg = a.dO + (long long)b * Lq * a.ldo + h * dh;
  const float* Kg = a.K + (long long)b * Lk * a.ldk + h * dh;
  const float* Vg = a.V + (long long)b * Lk * a.ldk + h * dh;
  for (int e = threadIdx.x; e < Lq * dh; e += blockDim.x) {
    const int i = e / dh, c = e % dh;
    Qs[i * ldh + c] = Qg[(long long)i * a.ldq + c];
    dOs[i * ldh + c] = dOg[(long long)i * a.ldo + c];
  }
  for (int e = threadIdx.x; e < Lk * dh; e += blockDim.x) {
    const int j = e / dh, c = e % dh;
    Ks[j * ldh + c] = Kg[(long long)j * a.ldk + c];
    Vs[j * ldh + c] = Vg[(long long)j * a.ldk + c];
  }
  for (int i = threadIdx.x; i < Lq; i += blockDim.x) qms[i] = a.q_mask[(long long)b * Lq + i];
  for (int j = threadIdx.x; j < Lk; j += blockDim.x) kms[j] = a.k_mask[(long long)b * Lk + j];
  __syncthreads();

  // ---- sweep 1: one warp per query row -> softmax stats, D_i = sum_j p_ij dP_ij, dQ
  for (int i = w; i < Lq; i += kAttnWarps) {
    const float qm = qms[i];
    float s[kAttnMaxChunks], dp[kAttnMaxChunks];
    float mx = -INFINITY;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const int j = ch * kWarp + lane;
      s[ch] = -INFINITY;
      dp[ch] = 0.f;
      if (ch * kWarp < Lk && j < Lk && attn_allowed(a, qm, kms[j], i, j)) {
        float acc = 0.f, acc2 = 0.f;
        for (int c = 0; c < dh; ++c) {
          acc = fmaf(Qs[i * ldh + c], Ks[j * ldh + c], acc);
          acc2 = fmaf(dOs[i * ldh + c], Vs[j * ldh + c], acc2);
        }
        s[ch] = acc / a.sqrt_dh;
        const unsigned long long elem = (((unsigned long long)b * a.H + h) * Lq + i) * Lk + j;
        dp[ch] = acc2 * drop_factor(a.drop, elem);
        mx = fmaxf(mx, s[ch]);
      }
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const float e = (s[ch] == -INFINITY) ? 0.f : expf(s[ch] - mx);
      s[ch] = e;
      sum += e;
    }
    sum = warp_sum(sum);
    const float inv = sum > 0.f ? 1.0f / sum : 0.f;
    float D = 0.f;
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      s[ch] *= inv;
      D = fmaf(s[ch], dp[ch], D);
    }
    D = warp_sum(D);
    __syncwarp();
#pragma unroll
    for (int ch = 0; ch < kAttnMaxChunks; ++ch) {
      const int j = ch * kWarp + lane;
      if (ch * kWarp < Lk && j < Lk) buf1[j] = s[ch] * (dp[ch] - D) / a.sqrt_dh;   // d(QK^T)_ij
    }
    __syncwarp();
    float* dQg = a.dQ + ((long long)b * Lq + i) * a.ldq + h * dh;
    for (int c = lane; c < dh; c += kWarp) {
      float acc = 0.f;
      for (int j = 0; j < Lk; ++j) acc = fmaf(buf1[j], Ks[j * ldh + c], acc);
      dQg[c] = acc;
    }
    if (lane == 0) {
      rmax[i] = mx;
      rinv[i] = inv;
      rD[i] = D;
    }
  }
  __syncthreads();

  // ---- sweep 2: one warp per key j -> dK_j = sum_i dS_ij Q_i, dV_j = sum_i dropout(W)_ij dO_i
  for (int j = w; j < Lk; j += kAttnWarps) {
    const float km = kms[j];
    __syncwarp();
    for (int i0 = 0; i0 < Lq; i0 += kWarp) {
      const int i = i0 + lane;
      if (i < Lq) {
        float ds = 0.f, pd = 0.f;
        if (attn_allowed(a, qms[i], km, i, j)) {
          float acc = 0.f, acc2 = 0.f;
          for (int c = 0; c < dh; ++c) {
            acc = fmaf(Qs[i * ldh + c], Ks[j * ldh + c], acc);
            acc2 = fmaf(dOs[i * ldh + c], Vs[j * ldh + c], acc2);
          }
          const float p = expf(acc / a.sqrt_dh - rmax[i]) * rinv[i];
          const unsigned long long elem = (((unsigned long long)b * a.H + h) * Lq + i) * Lk + j;
          const float f = drop_factor(a.drop, elem);
          ds = p * (acc2 * f - rD[i]) / a.sqrt_dh;
          pd = p * f;
        }
        buf1[i] = ds;
        buf2[i] = pd;
      }
    }
    __syncwarp();
    float* dKg = a.dK + ((long long)b * Lk + j) * a.ldk + h * dh;
    float* dVg = a.dV + ((long long)b * Lk + j) * a.ldk + h * dh;
    for (int c = lane; c < dh; c += kWarp) {
      float ak = 0.f, av = 0.f;
      for (int i = 0; i < Lq; ++i) {
        ak = fmaf(buf1[i], Qs[i * ldh + c], ak);
        av = fmaf(buf2[i], dOs[i * ldh + c], av);
      }
      dKg[c] = ak;
      dVg[c] = av;
    }
  }
}

inline size_t attention_fwd_smem(int Lk, int dh) {
  return sizeof(float) * ((size_t)2 * Lk * (dh + 1) + Lk + (size_t)kAttnWarps * (dh + Lk));
}
inline size_t attention_bwd_smem(int Lq, int Lk, int dh) {
  const int Lm = Lq > Lk ? Lq : Lk;
  return sizeof(float) * ((size_t)2 * Lq * (dh + 1) + (size_t)2 * Lk * (dh + 1) + 4 * (size_t)Lq + Lk +
                          (size_t)kAttnWarps * 2 * Lm);
}

}  // namespace carca
