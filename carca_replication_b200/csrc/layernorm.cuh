// LayerNorm over the feature dim (nn.LayerNorm(d), eps 1e-5, affine; src/carca.py:279,283,408).
// One warp per row, statistics by warp shuffle, two-pass variance (biased, as torch).
#pragma once
#include "common.cuh"

namespace carca {

__global__ void __launch_bounds__(256) layernorm_fwd_kernel(float* __restrict__ y, float* __restrict__ mean_out,
                                                            float* __restrict__ rstd_out,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ gamma,
                                                            const float* __restrict__ beta, int rows, int d) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) / kWarp;
  const int lane = threadIdx.x % kWarp;
  if (row >= rows) return;
  const float* xr = x + (long long)row * d;
  float s = 0.f;
  for (int j = lane; j < d; j += kWarp) s += xr[j];
  const float mean = warp_sum(s) / (float)d;
  float v = 0.f;
  for (int j = lane; j < d; j += kWarp) {
    const float c = xr[j] - mean;
    v = fmaf(c, c, v);
  }
  const float rstd = 1.0f / sqrtf(warp_sum(v) / (float)d + kLnEps);
  float* yr = y + (long long)row * d;
  for (int j = lane; j < d; j += kWarp) yr[j] = (xr[j] - mean) * rstd * gamma[j] + beta[j];
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
}

// dx (=|+=) rstd * (g - mean(g) - xhat * mean(g*xhat)),  g = dy*gamma;
// dgamma[j] += sum_rows dy*xhat, dbeta[j] += sum_rows dy   (block-level partials, then atomics).
// Dynamic smem: 2 * warps_per_block * d floats.
__global__ void __launch_bounds__(256) layernorm_bwd_kernel(float* __restrict__ dx, float* __restrict__ dgamma,
                                                            float* __restrict__ dbeta,
                                                            const float* __restrict__ dy,
                                                            const float* __restrict__ x,
                                                            const float* __restrict__ mean_in,
                                                            const float* __restrict__ rstd_in,
                                                            const float* __restrict__ gamma, int rows, int d,
                                                            int accumulate_dx) {
  CARCA_DYN_SMEM(float, sm);
  const int warps = blockDim.x / kWarp;
  const int w = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  float* sg = sm + (long long)w * d;              // per-warp dgamma partial
  float* sb = sm + (long long)(warps + w) * d;    // per-warp dbeta partial
  for (int j = lane; j < d; j += kWarp) { sg[j] = 0.f; sb[j] = 0.f; }
  for (int row = blockIdx.x * warps + w; row < rows; row += gridDim.x * warps) {
    const float* xr = x + (long long)row * d;
    const float* dyr = dy + (long long)row * d;
    const float mean = mean_in[row], rstd = rstd_in[row];
    float c1 = 0.f, c2 = 0.f;
    for (int j = lane; j < d; j += kWarp) {
      const float xh = (xr[j] - mean) * rstd;
      const float gj = dyr[j] * gamma[j];
      c1 += gj;
      c2 = fmaf(gj, xh, c2);
      sg[j] = fmaf(dyr[j], xh, sg[j]);
      sb[j] += dyr[j];
    }
    c1 = warp_sum(c1) / (float)d;
    c2 = warp_sum(c2) / (float)d;
    float* dxr = dx + (long long)row * d;
    for (int j = lane; j < d; j += kWarp) {
      const float xh = (xr[j] - mean) * rstd;
      const float val = rstd * (dyr[j] * gamma[j] - c1 - xh * c2);
      dxr[j] = accumulate_dx ? dxr[j] + val : val;
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < d; j += blockDim.x) {
    float a = 0.f, b = 0.f;
    for (int ww = 0; ww < warps; ++ww) {
      a += sm[(long long)ww * d + j];
      b += sm[(long long)(warps + ww) * d + j];
    }
    atomicAdd(dgamma + j, a);
    atomicAdd(dbeta + j, b);
  }
}

}  // namespace carca
