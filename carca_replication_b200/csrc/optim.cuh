// Adam over every parameter tensor of the model in ONE launch.
// Reference: torch.optim.Adam as scripts/training.py:174 builds it (lr 1e-3, betas (0.9, 0.98), weight_decay =
// l2_reg added to the gradient, no amsgrad) and src/train.py:96 steps it.  torch's own implementation is ~9
// multi-tensor launches per step; at the reference batch size that is a fifth of the whole training step.
// HBM-bound: 4 reads + 3 writes of 4 bytes per parameter; 4096-element chunks, 128-bit accesses.
#pragma once
#include "common.cuh"

namespace carca {

constexpr int kAdamMaxTensors = 64;
constexpr int kAdamChunk = 4096;

struct AdamArgs {
  float* p[kAdamMaxTensors];
  const float* g[kAdamMaxTensors];
  float* m[kAdamMaxTensors];
  float* v[kAdamMaxTensors];
  long long n[kAdamMaxTensors];
  int chunk0[kAdamMaxTensors + 1];   // first chunk of tensor i; chunk0[n_tensors] = total
  int n_tensors;
  float lr, beta1, beta2, eps, weight_decay;
  float* step[kAdamMaxTensors];      // per-tensor device counters (torch keeps one step count per parameter)
};

__global__ void adam_tick_kernel(const AdamArgs a) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < a.n_tensors) a.step[i][0] += 1.0f;
}

__global__ void __launch_bounds__(256) adam_step_kernel(const AdamArgs a) {
  int t = 0;
  while (t + 1 < a.n_tensors && (int)blockIdx.x >= a.chunk0[t + 1]) ++t;
  const long long base = (long long)((int)blockIdx.x - a.chunk0[t]) * kAdamChunk;
  const long long n = a.n[t];
  float* __restrict__ p = a.p[t];
  const float* __restrict__ g = a.g[t];
  float* __restrict__ m = a.m[t];
  float* __restrict__ v = a.v[t];
  const float step = a.step[t][0];
  const float bc1 = 1.0f - powf(a.beta1, step);
  const float bc2_sqrt = sqrtf(1.0f - powf(a.beta2, step));
  const float step_size = a.lr / bc1;
  const bool vec = ((((uintptr_t)p | (uintptr_t)g | (uintptr_t)m | (uintptr_t)v) & 15) == 0);
  auto one = [&](float& pv, float gv, float& mv, float& vv) {
    if (a.weight_decay != 0.f) gv = fmaf(a.weight_decay, pv, gv);
    mv = mv + (gv - mv) * (1.0f - a.beta1);
    vv = a.beta2 * vv + (1.0f - a.beta2) * gv * gv;
    pv -= step_size * mv / (sqrtf(vv) / bc2_sqrt + a.eps);
  };
  for (int e = threadIdx.x * 4; e < kAdamChunk; e += 256 * 4) {
    const long long i = base + e;
    if (i >= n) break;
    if (vec && i + 4 <= n) {
      float4 pv = *reinterpret_cast<float4*>(p + i);
      const float4 gv = *reinterpret_cast<const float4*>(g + i);
      float4 mv = *reinterpret_cast<float4*>(m + i);
      float4 vv = *reinterpret_cast<float4*>(v + i);
      one(pv.x, gv.x, mv.x, vv.x);
      one(pv.y, gv.y, mv.y, vv.y);
      one(pv.z, gv.z, mv.z, vv.z);
      one(pv.w, gv.w, mv.w, vv.w);
      *reinterpret_cast<float4*>(p + i) = pv;
      *reinterpret_cast<float4*>(m + i) = mv;
      *reinterpret_cast<float4*>(v + i) = vv;
    } else {
      for (long long j = i; j < n && j < i + 4; ++j) one(p[j], g[j], m[j], v[j]);
    }
  }
}

}  // namespace carca
