// bf16 inference pipeline over PACKED ROWS (d = 64 or 256): CARCA.forward in eval mode, src/carca.py:411-431 with
// :85-95 (embedding, folded), :228-265 (attention), :297-318 (encoder block), :338-365 (decoders).
//
// The fp32 path (fused_eval_tc.cuh) keeps a 128-row tile on one SM through the whole model; that dependent chain is
// latency-bound and needs every user to fit a 64-row bin.  This path is the opposite trade: the batch's VALID profile
// positions (plus position L-1, which the dot decoder reads) are packed, user after user, into one flat row array
// [R, d] and every stage of the model is one kernel over all R rows:
//
//   rows_pack        ids -> row list (user, position), segment start per row, (start, length) per user
//   rows_embed_ln    e = T[id] + Mc c (+ pos)           -> x (bf16 operand tiles), LN1(x) (fp32 rows + bf16 tiles)
//   (L <= 129: attention on the tensor cores, rows_attn_tc.cuh; d = 64 on top of that: the block tail FFN-1 -> FFN-2 + LN
//    -> next block's Q / K / V as ONE kernel, rows_ffn_chain.cuh — pack, embed, Q|K|V, 3 x (attention, chain), decoder)
//   per block:
//     rows_gemm x3   Q = LN1(x) Wq^T, K = x Wk^T, V = x Wv^T        (one grouped launch, tcgen05 kind::f16)
//     rows_attn      causal attention inside the row's segment + residual + LN2   (fp32 softmax / LN)
//     rows_gemm      F1 = LeakyReLU(s2 W1^T + b1)
//     rows_gemm      x' = F1 W2^T + b2 + s2, then the NEXT LayerNorm (LN1 of block b+1 or the final norm) in the epilogue
//   decoder:  K / V projections of the encoded profile (grouped rows_gemm, epilogues fold V into u = <V_h, wf_h> and
//             K into its context terms), then one kernel over (user, candidate) rows.
//
// rows_gemm is a warp-specialised tcgen05 GEMM: [R, d] x [d, d] with the WEIGHT MATRIX RESIDENT in shared memory for
// the whole kernel (128 KB at d = 256), A tiles streamed by the TMA engine (cp.async.bulk, 16 KB per 64-wide K chunk:
// activations are stored in HBM in the operand layout [tile][k/8][128 rows][8], so a chunk is one contiguous copy),
// fp32 accumulators double-buffered in TMEM (2 x d columns) so the epilogue of tile i (bias, activation, residual,
// LayerNorm, bf16 conversion: one thread per row, whole row visible in TMEM) overlaps the MMAs of tile i+1; two
// epilogue warpgroups alternate tiles.  Roles: warp 0 = bulk-copy producer, warp 1 = MMA issuer, warps 2..9 = epilogue.
// No segment or tile constraint: a user may have any number of valid positions (L <= 256).
#pragma once
#include "common.cuh"
#ifndef CARCA_EMU
#include <cuda_bf16.h>

#include "tmem_io.cuh"
#include "umma.cuh"

namespace carca {
namespace rows {

typedef __nv_bfloat16 bf16;
constexpr int TILE = 128;

// ---------------------------------------------------------------------------------------------- small helpers
__device__ __forceinline__ void unpack8(const uint4& p, float (&v)[8]) {
  const uint32_t w[4] = {p.x, p.y, p.z, p.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&h);
}
__device__ __forceinline__ uint4 pack8(const float* v) {
  return make_uint4(pack2(v[0], v[1]), pack2(v[2], v[3]), pack2(v[4], v[5]), pack2(v[6], v[7]));
}
// 256-bit global accesses (sm_100: LDG.256 / STG.256): one full 32-byte sector per thread and instruction
__device__ __forceinline__ void ldg256(const void* p, float (&v)[8]) {
  asm volatile("ld.global.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void stg256u(void* p, const uint4& a, const uint4& b) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(a.x), "r"(a.y), "r"(a.z), "r"(a.w),
               "r"(b.x), "r"(b.y), "r"(b.z), "r"(b.w)
               : "memory");
}
// element offset of (row r, k-group kg) in the operand-tiled layout [tile][D/8][128][8]
template <int D>
__device__ __forceinline__ long long tile_off(long long r, int kg) {
  return ((r >> 7) * (D / 8) + kg) * (128 * 8) + (r & 127) * 8;
}
template <int G>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
  for (int o = G / 2; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
  return v;
}
// Programmatic dependent launch (launch attribute cudaLaunchAttributeProgrammaticStreamSerialization): a kernel of the
// pipeline may start while its predecessor drains — its prologue (barrier init, TMEM allocation, parameter loads) runs
// under the predecessor's tail instead of after it.  pdl_wait() returns once every prerequisite grid has completed and
// its writes are visible: NOTHING a predecessor produced (including the device-side row count) is read before it and
// no global memory is written before it.  Both are no-ops for a kernel launched without the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// ---------------------------------------------------------------------------------------------- packing
// row_src[r] = pos | user << 8 | (id == 0) << 31;  row_seg[r] = first row of the user's segment;
// useg[u] = (first row, rows).  Segments are placed with one atomicAdd per user (order is irrelevant: users are
// independent and every row-wise stage is position-independent inside a tile).
__global__ void __launch_bounds__(256) rows_pack_kernel(int* __restrict__ row_src, int* __restrict__ row_seg,
                                                        int2* __restrict__ useg, int* __restrict__ n_rows,
                                                        const int* __restrict__ p_x, int B, int L,
                                                        int* __restrict__ row_id) {
  const int u = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (u >= B) return;
  const int* px = p_x + (long long)u * L;
  int n = 0;
  for (int p0 = 0; p0 < L; p0 += 32) {
    const int p = p0 + lane;
    const bool keep = p < L && (px[p] != 0 || p == L - 1);
    n += __popc(__ballot_sync(kFull, keep));
  }
  int base = 0;
  if (lane == 0) {
    base = atomicAdd(n_rows, n);
    useg[u] = make_int2(base, n);
  }
  base = __shfl_sync(kFull, base, 0);
  int k = 0;
  for (int p0 = 0; p0 < L; p0 += 32) {
    const int p = p0 + lane;
    const int id = p < L ? px[p] : 0;
    const bool keep = p < L && (id != 0 || p == L - 1);
    const uint32_t m = __ballot_sync(kFull, keep);
    if (keep) {
      const int r = base + k + __popc(m & ((1u << lane) - 1u));
      row_src[r] = p | (u << 8) | (id == 0 ? (int)0x80000000u : 0);
      row_seg[r] = base;
      row_id[r] = id;                 // (the embedding kernel gathers T[id] without going back to p_x)
    }
    k += __popc(m);
  }
}

// ---------------------------------------------------------------------------------------------- embedding + LN1
struct EmbedArgs {
  const float* T;           // folded item table [n_items, D] (T[i] = Wj [sqrt(d) E[i] | Wf_a attrs[i] + bf] + bj), fp32
  const float* Mc;          // folded context map [D][8]
  const float* pos;         // optional positional table [L, D]
  const int* p_x;
  const float* p_c;
  const int *row_src, *n_rows;
  const int* row_id;        // item id per row (written by the packing pass)
  const float *ln_g, *ln_b;
  bf16 *XA, *QA;            // operand tiles: x and LN(x)                 (bf16 flavour)
  float* Xf;                // x rows, fp32                                 (fp32 flavour)
  float* QN;                // LN(x) rows, fp32 (residual)
  int L, C;
};
// one group of D/8 lanes per row, 8 consecutive features per lane
template <int D, bool F32>
__global__ void __launch_bounds__(256) rows_embed_ln_kernel(const EmbedArgs a) {
  constexpr int G = D / 8, RPW = 32 / G;
  __shared__ float mct[8][D];
  pdl_trigger();
  for (int i = threadIdx.x; i < 8 * D; i += blockDim.x) mct[i % 8][i / 8] = a.Mc[i];
  __syncthreads();
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), n_warps = (int)(gridDim.x * blockDim.x) >> 5;
  pdl_wait();
  const int R = *a.n_rows;
  for (long long r0 = (long long)warp * RPW; r0 < R; r0 += (long long)n_warps * RPW) {
    const long long r = r0 + sub;
    const bool live = r < R;
    float x[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) x[e] = 0.f;
    if (live) {
      const int src = a.row_src[r];
      const int id = a.row_id[r];
      const int pos = src & 255, u = (src >> 8) & 0x7fffff;
      if (src >= 0) {
        ldg256(a.T + (long long)id * D + 8 * l, x);
        // (fixed trip count: all context loads go out together with the table row; a run-time `k < C` loop would
        // issue one, wait for it, issue the next)
        const float* c = a.p_c + ((long long)u * a.L + pos) * a.C;
        float cv[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) cv[k] = k < a.C ? __ldg(c + k) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
#pragma unroll
          for (int e = 0; e < 8; ++e) x[e] = fmaf(mct[k][8 * l + e], cv[k], x[e]);
        }
        if (a.pos) {
          const float4* pp = reinterpret_cast<const float4*>(a.pos + (long long)pos * D + 8 * l);
          const float4 p0 = __ldg(pp), p1 = __ldg(pp + 1);
          x[0] += p0.x; x[1] += p0.y; x[2] += p0.z; x[3] += p0.w;
          x[4] += p1.x; x[5] += p1.y; x[6] += p1.z; x[7] += p1.w;
        }
      }
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += x[e];
    const float mean = group_sum<G>(s) * (1.0f / D);
    float m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) m2 = fmaf(x[e] - mean, x[e] - mean, m2);
    const float rstd = rsqrtf(group_sum<G>(m2) * (1.0f / D) + kLnEps);
    if (live) {
      float q[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_g + 8 * l)), g1 = __ldg(reinterpret_cast<const float4*>(a.ln_g + 8 * l) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.ln_b + 8 * l)), b1 = __ldg(reinterpret_cast<const float4*>(a.ln_b + 8 * l) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int e = 0; e < 8; ++e) q[e] = (x[e] - mean) * rstd * gg[e] + bb[e];
      if (F32) {
        stg256(a.Xf + r * D + 8 * l, x);
      } else {
        *reinterpret_cast<uint4*>(a.XA + tile_off<D>(r, l)) = pack8(x);
        *reinterpret_cast<uint4*>(a.QA + tile_off<D>(r, l)) = pack8(q);
      }
      stg256(a.QN + r * D + 8 * l, q);
    }
  }
}

// ---------------------------------------------------------------------------------------------- attention + LN2
struct AttnRowsArgs {
  const void *Q, *K, *V;    // rows [R, D]: bf16 (bf16 flavour) or fp32
  int ldkv;                 // row stride of K and V in elements (fp32 flavour: K | V come from one fused GEMM, 2 D)
  const float* QN;          // LN1(x) rows (residual, src/carca.py:302)
  const int *row_src, *row_seg, *n_rows;
  const float *ln_g, *ln_b;
  float* S2;                // LN2 rows, fp32 (residual of the FFN, :316)
  bf16* S2A;                // LN2 operand tiles
  int residual;
};
// causal self-attention of one packed row over the rows of its own segment (src/carca.py:299 with :246-256: keys of
// the same user at positions <= the query's; a padding query row gives exactly 0), + LN1 residual, LayerNorm 2.
// One group of D/8 lanes per row; a lane owns 8 consecutive features, i.e. a slice of one head.
template <bool F32>
__device__ __forceinline__ void load_row8(const void* base, long long row, int ld, int l, bool on, float (&v)[8]) {
  if (!on) {
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] = 0.f;
    return;
  }
  if (F32) {
    ldg256(reinterpret_cast<const float*>(base) + row * ld + 8 * l, v);
  } else {
    unpack8(__ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const bf16*>(base) + row * ld) + l), v);
  }
}
template <int D, int H, bool F32>
__global__ void __launch_bounds__(256) rows_attn_ln_kernel(const AttnRowsArgs a) {
  constexpr int G = D / 8, RPW = 32 / G, DH = D / H, LPH = DH / 8;   // lanes per head
  static_assert(DH % 8 == 0 && LPH >= 1 && LPH <= 8, "head width");
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), n_warps = (int)(gridDim.x * blockDim.x) >> 5;
  const int R = *a.n_rows;
  const float sc = 1.4426950408889634f * rsqrtf((float)DH);
  for (long long r0 = (long long)warp * RPW; r0 < R; r0 += (long long)n_warps * RPW) {
    const long long r = r0 + sub;
    const bool live = r < R;
    const int src = live ? a.row_src[r] : (int)0x80000000u;
    const int s0 = live ? a.row_seg[r] : 0;
    // every lane of the warp runs the same number of iterations (shuffles inside): the longest of its rows
    int n_keys = (live && src >= 0) ? (int)(r - s0) + 1 : 0;
    int n_max = n_keys;
#pragma unroll
    for (int o = G; o < 32; o <<= 1) n_max = max(n_max, __shfl_xor_sync(kFull, n_max, o));
    float q[8], acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { q[e] = 0.f; acc[e] = 0.f; }
    if (n_keys > 0) {
      load_row8<F32>(a.Q, r, D, l, true, q);
#pragma unroll
      for (int e = 0; e < 8; ++e) q[e] *= sc;
    }
    // keys in chunks of KC: the chunk's K and V rows are loaded first (2 KC independent loads in flight per lane),
    // then the KC scores, one rescale of the running softmax per chunk
    constexpr int KC = 4;
    float m = -INFINITY, z = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < n_max; j0 += KC) {
      float kv[KC][8], vv[KC][8], sj[KC];
#pragma unroll
      for (int i = 0; i < KC; ++i) load_row8<F32>(a.K, (long long)s0 + j0 + i, a.ldkv, l, j0 + i < n_keys, kv[i]);
#pragma unroll
      for (int i = 0; i < KC; ++i) load_row8<F32>(a.V, (long long)s0 + j0 + i, a.ldkv, l, j0 + i < n_keys, vv[i]);
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        float t = 0.f;
#pragma unroll
        for (int e = 0; e < 8; ++e) t = fmaf(q[e], kv[i][e], t);
        sj[i] = t;
      }
#pragma unroll
      for (int o = LPH / 2; o > 0; o >>= 1)
#pragma unroll
        for (int i = 0; i < KC; ++i) sj[i] += __shfl_xor_sync(kFull, sj[i], o);
      float mn = m;
#pragma unroll
      for (int i = 0; i < KC; ++i) {
        sj[i] = j0 + i < n_keys ? sj[i] : -INFINITY;
        mn = fmaxf(mn, sj[i]);
      }
      if (mn > -INFINITY) {
        const float corr = ex2f(m - mn);
        float pj[KC], ps = 0.f;
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          pj[i] = ex2f(sj[i] - mn);
          ps += pj[i];
        }
        z = fmaf(z, corr, ps);
#pragma unroll
        for (int e = 0; e < 8; ++e) {
          float t = acc[e] * corr;
#pragma unroll
          for (int i = 0; i < KC; ++i) t = fmaf(pj[i], vv[i][e], t);
          acc[e] = t;
        }
        m = mn;
      }
    }
    float v[8];
    {
      const float inv = z > 0.f ? 1.0f / z : 0.f;
      float qn[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
      if (live && a.residual) ldg256(a.QN + r * D + 8 * l, qn);
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = fmaf(acc[e], inv, qn[e]);
    }
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
    const float mean = group_sum<G>(s) * (1.0f / D);
    float m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) m2 = fmaf(v[e] - mean, v[e] - mean, m2);
    const float rstd = rsqrtf(group_sum<G>(m2) * (1.0f / D) + kLnEps);
    if (live) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(a.ln_g + 8 * l)), g1 = __ldg(reinterpret_cast<const float4*>(a.ln_g + 8 * l) + 1);
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(a.ln_b + 8 * l)), b1 = __ldg(reinterpret_cast<const float4*>(a.ln_b + 8 * l) + 1);
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float o[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
      stg256(a.S2 + r * D + 8 * l, o);
      if (!F32) *reinterpret_cast<uint4*>(a.S2A + tile_off<D>(r, l)) = pack8(o);
    }
  }
}

// y[r] = LayerNorm(x[r]) over fp32 rows (fp32 flavour: LN1 of the next block / the final norm, src/carca.py:298,421)
template <int D>
__global__ void __launch_bounds__(256) rows_ln_kernel(float* __restrict__ Y, const float* __restrict__ X,
                                                      const float* __restrict__ g, const float* __restrict__ b,
                                                      const int* __restrict__ n_rows) {
  constexpr int G = D / 8, RPW = 32 / G;
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), n_warps = (int)(gridDim.x * blockDim.x) >> 5;
  const int R = *n_rows;
  for (long long r0 = (long long)warp * RPW; r0 < R; r0 += (long long)n_warps * RPW) {
    const long long r = r0 + sub;
    const bool live = r < R;
    float v[8];
    load_row8<true>(X, r, D, l, live, v);
    float s = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) s += v[e];
    const float mean = group_sum<G>(s) * (1.0f / D);
    float m2 = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) m2 = fmaf(v[e] - mean, v[e] - mean, m2);
    const float rstd = rsqrtf(group_sum<G>(m2) * (1.0f / D) + kLnEps);
    if (live) {
      float gg[8], bb[8], o[8];
      {
        const float4 g0 = __ldg(reinterpret_cast<const float4*>(g + 8 * l)), g1 = __ldg(reinterpret_cast<const float4*>(g + 8 * l) + 1);
        const float4 b0 = __ldg(reinterpret_cast<const float4*>(b + 8 * l)), b1 = __ldg(reinterpret_cast<const float4*>(b + 8 * l) + 1);
        gg[0] = g0.x; gg[1] = g0.y; gg[2] = g0.z; gg[3] = g0.w; gg[4] = g1.x; gg[5] = g1.y; gg[6] = g1.z; gg[7] = g1.w;
        bb[0] = b0.x; bb[1] = b0.y; bb[2] = b0.z; bb[3] = b0.w; bb[4] = b1.x; bb[5] = b1.y; bb[6] = b1.z; bb[7] = b1.w;
      }
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = (v[e] - mean) * rstd * gg[e] + bb[e];
      stg256(Y + r * D + 8 * l, o);
    }
  }
}

// fp32 flavour: per packed row the decoder's value fold u[r][h] = <V_h[r], wf_h> and the context terms of its key
// km[r][h][k] = <K_h[r], McQ_h[:, k]> (the bf16 flavour gets both from its GEMM epilogues).  One group of D/8 lanes per row.
template <int D, int H>
__global__ void __launch_bounds__(256) rows_fold_kv_kernel(float* __restrict__ U, float* __restrict__ KM,
                                                           const float* __restrict__ Kd, const float* __restrict__ Vd, int ld,
                                                           const float* __restrict__ wf, const float* __restrict__ McQ,
                                                           const int* __restrict__ n_rows) {
  constexpr int G = D / 8, RPW = 32 / G, DH = D / H, LPH = DH / 8;
  const int lane = threadIdx.x & 31, l = lane % G, sub = lane / G;
  const int warp = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5), n_warps = (int)(gridDim.x * blockDim.x) >> 5;
  const int R = *n_rows;
  float w8[8], mq[8][8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    w8[e] = __ldg(wf + 8 * l + e);
#pragma unroll
    for (int k = 0; k < 8; ++k) mq[e][k] = __ldg(McQ + (8 * l + e) * 8 + k);
  }
  for (long long r0 = (long long)warp * RPW; r0 < R; r0 += (long long)n_warps * RPW) {
    const long long r = r0 + sub;
    const bool live = r < R;
    float kv[8], vv[8];
    load_row8<true>(Kd, r, ld, l, live, kv);
    load_row8<true>(Vd, r, ld, l, live, vv);
    float u = 0.f, km[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      u = fmaf(vv[e], w8[e], u);
#pragma unroll
      for (int k = 0; k < 8; ++k) km[k] = fmaf(kv[e], mq[e][k], km[k]);
    }
#pragma unroll
    for (int o = LPH / 2; o > 0; o >>= 1) {
      u += __shfl_xor_sync(kFull, u, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) km[k] += __shfl_xor_sync(kFull, km[k], o);
    }
    if (live && l % LPH == 0) {
      const int h = l / LPH;
      U[r * H + h] = u;
      float4* o = reinterpret_cast<float4*>(KM + (r * H + h) * 8);
      o[0] = make_float4(km[0], km[1], km[2], km[3]);
      o[1] = make_float4(km[4], km[5], km[6], km[7]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- tcgen05 GEMM
enum { EPI_ROWS = 0, EPI_LRELU_TILE = 1, EPI_LN = 2, EPI_VDOT = 3, EPI_KDEC = 4, EPI_BIAS_TILE = 5, EPI_KMAJ = 6, EPI_VMN = 7 };
// EPI_BIAS_TILE / EPI_KMAJ / EPI_VMN feed the tensor-core attention (rows_attn_tc.cuh): Q as operand tiles, K as one
// K-major operand over all rows [D/8][ld_rows][8], V as MN-major 8 x 8 blocks per head [H][ld_rows/8][DH/8][8][8].
// They write EVERY row of a tile, rows past the batch as exact zeros (an MMA reads them; 0 * garbage must stay 0).

struct GemmJob {
  const bf16* A;            // operand tiles [n_tiles][D/8][128][8]
  const bf16* W;            // packed weight [D/8][D (out features)][8]
  const float* bias;        // [D]
  int epi;
  bf16* out_rows;           // EPI_ROWS: bf16 rows [R, D]
  bf16* out_tile;           // EPI_LRELU_TILE / EPI_BIAS_TILE: activation tiles; EPI_LN: tiles of the pre-norm value x'
                            //   (may be null); EPI_KMAJ / EPI_VMN: the attention operand
  long long ld_rows;        // EPI_KMAJ / EPI_VMN: rows of the operand (multiple of 128)
  const float* resid;       // EPI_LN: fp32 rows added before the norm (null: no residual)
  const float *ln_g, *ln_b; // EPI_LN
  float* out_f32;           // EPI_LN: LN rows fp32; EPI_KDEC: key rows fp32 [R, D]
  bf16* out_tile2;          // EPI_LN: LN tiles
  const float* wf;          // EPI_VDOT: scorer weight [D]
  float* U;                 // EPI_VDOT: u[r][h] = <V_h[r], wf_h>
  const float* McQ;         // EPI_KDEC: context map of the query side [D][8]
  float* KM;                // EPI_KDEC: km[r][h][k] = <K_h[r], McQ_h[:, k]>
};
struct GemmArgs {
  GemmJob job[3];
  int n_jobs;
  const int* n_rows;
  int H;
  int* status;
};

template <int D>
struct GemmCfg {
  static constexpr int KG = D / 8;
  static constexpr int CHUNK_KG = 8;                         // 64 k per A stage
  static constexpr int CHUNKS = KG / CHUNK_KG;
  static constexpr int STAGE_BYTES = CHUNK_KG * TILE * 16;   // 16 KB
  static constexpr int STAGES = D == 256 ? 5 : 4;
  static constexpr int W_BYTES = KG * D * 16;
  static constexpr int PARAM_FLOATS = 12 * D;                // bias | ln_g | ln_b | wf | McQ^T [8][D]
  static constexpr int TMEM_COLS = 2 * D;
  static constexpr size_t SMEM = (size_t)W_BYTES + (size_t)STAGES * STAGE_BYTES + PARAM_FLOATS * 4 + 256;
};
constexpr int GEMM_THREADS = 320;   // producer warp, MMA warp, 8 epilogue warps

__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(umma::smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(umma::smem_u32(bar)) : "memory");
}
// TMA engine, non-tensor form: one contiguous global -> shared copy that completes on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   umma::smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(umma::smem_u32(bar))
               : "memory");
}
__host__ __device__ constexpr uint32_t idesc_bf16(int n) {   // bf16 x bf16 -> f32, both K-major, M = 128
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((128u >> 4) << 24);
}
__device__ __forceinline__ void mma_bf16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, bool acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      :
      : "r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"((uint32_t)acc)
      : "memory");
}
// wait bounded in TIME (a failed mbarrier.try_wait may itself suspend for a while, so a spin COUNT bounds nothing):
// on a timeout (~0.25 s) the caller's bit is set in the status word and the role stops — never a hang
__device__ __forceinline__ bool wait_or_flag(uint64_t* bar, uint32_t parity, int* status, int bit = 2) {
  const long long t0 = clock64();
#pragma unroll 1
  for (;;) {
#pragma unroll 1
    for (int i = 0; i < 64; ++i)
      if (umma::mbar_try(bar, parity)) return true;
    if (clock64() - t0 > 500000000ll) break;
  }
  atomicOr(status, bit);
  return false;
}

template <int D>
// (d = 64: two CTAs per SM — 102 registers per thread; without the bound the kernel takes 160 and, at one CTA per SM,
// a grid sized for more runs in waves that each pay the prologue)
__global__ void __launch_bounds__(GEMM_THREADS, (D == 64 ? 2 : 1)) rows_gemm_kernel(const GemmArgs a) {
  using Cfg = GemmCfg<D>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  unsigned char* w_smem = smem_raw;
  unsigned char* a_smem = smem_raw + Cfg::W_BYTES;
  float* prm = reinterpret_cast<float*>(a_smem + Cfg::STAGES * Cfg::STAGE_BYTES);
  uint64_t* bars = reinterpret_cast<uint64_t*>(prm + Cfg::PARAM_FLOATS);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = bars + Cfg::STAGES;        // [STAGES]
  uint64_t* acc_full = empty + Cfg::STAGES;    // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* w_full = acc_empty + 2;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(w_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int jid = blockIdx.x % a.n_jobs, slot = blockIdx.x / a.n_jobs, n_slots = gridDim.x / a.n_jobs;
  const GemmJob& J = a.job[jid];
  if (slot >= n_slots) return;                 // (gridDim.x not a multiple of n_jobs)
  pdl_trigger();

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < Cfg::STAGES; ++s) { umma::mbar_init(&full[s], 1); umma::mbar_init(&empty[s], 1); }
    for (int s = 0; s < 2; ++s) { umma::mbar_init(&acc_full[s], 1); umma::mbar_init(&acc_empty[s], 4); }
    umma::mbar_init(w_full, 1);
  }
  if (warp == 1) umma::tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  // epilogue parameters -> shared memory
  for (int i = threadIdx.x; i < D; i += GEMM_THREADS) {
    prm[i] = J.bias ? J.bias[i] : 0.f;
    prm[D + i] = J.ln_g ? J.ln_g[i] : 1.f;
    prm[2 * D + i] = J.ln_b ? J.ln_b[i] : 0.f;
    prm[3 * D + i] = J.wf ? J.wf[i] : 0.f;
  }
  if (J.epi == EPI_KDEC)
    for (int i = threadIdx.x; i < 8 * D; i += GEMM_THREADS) prm[4 * D + (i % 8) * D + i / 8] = J.McQ[i];
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = *tmem_slot;
  pdl_wait();                                  // the predecessor's rows (and the row count) from here on
  const int R = *a.n_rows;
  const int n_tiles = (R + TILE - 1) / TILE;

  if (warp == 0) {
    // ===== producer: the weight matrix once, then the A chunks of this CTA's tiles through the stage ring
    if (lane == 0 && slot < n_tiles) {
      mbar_expect_tx(w_full, Cfg::W_BYTES);
      for (int off = 0; off < Cfg::W_BYTES; off += 16384)
        bulk_g2s(w_smem + off, reinterpret_cast<const unsigned char*>(J.W) + off, min(16384, Cfg::W_BYTES - off), w_full);
      uint32_t it = 0;
      for (int tile = slot; tile < n_tiles; tile += n_slots) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(J.A) + (size_t)tile * (TILE * D * 2);
        for (int c = 0; c < Cfg::CHUNKS; ++c, ++it) {
          const uint32_t s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1u;
          if (!wait_or_flag(&empty[s], ph ^ 1u, a.status)) break;
          mbar_expect_tx(&full[s], Cfg::STAGE_BYTES);
          bulk_g2s(a_smem + s * Cfg::STAGE_BYTES, src + (size_t)c * Cfg::STAGE_BYTES, Cfg::STAGE_BYTES, &full[s]);
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one elected lane)
    if (slot < n_tiles) {
      constexpr uint32_t idesc = idesc_bf16(D);
      constexpr uint32_t a_lbo = TILE * 16, b_lbo = D * 16;
      const uint32_t w_addr = umma::smem_u32(w_smem), a_addr0 = umma::smem_u32(a_smem);
      bool ok = wait_or_flag(w_full, 0, a.status);
      uint32_t it = 0, t = 0;
      for (int tile = slot; tile < n_tiles && ok; tile += n_slots, ++t) {
        const uint32_t as = t & 1u, aph = (t >> 1) & 1u;
        ok = wait_or_flag(&acc_empty[as], aph ^ 1u, a.status);       // epilogue drained this accumulator
        umma::fence_after_sync();
        const uint32_t d_tmem = tmem0 + as * D;
        for (int c = 0; c < Cfg::CHUNKS && ok; ++c, ++it) {
          const uint32_t s = it % Cfg::STAGES, ph = (it / Cfg::STAGES) & 1u;
          ok = wait_or_flag(&full[s], ph, a.status);
          umma::fence_after_sync();
          if (umma::elect_one()) {
            const uint32_t a_addr = a_addr0 + s * Cfg::STAGE_BYTES;
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              const uint64_t da = umma::smem_desc(a_addr + ks * 2 * a_lbo, a_lbo, 128);
              const uint64_t db = umma::smem_desc(w_addr + (c * 8 + ks * 2) * b_lbo, b_lbo, 128);
              mma_bf16_ss(d_tmem, da, db, idesc, !(c == 0 && ks == 0));
            }
            umma::commit(&empty[s]);                                 // stage free when these MMAs have read it
            if (c == Cfg::CHUNKS - 1) umma::commit(&acc_full[as]);   // accumulator complete
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===== epilogue: group g = (warp - 2) / 4 takes the tiles whose accumulator stage is g
    const int g = (warp - 2) >> 2, quarter = warp & 3;
    const int row_in_tile = 32 * quarter + lane;
    const uint32_t lane_base = tmem0 + ((uint32_t)(32 * quarter) << 16) + g * D;
    const int H = a.H, DH = D / H;
    uint32_t t = 0;
    bool ok = true;
    for (int tile = slot; tile < n_tiles && ok; tile += n_slots, ++t) {
      if ((int)(t & 1u) != g) continue;
      const uint32_t aph = (t >> 1) & 1u;
      const long long r = (long long)tile * TILE + row_in_tile;
      const bool live = r < R;
      if (J.epi == EPI_LN && J.resid != nullptr && live) {
        // the residual row is the epilogue's one dependent global read: sent to L2 while the tile's MMAs still run
        // (ncu: 30 % of this kernel's samples sat on its first use when only the next chunk was requested ahead)
#pragma unroll
        for (int c = 0; c < D / 32; ++c) asm volatile("prefetch.global.L2 [%0];" ::"l"(J.resid + r * D + 32 * c));
      }
      ok = wait_or_flag(&acc_full[g], aph, a.status);
      umma::fence_after_sync();
      if (J.epi == EPI_ROWS || J.epi == EPI_LRELU_TILE || J.epi >= EPI_BIAS_TILE) {
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float v[32];
          umma::tmem_ld_1x32(lane_base + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            v[e] += prm[32 * c + e];
            if (J.epi == EPI_LRELU_TILE) v[e] = v[e] > 0.f ? v[e] : kLeakySlope * v[e];
            if (J.epi >= EPI_KMAJ) v[e] = live ? v[e] : 0.f;
          }
          if (J.epi == EPI_ROWS) {
            if (live) {   // 64 contiguous bytes of this thread's row as two full-sector stores
              bf16* o = J.out_rows + r * D + 32 * c;
              stg256u(o, pack8(&v[0]), pack8(&v[8]));
              stg256u(o + 16, pack8(&v[16]), pack8(&v[24]));
            }
          } else if (J.epi == EPI_KMAJ) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
              *reinterpret_cast<uint4*>(J.out_tile + ((long long)(4 * c + q) * J.ld_rows + r) * 8) = pack8(&v[8 * q]);
          } else if (J.epi == EPI_VMN) {
            const int FG = DH / 8;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int g = 4 * c + q, h = g / FG, j = g % FG;
              *reinterpret_cast<uint4*>(J.out_tile + (((long long)h * (J.ld_rows >> 3) + (r >> 3)) * FG + j) * 64 + (r & 7) * 8) =
                  pack8(&v[8 * q]);
            }
          } else {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(J.out_tile + tile_off<D>(r, 4 * c + q)) = pack8(&v[8 * q]);
          }
        }
      } else if (J.epi == EPI_LN) {
        // pass 1: x' = acc + bias (+ residual) back into TMEM with the row's sum and sum of squares (fp32; LayerNorm
        // inputs are O(1) with |mean| << std, so E[x^2] - mean^2 loses nothing at this precision).  The residual of
        // chunk c+1 is requested before chunk c is processed: its latency hides behind the TMEM round trip.
        float sum = 0.f, sq = 0.f;
        float rs[32];
        const bool use_res = J.resid != nullptr && live;
        if (use_res) {
#pragma unroll
          for (int q = 0; q < 4; ++q) ldg256(J.resid + r * D + 8 * q, *reinterpret_cast<float(*)[8]>(&rs[8 * q]));
        }
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float v[32];
          umma::tmem_ld_1x32(lane_base + 32 * c, v);
          if (use_res) {
#pragma unroll
            for (int e = 0; e < 32; ++e) v[e] += rs[e];
            if (c + 1 < D / 32) {
#pragma unroll
              for (int q = 0; q < 4; ++q)
                ldg256(J.resid + r * D + 32 * (c + 1) + 8 * q, *reinterpret_cast<float(*)[8]>(&rs[8 * q]));
            }
          }
#pragma unroll
          for (int e = 0; e < 32; ++e) {
            v[e] += prm[32 * c + e];
            sum += v[e];
            sq = fmaf(v[e], v[e], sq);
          }
          umma::tmem_st_x32(lane_base + 32 * c, v, 0);
          if (J.out_tile) {
#pragma unroll
            for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(J.out_tile + tile_off<D>(r, 4 * c + q)) = pack8(&v[8 * q]);
          }
        }
        umma::tmem_st_wait();
        const float mean = sum * (1.0f / D);
        const float rstd = rsqrtf(fmaxf(sq * (1.0f / D) - mean * mean, 0.f) + kLnEps);
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float v[32];
          umma::tmem_ld_1x32(lane_base + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] = (v[e] - mean) * rstd * prm[D + 32 * c + e] + prm[2 * D + 32 * c + e];
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) stg256(J.out_f32 + r * D + 32 * c + 8 * q, &v[8 * q]);
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) *reinterpret_cast<uint4*>(J.out_tile2 + tile_off<D>(r, 4 * c + q)) = pack8(&v[8 * q]);
        }
      } else if (J.epi == EPI_VDOT) {
        float u = 0.f;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float v[32];
          umma::tmem_ld_1x32(lane_base + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) u = fmaf(v[e] + prm[32 * c + e], prm[3 * D + 32 * c + e], u);
          if ((32 * (c + 1)) % DH == 0) {
            if (live) J.U[r * H + (32 * c) / DH] = u;
            u = 0.f;
          }
        }
      } else {   // EPI_KDEC
        float km[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) km[k] = 0.f;
#pragma unroll 1
        for (int c = 0; c < D / 32; ++c) {
          float v[32];
          umma::tmem_ld_1x32(lane_base + 32 * c, v);
#pragma unroll
          for (int e = 0; e < 32; ++e) v[e] += prm[32 * c + e];
          if (live) {
#pragma unroll
            for (int q = 0; q < 4; ++q) stg256(J.out_f32 + r * D + 32 * c + 8 * q, &v[8 * q]);
          }
#pragma unroll
          for (int k = 0; k < 8; ++k) {
            const float* mk = prm + 4 * D + k * D + 32 * c;
#pragma unroll
            for (int e = 0; e < 32; ++e) km[k] = fmaf(v[e], mk[e], km[k]);
          }
          if ((32 * (c + 1)) % DH == 0) {
            if (live) {
              float4* o = reinterpret_cast<float4*>(J.KM + (r * H + (32 * c) / DH) * 8);
              o[0] = make_float4(km[0], km[1], km[2], km[3]);
              o[1] = make_float4(km[4], km[5], km[6], km[7]);
            }
#pragma unroll
            for (int k = 0; k < 8; ++k) km[k] = 0.f;
          }
        }
      }
      // accumulator stage g is free again
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[g]);
    }
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 1) umma::tmem_free(tmem0, Cfg::TMEM_COLS);
}

// ---------------------------------------------------------------------------------------------- decoders
struct DecodeArgs {
  // cross-attention (src/carca.py:338-347): keys of the encoded profile and the folded candidate tables
  const float* Kd;          // decoder keys, fp32 rows (row stride ldk)
  int ldk;
  const float* U;           // u[r][h] = <V_h[r], wf_h>            (bf16 flavour: from the GEMM epilogue)
  const float* KM;          // km[r][h][k]
  const float* TQ;          // WQ T[i] + bq  [n_items, D] fp32
  const float* tw;          // <T[i], wf>    [n_items]
  const float* mcw;         // wf Mc         [8]
  const float* dbf;         // scorer bias   [1]
  // dot product (:358-365)
  const float* PE;          // encoded profile rows fp32 [R, D]
  const float* Tf;          // folded item table fp32
  const float* Mc;          // [D][8]
  const int2* useg;
  const int* row_src;
  const int* o_x;
  const float* o_c;
  long long oc_user, oc_tgt;
  float* y;
  long long ldy;
  int col0, B, T, C, L, cat_lo, residual_ca;
};
// keys staged per pass: the decoder (the stage whose rounding lands directly on the logit) runs in fp32 on fp32 tables
template <int D> struct DecCfg {
  static constexpr int KEYS = D >= 128 ? 32 : 64;
  static constexpr int threads(int H) { return H > 4 ? 32 * H : 128; }
};

// One CTA per (user, slice of 128 / H candidates); thread = (candidate, head) with the head constant per WARP
// (warp w serves head w % H), so the key reads from shared memory are warp-wide broadcasts (one address per
// instruction, no bank conflicts) and every thread carries only its head's query slice (DH floats).  A thread first
// issues the gather of that slice (TQ[id][h*DH ..] as 256-bit loads) and only then helps staging the user's keys: the
// gather latency hides behind the staging and its barrier.  Softmax runs online; the attention output is folded into
// u_h[j] = <V_h[j], wf_h> (src/carca.py:343-345: the scorer is Linear(d, 1) on attn + o); the H partial results of a
// candidate meet in shared memory.
template <int D, int H, bool F32>
__global__ void __launch_bounds__(DecCfg<D>::threads(H)) rows_decode_ca_kernel(const DecodeArgs a) {
  constexpr int NT = DecCfg<D>::threads(H);      // one warp per head at least: 8 heads -> 256 threads
  constexpr int DH = D / H, DEC_KEYS = DecCfg<D>::KEYS, CPB = NT / H;
  __shared__ __align__(16) float ks[DEC_KEYS][D];
  __shared__ float us[DEC_KEYS][H];
  __shared__ __align__(16) float kms[DEC_KEYS][H][8];
  __shared__ float kvalid[DEC_KEYS];
  __shared__ float part[H][CPB];
  const float sc = 1.4426950408889634f * rsqrtf((float)DH);
  const int n_slices = (a.T + CPB - 1) / CPB;
  const float bfv = __ldg(a.dbf);
  const int wid = threadIdx.x >> 5;
  const int h = wid % H, cl = (wid / H) * 32 + (threadIdx.x & 31);
  for (long long item = blockIdx.x; item < (long long)a.B * n_slices; item += gridDim.x) {
    const int u = (int)(item / n_slices), t = (int)(item % n_slices) * CPB + cl;
    const int2 sg = a.useg[u];
    const bool has = t < a.T;
    const int id = !has ? 0 : (a.cat_lo > 0 ? a.cat_lo + t : __ldg(a.o_x + (long long)u * a.T + t));
    float q[DH];
    if (id != 0) {
      const float* qp = a.TQ + (long long)id * D + h * DH;
#pragma unroll
      for (int i = 0; i < DH / 8; ++i) ldg256(qp + 8 * i, *reinterpret_cast<float(*)[8]>(&q[8 * i]));
    }
    float cv[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) cv[k] = 0.f;
    if (id != 0) {
      const float* c = a.o_c + (long long)u * a.oc_user + (long long)t * a.oc_tgt;
      for (int k = 0; k < a.C; ++k) cv[k] = __ldg(c + k);
    }
    float m = -INFINITY, z = 0.f, d = 0.f;
    for (int k0 = 0; k0 < sg.y; k0 += DEC_KEYS) {
      const int nk = min(DEC_KEYS, sg.y - k0);
      __syncthreads();
      for (int i = threadIdx.x; i < nk * (D / 4); i += NT)
        reinterpret_cast<float4*>(&ks[0][0])[i] =
            __ldg(reinterpret_cast<const float4*>(a.Kd + (long long)(sg.x + k0 + i / (D / 4)) * a.ldk) + i % (D / 4));
      for (int i = threadIdx.x; i < nk * H; i += NT) us[i / H][i % H] = a.U[(long long)(sg.x + k0) * H + i];
      for (int i = threadIdx.x; i < nk * H * 8; i += NT) (&kms[0][0][0])[i] = a.KM[(long long)(sg.x + k0) * H * 8 + i];
      for (int i = threadIdx.x; i < nk; i += NT) kvalid[i] = a.row_src[sg.x + k0 + i] >= 0 ? 1.f : 0.f;
      __syncthreads();
      if (id != 0) {
        for (int j = 0; j < nk; ++j) {
          if (kvalid[j] == 0.f) continue;                   // padding key (position L-1 of a short window)
          const float4* kp = reinterpret_cast<const float4*>(&ks[j][h * DH]);
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int i = 0; i < DH / 4; ++i) {
            const float4 k4 = kp[i];
            s0 = fmaf(q[4 * i], k4.x, s0); s1 = fmaf(q[4 * i + 1], k4.y, s1);
            s0 = fmaf(q[4 * i + 2], k4.z, s0); s1 = fmaf(q[4 * i + 3], k4.w, s1);
          }
          float s = s0 + s1;
#pragma unroll
          for (int k = 0; k < 8; ++k) s = fmaf(cv[k], kms[j][h][k], s);
          s *= sc;
          const float mn = fmaxf(m, s);
          const float corr = ex2f(m - mn), p = ex2f(s - mn);
          z = fmaf(z, corr, p);
          d = fmaf(d, corr, p * us[j][h]);
          m = mn;
        }
      }
    }
    part[h][cl] = (id != 0 && z > 0.f) ? d / z : 0.f;
    __syncthreads();
    if (has && h == 0) {
      float acc = bfv;
      if (id != 0) {
#pragma unroll
        for (int hh = 0; hh < H; ++hh) acc += part[hh][cl];
        if (a.residual_ca) {
          acc += __ldg(a.tw + id);
          for (int k = 0; k < a.C; ++k) acc = fmaf(__ldg(a.mcw + k), cv[k], acc);
        }
      }
      a.y[(long long)u * a.ldy + a.col0 + t] = 1.0f / (1.0f + expf(-acc));
    }
  }
}

// dot decoder (src/carca.py:362): y = sigmoid(<p[L-1], e_t>), e_t = T[id] + Mc c_t (0 for a padding candidate)
template <int D>
__global__ void __launch_bounds__(128) rows_decode_dot_kernel(const DecodeArgs a) {
  __shared__ __align__(16) float pl[D];
  __shared__ float pmc[8];
  const int n_slices = (a.T + 127) / 128;
  for (long long item = blockIdx.x; item < (long long)a.B * n_slices; item += gridDim.x) {
    const int u = (int)(item / n_slices), t = (int)(item % n_slices) * 128 + threadIdx.x;
    const int2 sg = a.useg[u];
    const long long last = (long long)sg.x + sg.y - 1;     // position L-1 is always the segment's last row
    __syncthreads();
    for (int i = threadIdx.x; i < D; i += 128) pl[i] = a.PE[last * D + i];
    __syncthreads();
    if (threadIdx.x < 8) {
      float s = 0.f;
      if (threadIdx.x < a.C)
        for (int i = 0; i < D; ++i) s = fmaf(pl[i], a.Mc[i * 8 + threadIdx.x], s);
      pmc[threadIdx.x] = s;
    }
    __syncthreads();
    if (t < a.T) {
      const int id = a.cat_lo > 0 ? a.cat_lo + t : __ldg(a.o_x + (long long)u * a.T + t);
      float acc = 0.f;
      if (id != 0) {
        const float4* tp = reinterpret_cast<const float4*>(a.Tf + (long long)id * D);
        float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
        for (int i = 0; i < D / 4; ++i) {
          const float4 e4 = __ldg(tp + i);
          s0 = fmaf(pl[4 * i], e4.x, s0); s1 = fmaf(pl[4 * i + 1], e4.y, s1);
          s0 = fmaf(pl[4 * i + 2], e4.z, s0); s1 = fmaf(pl[4 * i + 3], e4.w, s1);
        }
        acc = s0 + s1;
        const float* c = a.o_c + (long long)u * a.oc_user + (long long)t * a.oc_tgt;
        float c8[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) c8[k] = k < a.C ? __ldg(c + k) : 0.f;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc = fmaf(k < a.C ? pmc[k] : 0.f, c8[k], acc);
      }
      a.y[(long long)u * a.ldy + a.col0 + t] = 1.0f / (1.0f + expf(-acc));
    }
  }
}

// ---------------------------------------------------------------------------------------------- plan preparation
// fp32 [n, D] -> bf16
__global__ void __launch_bounds__(256) to_bf16_kernel(bf16* __restrict__ dst, const float* __restrict__ src, long long n) {
  const long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (i + 3 < n) {
    const float4 v = *reinterpret_cast<const float4*>(src + i);
    *reinterpret_cast<uint2*>(dst + i) = make_uint2(pack2(v.x, v.y), pack2(v.z, v.w));
  } else {
    for (long long j = i; j < n; ++j) dst[j] = __float2bfloat16(src[j]);
  }
}
// W [N = D out][K = D in] fp32 (nn.Linear / Conv1d k=1 layout) -> B operand [K/8][N][8] bf16
__global__ void __launch_bounds__(256) pack_weight_bf16_kernel(bf16* __restrict__ dst, const float* __restrict__ W, int D) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= D * D) return;
  const int n = i / D, k = i % D;
  dst[((long long)(k / 8) * D + n) * 8 + (k % 8)] = __float2bfloat16(W[i]);
}

}  // namespace rows
}  // namespace carca
#endif  // CARCA_EMU

// ---------------------------------------------------------------------------------------------- tcgen05 decoder
#ifndef CARCA_EMU
namespace carca {
namespace rows {

// Cross-attention decoder of the bf16 flavour on the tensor cores (src/carca.py:338-347).  Work item = (user, tile of
// 128 candidates).  Four producer warps (thread = candidate row) GATHER the candidates' folded query rows TQ[id] (bf16
// table) straight into the K-major operand layout with 16-byte cp.async copies and stage the user's keys (fp32 rows ->
// bf16 operand) and their per-key terms; one elected lane issues per head h  S_h[128 x 64] = TQ_h K_h^T
// (tcgen05.mma kind::f16, M = 128, N = 64, K = d / H) into one of two TMEM stages; four epilogue warps (thread =
// candidate) run the softmax over the user's keys, fold the attention output through u_h[j] = <V_h[j], wf_h> and write
// sigmoid(att + <o, wf> + bf).  Users with more than 64 keys take several key chunks (online softmax across chunks).
// Two stages: the gather of item i+1 runs under the MMAs / epilogue of item i.
constexpr int DT_KEYS = 64;
constexpr int DT_THREADS = 416;   // warps 0..3 epilogue, 4 MMA, 5..8 producers of stage 0, 9..12 producers of stage 1

template <int D, int H>
struct DecTcSmem {
  static constexpr int NST = 2;                    // stages (96 KB each at d = 256; two CTAs per SM at d = 64)
  static constexpr int A_BYTES = (D / 8) * 128 * 16;
  static constexpr int K_BYTES = (D / 8) * DT_KEYS * 16;
  unsigned char a[NST][A_BYTES];
  unsigned char k[NST][K_BYTES];
  float kc[NST][H][DT_KEYS];        // context term of the key per head (one context row per user), -1e30: masked
  float uu[NST][H][DT_KEYS];        // u_h[j] = <V_h[j], wf_h>
  alignas(16) float km[NST][DT_KEYS][H][8];   // per-candidate context: km[j][h][k] = <K_h[j], McQ_h[:, k]>
  // per candidate row, written by its producer thread: item id, residual term <o, wf> = tw[id] + <mcw, c>, context
  int rid[NST][128];
  float rres[NST][128];
  alignas(16) float rcv[NST][128][8];
  uint64_t full[NST], empty[NST], acc_full[NST], acc_empty[NST];
  uint32_t tmem_slot;
  int2 useg_c[64];                  // (first row, rows) of the users of this CTA's first 64 work items
};

struct DecTcArgs {
  const bf16* TQb;          // WQ T[i] + bq, bf16 [n_items, D]
  const float* Kd;          // decoder keys, fp32 rows [R, D]
  const float* U;           // [R, H]
  const float* KM;          // [R, H, 8]
  const float* tw;
  const float* mcw;
  const float* dbf;
  const int2* useg;
  const int* row_src;
  const int* o_x;
  const float* o_c;
  long long oc_user, oc_tgt;
  float* y;
  long long ldy;
  int col0, B, T, C, cat_lo, residual_ca;
  int* status;
};

template <int D, int H>
// (d = 64: two CTAs per SM)
__global__ void __launch_bounds__(DT_THREADS, (D >= 256 ? 1 : 2)) rows_decode_tc_kernel(const DecTcArgs a) {
  using SM = DecTcSmem<D, H>;
  constexpr int DH = D / H, KG = D / 8, NST = SM::NST;
  constexpr float kMask = -1.0e30f;
  extern __shared__ __align__(128) unsigned char dec_raw[];
  SM& s = *reinterpret_cast<SM*>(dec_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tiles = (a.T + 127) / 128;
  const bool uctx = a.oc_tgt == 0;

  pdl_trigger();
  if (threadIdx.x == 0) {
    for (int i = 0; i < NST; ++i) {
      umma::mbar_init(&s.full[i], 128);       // every producer thread arrives when its copies have landed
      umma::mbar_init(&s.empty[i], 1);        // tcgen05.commit: the MMAs have read the stage
      umma::mbar_init(&s.acc_full[i], 1);
      umma::mbar_init(&s.acc_empty[i], 4);    // one arrival per epilogue warp
    }
  }
  if (warp == 4) umma::tmem_alloc(&s.tmem_slot, NST * H * DT_KEYS);
  // key columns past a user's last key are masked by their -1e30 term, but what the MMA reads there must be finite
  for (int i = threadIdx.x; i < NST * SM::K_BYTES / 16; i += DT_THREADS) reinterpret_cast<uint4*>(s.k[0])[i] = make_uint4(0, 0, 0, 0);
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = s.tmem_slot;
  const float sc = 1.4426950408889634f * rsqrtf((float)DH);
  pdl_wait();
  // every role walks the same list of work items and needs each user's segment first: fetched once per CTA, up front
  // (in the step loop that load headed the producers' chain of dependent loads)
  for (int i = threadIdx.x; i < 64; i += DT_THREADS) {
    const unsigned item = blockIdx.x + (unsigned)i * gridDim.x;
    if (item < (unsigned)a.B * (unsigned)n_tiles) s.useg_c[i] = a.useg[n_tiles == 1 ? item : item / (unsigned)n_tiles];
  }
  __syncthreads();

  // the sequence of (user, candidate tile, key chunk) steps of this CTA: identical in every role
  // step `it`: user u, tile, key chunk k0 .. k0 + nk
  auto for_each_step = [&](auto&& body) {
    uint32_t it = 0;
    const unsigned n_items = (unsigned)a.B * (unsigned)n_tiles;       // (B < 2^23, a few tiles per user)
    int ic = 0;
    for (unsigned item = blockIdx.x; item < n_items; item += gridDim.x, ++ic) {
      const int u = n_tiles == 1 ? (int)item : (int)(item / (unsigned)n_tiles);
      const int t0 = n_tiles == 1 ? 0 : (int)(item % (unsigned)n_tiles) * 128;
      const int2 sg = ic < 64 ? s.useg_c[ic] : a.useg[u];
      for (int k0 = 0; k0 < max(sg.y, 1); k0 += DT_KEYS, ++it)
        if (!body(it, u, t0, sg, k0, min(DT_KEYS, sg.y - k0), k0 + DT_KEYS >= sg.y)) return;
    }
  };

  if (warp >= 5) {
    // ===== producers: thread r gathers candidate row r of the tile and helps staging the keys.  A step is a chain of
    // dependent global loads (candidate id -> table row, user segment -> key rows -> their terms): the two producer
    // groups each own one stage and take every other step, so two such chains are always in flight per CTA
    static_assert(NST == 2, "one producer group per stage");
    const int grp = (warp - 5) >> 2, r = threadIdx.x - 160 - 128 * grp;
    float mcw_r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) mcw_r[k] = k < a.C ? __ldg(a.mcw + k) : 0.f;
    for_each_step([&](uint32_t it, int u, int t0, int2 sg, int k0, int nk, bool last) {
      const int st = it % NST, ph = (it / NST) & 1;
      if (st != grp) return true;
      // Everything that comes from global memory into REGISTERS is requested before the stage is waited for: the id ->
      // table-row chain and the key rows then overlap the MMAs / epilogue still running on this stage.
      const int t = t0 + r;
      const int id = t < a.T ? (a.cat_lo > 0 ? a.cat_lo + t : __ldg(a.o_x + (long long)u * a.T + t)) : 0;
      constexpr int KPT = D == 64 ? KG / 2 : 0;      // key k-groups of this thread held in registers (d = 64 only)
      uint4 kreg[KPT > 0 ? KPT : 1];
      if (KPT > 0 && (r >> 1) < nk) {
        const float* kr = a.Kd + (long long)(sg.x + k0 + (r >> 1)) * D;
#pragma unroll
        for (int q = 0; q < KPT; ++q) {
          float v[8];
          ldg256(kr + 8 * ((r & 1) * KPT + q), v);
          kreg[q] = pack8(v);
        }
      }
      float cv[8], res = 0.f;
      {
#pragma unroll
        for (int k = 0; k < 8; ++k) cv[k] = 0.f;
        if (id != 0) {
          const float* c = a.o_c + (long long)u * a.oc_user + (long long)t * a.oc_tgt;
#pragma unroll
          for (int k = 0; k < 8; ++k) cv[k] = k < a.C ? __ldg(c + k) : 0.f;
          if (a.residual_ca) res = __ldg(a.tw + id);     // (second level of the id chain: first USED after the stage wait)
        }
      }
      // per-key terms of the user's keys: row flag -> value fold / context terms, another dependent chain
      constexpr int TERMS = (DT_KEYS * H + 127) / 128;
      float t_kc[TERMS], t_uu[TERMS];
      float4 t_m0[TERMS], t_m1[TERMS];
#pragma unroll
      for (int q = 0; q < TERMS; ++q) {
        const int i = r + 128 * q;
        const int j = i / H, h = i % H;
        t_kc[q] = kMask; t_uu[q] = 0.f;
        t_m0[q] = make_float4(0.f, 0.f, 0.f, 0.f); t_m1[q] = t_m0[q];
        if (i < DT_KEYS * H && j < nk) {
          // (the row's flag, value fold and context terms are requested together: one level of latency, not two; and
          // fixed trip counts — a run-time `k < C` loop issues one load, waits for it, issues the next: the per-role
          // clocks showed ~3,000 cycles per step in these few loops)
          const long long row = (long long)sg.x + k0 + j;
          const int flag = a.row_src[row];
          const float uval = a.U[row * H + h];
          const float4* kmp = reinterpret_cast<const float4*>(a.KM + (row * H + h) * 8);
          const float4 m0 = __ldg(kmp), m1 = __ldg(kmp + 1);
          if (flag < 0) continue;                    // padding key (position L-1 of a short window): stays masked
          t_uu[q] = uval;
          if (uctx) {
            const float* cu = a.o_c + (long long)u * a.oc_user;
            const float mm[8] = {m0.x, m0.y, m0.z, m0.w, m1.x, m1.y, m1.z, m1.w};
            float cu8[8];
#pragma unroll
            for (int k = 0; k < 8; ++k) cu8[k] = k < a.C ? __ldg(cu + k) : 0.f;
            float kcv = 0.f;
#pragma unroll
            for (int k = 0; k < 8; ++k) kcv = fmaf(mm[k], cu8[k], kcv);
            t_kc[q] = kcv;
          } else {
            t_kc[q] = 0.f;
            t_m0[q] = m0; t_m1[q] = m1;
          }
        }
      }
      if (!wait_or_flag(&s.empty[st], ph ^ 1, a.status, 4)) return false;
      if (!wait_or_flag(&s.acc_empty[st], ph ^ 1, a.status, 4)) return false;
      {   // what the epilogue needs of this row, so that it reads no global memory
        if (id != 0 && a.residual_ca) {
#pragma unroll
          for (int k = 0; k < 8; ++k) res = fmaf(mcw_r[k], cv[k], res);      // (cv[k] = 0 for k >= C)
        }
        s.rid[st][r] = t < a.T ? id : -1;
        s.rres[st][r] = res;
        if (!uctx) {
          reinterpret_cast<float4*>(s.rcv[st][r])[0] = make_float4(cv[0], cv[1], cv[2], cv[3]);
          reinterpret_cast<float4*>(s.rcv[st][r])[1] = make_float4(cv[4], cv[5], cv[6], cv[7]);
        }
      }
      if (id != 0) {
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.TQb + (long long)id * D);
        const uint32_t dst = umma::smem_u32(s.a[st]) + r * 16;
#pragma unroll 8
        for (int kg = 0; kg < KG; ++kg)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + kg * 2048), "l"(src + kg * 16) : "memory");
      }
      asm volatile("cp.async.commit_group;" ::: "memory");
      // keys of the chunk: fp32 rows -> bf16 operand [k/8][64 keys][8]; key j = r / 2, half of the k-groups each
      {
        const int j = r >> 1, hf = r & 1;
        uint4* kd = reinterpret_cast<uint4*>(s.k[st]);
        if (j < nk) {
          if (KPT > 0) {
#pragma unroll
            for (int q = 0; q < KPT; ++q) kd[(hf * KPT + q) * DT_KEYS + j] = kreg[q];
          } else {
            const float* kr = a.Kd + (long long)(sg.x + k0 + j) * D;
#pragma unroll 4
            for (int kg = hf * (KG / 2); kg < (hf + 1) * (KG / 2); ++kg) {
              float v[8];
              ldg256(kr + 8 * kg, v);
              kd[kg * DT_KEYS + j] = pack8(v);
            }
          }
        }
      }
      // per-key terms (loaded before the stage wait, see above)
#pragma unroll
      for (int q = 0; q < TERMS; ++q) {
        const int i = r + 128 * q;
        if (i < DT_KEYS * H) {
          const int j = i / H, h = i % H;
          if (!uctx) {   // masked keys: their context terms are still READ by the epilogue (whole chunks of 8 columns)
            reinterpret_cast<float4*>(&s.km[st][j][h][0])[0] = t_m0[q];
            reinterpret_cast<float4*>(&s.km[st][j][h][0])[1] = t_m1[q];
          }
          s.kc[st][h][j] = t_kc[q];
          s.uu[st][h][j] = t_uu[q];
        }
      }
      // this thread's stores are visible to the tensor core's (async-proxy) reads; its ARRIVAL is deferred to the
      // completion of its gather copies (cp.async.mbarrier.arrive.noinc: one of the barrier's 128 expected arrivals,
      // delivered by the copy unit), so the thread moves on to the next stage's gather without waiting for this one
      umma::fence_smem_to_async();
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(umma::smem_u32(&s.full[st])) : "memory");
      return true;
    });
  } else if (warp == 4) {
    // ===== MMA issuer
    constexpr uint32_t idesc = idesc_bf16(DT_KEYS);
    for_each_step([&](uint32_t it, int, int, int2, int, int, bool) {
      const int st = it % NST, ph = (it / NST) & 1;
      if (!wait_or_flag(&s.full[st], ph, a.status, 16)) return false;
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t ab = umma::smem_u32(s.a[st]), kb = umma::smem_u32(s.k[st]);
#pragma unroll
        for (int h = 0; h < H; ++h) {
          const uint32_t d = tmem0 + st * (H * DT_KEYS) + h * DT_KEYS;
#pragma unroll
          for (int ks = 0; ks < DH / 16; ++ks) {
            const int kg = h * (DH / 8) + 2 * ks;
            mma_bf16_ss(d, umma::smem_desc(ab + kg * 2048, 2048, 128), umma::smem_desc(kb + kg * (DT_KEYS * 16), DT_KEYS * 16, 128),
                        idesc, ks != 0);
          }
        }
        umma::commit(&s.empty[st]);
        umma::commit(&s.acc_full[st]);
      }
      __syncwarp();
      return true;
    });
  } else {
    // ===== epilogue: thread = candidate row
    const int r = 32 * warp + lane;
    const float bfv = __ldg(a.dbf);
    float m[H], z[H], dd[H];
#pragma unroll
    for (int h = 0; h < H; ++h) { m[h] = kMask; z[h] = 0.f; dd[h] = 0.f; }
    for_each_step([&](uint32_t it, int u, int t0, int2, int, int nk, bool last) {
      const int st = it % NST, ph = (it / NST) & 1;
      if (!wait_or_flag(&s.acc_full[st], ph, a.status, 32)) return false;
      umma::fence_after_sync();
      const int t = t0 + r;
      const int rid = s.rid[st][r];
      const bool has = rid >= 0;
      const int id = has ? rid : 0;
      float cv[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) cv[k] = 0.f;
      if (!uctx) {
        const float4 c0 = reinterpret_cast<const float4*>(s.rcv[st][r])[0], c1 = reinterpret_cast<const float4*>(s.rcv[st][r])[1];
        cv[0] = c0.x; cv[1] = c0.y; cv[2] = c0.z; cv[3] = c0.w; cv[4] = c1.x; cv[5] = c1.y; cv[6] = c1.z; cv[7] = c1.w;
      }
      const float res = s.rres[st][r];
      const uint32_t tb = tmem0 + ((uint32_t)(32 * warp) << 16) + st * (H * DT_KEYS);
      // Online softmax per CHUNK of 8 key columns, not per column: the 8 scores, their maximum (a tree), the 8
      // exponentials and the two sums are independent instructions; only one rescale per chunk sits on the running
      // (m, z, dd) chain (a per-column update is a serial max -> ex2 -> fma chain per key: per-role clocks showed the
      // epilogue warps busy 70 % of the kernel with it, 48 % after the change).
#pragma unroll 1
      for (int c0 = 0; c0 < nk; c0 += 8) {
#pragma unroll
        for (int h = 0; h < H; ++h) {
          float v[8];
          umma::tmem_ld_1x8(tb + h * DT_KEYS + c0, v);
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            float sv = v[e] + s.kc[st][h][c0 + e];          // masked / missing keys carry -1e30
            if (!uctx) {
#pragma unroll
              for (int k = 0; k < 8; ++k) sv = fmaf(cv[k], s.km[st][c0 + e][h][k], sv);
            }
            v[e] = sv * sc;
          }
          const float cm = fmaxf(fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3])), fmaxf(fmaxf(v[4], v[5]), fmaxf(v[6], v[7])));
          const float mn = fmaxf(m[h], cm);
          const float corr = ex2f(m[h] - mn);
          float ps0 = 0.f, ps1 = 0.f, pd0 = 0.f, pd1 = 0.f;
#pragma unroll
          for (int e = 0; e < 8; e += 2) {
            const float p0 = ex2f(v[e] - mn), p1 = ex2f(v[e + 1] - mn);
            ps0 += p0; ps1 += p1;
            pd0 = fmaf(p0, s.uu[st][h][c0 + e], pd0);
            pd1 = fmaf(p1, s.uu[st][h][c0 + e + 1], pd1);
          }
          z[h] = fmaf(z[h], corr, ps0 + ps1);
          dd[h] = fmaf(dd[h], corr, pd0 + pd1);
          m[h] = mn;
        }
      }
      if (last) {
        if (has) {
          float acc = bfv;
          if (id != 0) {
#pragma unroll
            for (int h = 0; h < H; ++h) acc += m[h] > -1.0e29f * 1.4426950408889634f ? dd[h] / z[h] : 0.f;
            acc += res;
          }
          a.y[(long long)u * a.ldy + a.col0 + t] = 1.0f / (1.0f + expf(-acc));
        }
#pragma unroll
        for (int h = 0; h < H; ++h) { m[h] = kMask; z[h] = 0.f; dd[h] = 0.f; }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s.acc_empty[st]);
      return true;
    });
  }
  umma::fence_before_sync();
  __syncthreads();
  if (warp == 4) umma::tmem_free(tmem0, NST * H * DT_KEYS);
}

}  // namespace rows
}  // namespace carca
#endif
