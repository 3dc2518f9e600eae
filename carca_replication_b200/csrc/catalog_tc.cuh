// Full-catalog scoring on the tensor cores (BASELINE configs[3]; d = 64): for an item shard [item_lo, item_lo + n) and
// ALL users of the call, the number of items a stable descending sort places before each user's positive
// (src/train.py:15-32 applied to eval-mode CARCA.forward over candidate chunks, src/carca.py:424-431) — without ever
// materialising the [users, items] score matrix.
//
// Scores are a GEMM between the (folded) item table and per-user column vectors:
//   dot decoder  (src/carca.py:362):   logit[i, u]   = <T[i], p_u[L-1]> + <Mc c_u, p_u[L-1]>
//   cross-attention (:338-347):        s_h[i, (u,j)] = <TQ_h[i], K_h[u, j]> + <McQ_h c_u, K_h[u, j]>   (per head h, key j)
//                                      logit[i, u]   = sum_h softmax_j(s_h / sqrt(dh)) . u_h[u, j] + tw[i] + <mcw, c_u> + bf
// Columns (users, or the keys of users) are packed into groups of 128 that never split a user; their operands are
// written once per call in the tensor core's K-major layout (3xTF32: hi | lo planes, scale folded in), so a group is
// one 64 KB bulk copy.  One CTA = (128-item tile, a contiguous range of column groups): the item tile is staged once
// (fp32 -> hi / lo), then per group 24 tcgen05.mma (kind::tf32, M = N = 128, hi*hi + lo*hi + hi*lo) fill 256 TMEM
// columns while the previous group's columns are reduced by the other epilogue warpgroup: thread = item row, online
// softmax along each user's columns, sigmoid, comparison with the positive's probability (ties by item id, as the
// stable sort), warp ballot, one atomicAdd per (warp, user).  Roles: warp 0 bulk-copy producer, warp 1 MMA issuer,
// warps 2..9 two epilogue warpgroups.
#pragma once
#include "common.cuh"
#ifndef CARCA_EMU
#include "rows_bf16.cuh"
#include "tmem_io.cuh"
#include "umma.cuh"

namespace carca {
namespace cat {

constexpr int CT = 128;                       // items per tile = columns per group
constexpr int GROUP_FLOATS = 2 * 16 * CT * 4; // [hi | lo][16 k-chunks (2 heads x 8, or 16 of one vector)][128][4]
constexpr int A_LBO = CT * 16 + 16;           // padded chunk stride of the thread-staged A operand (bytes)
constexpr int CAT_THREADS = 320;

// Column metadata of a group: 8 planes of 128 words (one bulk copy, read as float4 = 4 columns per instruction)
//   0 user | last << 30 (int bits; -1: unused column)     1, 2  kc_h: context term of the key per head, scaled; -inf: a
//   padding key (ca) / kc0 = <Mc c_u, p_last> (dot)        3, 4  u_h = <V_h[key], wf_h>
//   on the user's LAST column also: 5 probability of the positive, 6 the positive's item id (int bits), 7 <mcw, c_u>
constexpr int META_PLANES = 8;
constexpr int META_WORDS = META_PLANES * CT;   // per group
enum { MP_USER = 0, MP_KC0, MP_KC1, MP_U0, MP_U1, MP_YPOS, MP_PITEM, MP_CW };

struct CatArgs {
  const float* TA;          // item-side table: TQ (ca) or T (dot), [n_items, 64]
  const float* tw;          // ca: <T[i], wf>
  const float* Bt;          // column operand groups
  const float* meta;        // [n_groups][8][128]
  const int* n_groups;
  int* counts;
  const float* dbf;         // ca: scorer bias [1]
  int item_lo, n_items, residual_ca, parts, decoder;
  int* status;
  volatile int* dbg;        // optional progress markers in mapped host memory (development): 8 ints per CTA
};
__device__ __forceinline__ void mark(const CatArgs& a, int slot, int v) {
  if (a.dbg) {
    a.dbg[blockIdx.x * 8 + slot] = v;
    __threadfence_system();
  }
}

// ---- column assignment: users -> 128-column groups that never split a user (next-fit inside chunks of 64 users)
__global__ void __launch_bounds__(32) cat_assign_kernel(int* __restrict__ ucol, int* __restrict__ n_groups,
                                                        const int2* __restrict__ useg, int B) {
  if (threadIdx.x != 0) return;
  const int u0 = blockIdx.x * 64, u1 = min(B, u0 + 64);
  int g = 0, cur = 0;
  for (int u = u0; u < u1; ++u) {
    const int n = min(useg[u].y, CT);
    if (cur + n > CT) { ++g; cur = 0; }
    ucol[u] = g * CT + cur;     // group index relative to the chunk for now
    cur += n;
  }
  const int base = atomicAdd(n_groups, g + 1);
  for (int u = u0; u < u1; ++u) ucol[u] += base * CT;
}

constexpr float kMasked = -1.0e30f;   // context term of a masked / unused column: its softmax weight underflows to 0
// every column starts as unused: user -1, masked in both heads, zero value fold
__global__ void __launch_bounds__(256) cat_clear_meta_kernel(float* __restrict__ meta, long long n_cols) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_cols) return;
  float* mp = meta + (i / CT) * META_WORDS + (i % CT);
  mp[MP_USER * CT] = __int_as_float(-1);
  mp[MP_KC0 * CT] = kMasked;
  mp[MP_KC1 * CT] = kMasked;
  mp[MP_U0 * CT] = 0.f;
  mp[MP_U1 * CT] = 0.f;
}

struct FillArgs {
  float* Bt;
  float* meta;
  const int* ucol;
  const int2* useg;
  const int *row_src, *n_rows;
  const float *Kd, *Vd;       // ca: decoder keys / values, fp32 rows (row stride ld)
  int ld;
  const float *wf, *McQ;      // ca
  const float* ctx_user;      // [B, C]
  const float* PE;            // dot: encoded profile rows
  const float* Mc;            // dot
  const float* mcw;
  const float* y_pos;         // per user: probability of the positive
  const int* pos_item;
  int* n_groups_out;          // dot: receives ceil(B / 128)
  int B, C;
  float sc;                   // ca: log2(e) / sqrt(dh)
};
// ca: one warp per packed row (key)
__global__ void __launch_bounds__(256) cat_fill_ca_kernel(const FillArgs a) {
  const int lane = threadIdx.x & 31;
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (r >= *a.n_rows) return;
  const int src = a.row_src[r];
  const int u = (src >> 8) & 0x7fffff;
  const int2 sg = a.useg[u];
  const int j = (int)(r - sg.x);
  if (j >= CT) return;                           // (L <= 128 is required by the entry point)
  const int col = a.ucol[u] + j, g = col / CT, c = col % CT;
  // lane l < 16: float4 l of the key row -> head l / 8, chunk l % 8
  float4 kq = make_float4(0.f, 0.f, 0.f, 0.f), vq = kq;
  if (lane < 16) {
    kq = __ldg(reinterpret_cast<const float4*>(a.Kd + r * a.ld) + lane);
    vq = __ldg(reinterpret_cast<const float4*>(a.Vd + r * a.ld) + lane);
  }
  // context term of the key and the value fold, per head (lanes 0..7 head 0, 8..15 head 1)
  float kc = 0.f, uu = 0.f;
  if (lane < 16) {
    const float* cu = a.ctx_user + (long long)u * a.C;
    const float kk[4] = {kq.x, kq.y, kq.z, kq.w}, vv[4] = {vq.x, vq.y, vq.z, vq.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = 4 * lane + e;
      float q = 0.f;
      for (int k = 0; k < a.C; ++k) q = fmaf(__ldg(a.McQ + n * 8 + k), __ldg(cu + k), q);
      kc = fmaf(kk[e], q, kc);
      uu = fmaf(vv[e], __ldg(a.wf + n), uu);
    }
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) {
    kc += __shfl_xor_sync(kFull, kc, o);
    uu += __shfl_xor_sync(kFull, uu, o);
  }
  const float kc1 = __shfl_sync(kFull, kc, 8), u1 = __shfl_sync(kFull, uu, 8);
  if (lane < 16) {
    const float4 hs = make_float4(kq.x * a.sc, kq.y * a.sc, kq.z * a.sc, kq.w * a.sc);
    const float4 hi = make_float4(umma::tf32_hi(hs.x), umma::tf32_hi(hs.y), umma::tf32_hi(hs.z), umma::tf32_hi(hs.w));
    const float4 lo = make_float4(hs.x - hi.x, hs.y - hi.y, hs.z - hi.z, hs.w - hi.w);
    float4* base = reinterpret_cast<float4*>(a.Bt + (long long)g * GROUP_FLOATS);
    base[(0 * 16 + lane) * CT + c] = hi;
    base[(1 * 16 + lane) * CT + c] = lo;
  }
  if (lane == 0) {
    float* mp = a.meta + (long long)g * META_WORDS + c;
    const bool pad = src < 0;                    // position L-1 of a short window: a padding key (masked)
    const bool last = j == min(sg.y, CT) - 1;
    mp[MP_USER * CT] = __int_as_float(u | (last ? (1 << 30) : 0));
    mp[MP_KC0 * CT] = pad ? kMasked : kc * a.sc;
    mp[MP_KC1 * CT] = pad ? kMasked : kc1 * a.sc;
    mp[MP_U0 * CT] = uu;
    mp[MP_U1 * CT] = u1;
    if (last) {
      float w = 0.f;
      for (int k = 0; k < a.C; ++k) w = fmaf(__ldg(a.mcw + k), __ldg(a.ctx_user + (long long)u * a.C + k), w);
      mp[MP_YPOS * CT] = a.y_pos[u];
      mp[MP_PITEM * CT] = __int_as_float(a.pos_item[u]);
      mp[MP_CW * CT] = w;
    }
  }
}
// dot: one warp per user; column = user
__global__ void __launch_bounds__(256) cat_fill_dot_kernel(const FillArgs a) {
  const int lane = threadIdx.x & 31;
  const int u = (int)(((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (u >= a.B) return;
  const int2 sg = a.useg[u];
  const long long r = (long long)sg.x + sg.y - 1;      // position L-1 is the segment's last row
  const int g = u / CT, c = u % CT;
  float4 pq = make_float4(0.f, 0.f, 0.f, 0.f);
  float kc = 0.f;
  if (lane < 16) {
    pq = __ldg(reinterpret_cast<const float4*>(a.PE + r * 64) + lane);
    const float* cu = a.ctx_user + (long long)u * a.C;
    const float pp[4] = {pq.x, pq.y, pq.z, pq.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int n = 4 * lane + e;
      float q = 0.f;
      for (int k = 0; k < a.C; ++k) q = fmaf(__ldg(a.Mc + n * 8 + k), __ldg(cu + k), q);
      kc = fmaf(pp[e], q, kc);
    }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) kc += __shfl_xor_sync(kFull, kc, o);
  if (lane < 16) {
    const float4 hi = make_float4(umma::tf32_hi(pq.x), umma::tf32_hi(pq.y), umma::tf32_hi(pq.z), umma::tf32_hi(pq.w));
    const float4 lo = make_float4(pq.x - hi.x, pq.y - hi.y, pq.z - hi.z, pq.w - hi.w);
    float4* base = reinterpret_cast<float4*>(a.Bt + (long long)g * GROUP_FLOATS);
    base[(0 * 16 + lane) * CT + c] = hi;
    base[(1 * 16 + lane) * CT + c] = lo;
  }
  if (lane == 0) {
    float* mp = a.meta + (long long)g * META_WORDS + c;
    mp[MP_USER * CT] = __int_as_float(u | (1 << 30));
    mp[MP_KC0 * CT] = kc;
    mp[MP_YPOS * CT] = a.y_pos[u];
    mp[MP_PITEM * CT] = __int_as_float(a.pos_item[u]);
    if (u == 0) *a.n_groups_out = (a.B + CT - 1) / CT;
  }
}

struct CatSmem {
  float a[2][16 * A_LBO / 4];          // item tile: [hi | lo][16 chunks][128 rows (+pad)][4]
  float b[2][GROUP_FLOATS];            // two column-group stages
  alignas(16) float meta[2][META_WORDS];
  uint64_t b_full[2], b_empty[2], acc_full[2], acc_empty[2];
  uint32_t tmem_slot;
};

// DEC: 0 = dot (one column per user, K = 64), 1 = cross-attention (two heads, K = 32 each)
template <int DEC>
__global__ void __launch_bounds__(CAT_THREADS, 1) catalog_tc_kernel(const CatArgs a) {
  extern __shared__ __align__(128) unsigned char cat_raw[];
  CatSmem& s = *reinterpret_cast<CatSmem*>(cat_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tile = blockIdx.x / a.parts, part = blockIdx.x % a.parts;
  const int n_groups = *a.n_groups;
  const int per = (n_groups + a.parts - 1) / a.parts;
  const int g0 = part * per, g1 = min(n_groups, g0 + per);
  if (g0 >= g1) return;
  const int item0 = a.item_lo + tile * CT;

  if (threadIdx.x == 0) {
    for (int i = 0; i < 2; ++i) {
      umma::mbar_init(&s.b_full[i], 1); umma::mbar_init(&s.b_empty[i], 1);
      umma::mbar_init(&s.acc_full[i], 1); umma::mbar_init(&s.acc_empty[i], 4);
    }
  }
  if (threadIdx.x == 0) mark(a, 0, 1 + (g1 - g0));
  if (warp == 1) umma::tmem_alloc(&s.tmem_slot, 512);
  if (threadIdx.x == 32) mark(a, 1, 1);
  // item tile -> hi / lo planes (all threads; rows past the shard are zero)
  for (int i = threadIdx.x; i < CT * 16; i += CAT_THREADS) {
    const int r = i / 16, q = i % 16;
    const int item = item0 + r;
    float4 x = make_float4(0.f, 0.f, 0.f, 0.f);
    if (item < a.item_lo + a.n_items) x = __ldg(reinterpret_cast<const float4*>(a.TA + (long long)item * 64) + q);
    const float4 hi = make_float4(umma::tf32_hi(x.x), umma::tf32_hi(x.y), umma::tf32_hi(x.z), umma::tf32_hi(x.w));
    const float4 lo = make_float4(x.x - hi.x, x.y - hi.y, x.z - hi.z, x.w - hi.w);
    *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s.a[0]) + q * A_LBO + r * 16) = hi;
    *reinterpret_cast<float4*>(reinterpret_cast<unsigned char*>(s.a[1]) + q * A_LBO + r * 16) = lo;
  }
  umma::fence_smem_to_async();
  umma::fence_before_sync();
  __syncthreads();
  umma::fence_after_sync();
  const uint32_t tmem0 = s.tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
      for (int g = g0, it = 0; g < g1; ++g, ++it) {
        const int st = it & 1, ph = (it >> 1) & 1;
        // the stage is free when the MMAs have read its operands AND the epilogue has read its column metadata
        if (!rows::wait_or_flag(&s.b_empty[st], ph ^ 1, a.status, 4)) break;
        if (!rows::wait_or_flag(&s.acc_empty[st], ph ^ 1, a.status, 4)) break;
        rows::mbar_expect_tx(&s.b_full[st], GROUP_FLOATS * 4 + META_WORDS * 4);
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.Bt + (long long)g * GROUP_FLOATS);
        for (int off = 0; off < GROUP_FLOATS * 4; off += 16384)
          rows::bulk_g2s(reinterpret_cast<unsigned char*>(s.b[st]) + off, src + off, 16384, &s.b_full[st]);
        rows::bulk_g2s(s.meta[st], a.meta + (long long)g * META_WORDS, META_WORDS * 4, &s.b_full[st]);
        mark(a, 2, it + 1);
      }
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc = umma::idesc_tf32(CT);
    constexpr uint32_t b_lbo = CT * 16;
    bool ok = true;
    for (int g = g0, it = 0; g < g1 && ok; ++g, ++it) {
      const int st = it & 1, ph = (it >> 1) & 1;
      ok = rows::wait_or_flag(&s.acc_empty[st], ph ^ 1, a.status, 8);
      ok = ok && rows::wait_or_flag(&s.b_full[st], ph, a.status, 16);
      umma::fence_after_sync();
      if (umma::elect_one()) {
        const uint32_t a_hi = umma::smem_u32(s.a[0]), a_lo = umma::smem_u32(s.a[1]);
        const uint32_t b_hi = umma::smem_u32(s.b[st]), b_lo = b_hi + 16 * b_lbo;
        constexpr int NH = DEC == 1 ? 2 : 1, KS = DEC == 1 ? 4 : 8;   // heads, K = 8 steps per head
#pragma unroll
        for (int h = 0; h < NH; ++h) {
          const uint32_t d = tmem0 + st * 256 + h * CT;
#pragma unroll
          for (int p = 0; p < 3; ++p) {
            const uint32_t ab = (p == 1 ? a_lo : a_hi) + h * 8 * A_LBO;
            const uint32_t bb = (p == 2 ? b_lo : b_hi) + h * 8 * b_lbo;
#pragma unroll
            for (int ks = 0; ks < KS; ++ks)
              umma::mma_tf32(d, umma::smem_desc(ab + ks * 2 * A_LBO, A_LBO, 128), umma::smem_desc(bb + ks * 2 * b_lbo, b_lbo, 128),
                             idesc, !(p == 0 && ks == 0));
          }
        }
        umma::commit(&s.b_empty[st]);
        umma::commit(&s.acc_full[st]);
        mark(a, 3, it + 1);
      }
      __syncwarp();
    }
  } else {
    const int wg = (warp - 2) >> 2, quarter = warp & 3;
    const int row = 32 * quarter + lane;
    const int item = item0 + row;
    const bool live = item < a.item_lo + a.n_items;
    const float twv = (DEC == 1 && live && a.residual_ca) ? __ldg(a.tw + item) : 0.f;
    const float bfv = DEC == 1 ? __ldg(a.dbf) : 0.f;
    bool ok = true;
    for (int g = g0, it = 0; g < g1 && ok; ++g, ++it) {
      const int st = it & 1, ph = (it >> 1) & 1;
      if (st != wg) continue;
      ok = rows::wait_or_flag(&s.acc_full[st], ph, a.status, 32);
      umma::fence_after_sync();
      const uint32_t tb = tmem0 + ((uint32_t)(32 * quarter) << 16) + st * 256;
      const float4* mt = reinterpret_cast<const float4*>(s.meta[st]);
      // running softmax state of the open user, per head.  Masked / unused columns carry kc = -1e30: their weight
      // underflows to exactly 0 as soon as a real key has been seen, and a user with no real key at all keeps
      // m <= -1e29, which the finalisation turns into the all-masked attention row 0 (src/carca.py:256) — so the
      // per-column math needs no branch; the only (warp-uniform) branch is the user's last column.
      float m0 = kMasked, z0 = 0.f, d0 = 0.f, m1 = kMasked, z1 = 0.f, d1 = 0.f;
#pragma unroll 1
      for (int c0 = 0; c0 < CT; c0 += 16) {
        float v0[16], v1[16];
        umma::tmem_ld_1x16(tb + c0, v0);
        if (DEC == 1) umma::tmem_ld_1x16(tb + CT + c0, v1);
#pragma unroll
        for (int e4 = 0; e4 < 16; e4 += 4) {
          const int q4 = (c0 + e4) / 4;                    // metadata of 4 columns: broadcast reads
          const float4 fu = mt[MP_USER * (CT / 4) + q4], fk0 = mt[MP_KC0 * (CT / 4) + q4];
          const float uu[4] = {fu.x, fu.y, fu.z, fu.w}, k0[4] = {fk0.x, fk0.y, fk0.z, fk0.w};
          float k1[4], a0[4], a1[4];
          if (DEC == 1) {
            const float4 fk1 = mt[MP_KC1 * (CT / 4) + q4], fa0 = mt[MP_U0 * (CT / 4) + q4], fa1 = mt[MP_U1 * (CT / 4) + q4];
            k1[0] = fk1.x; k1[1] = fk1.y; k1[2] = fk1.z; k1[3] = fk1.w;
            a0[0] = fa0.x; a0[1] = fa0.y; a0[2] = fa0.z; a0[3] = fa0.w;
            a1[0] = fa1.x; a1[1] = fa1.y; a1[2] = fa1.z; a1[3] = fa1.w;
          }
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int e = e4 + q;
            const int uw = __float_as_int(uu[q]);
            float logit;
            if (DEC == 1) {
              const float s0 = v0[e] + k0[q], s1 = v1[e] + k1[q];
              const float n0 = fmaxf(m0, s0), n1 = fmaxf(m1, s1);
              const float c0f = rows::ex2f(m0 - n0), p0 = rows::ex2f(s0 - n0);
              const float c1f = rows::ex2f(m1 - n1), p1 = rows::ex2f(s1 - n1);
              z0 = fmaf(z0, c0f, p0); d0 = fmaf(d0, c0f, p0 * a0[q]); m0 = n0;
              z1 = fmaf(z1, c1f, p1); d1 = fmaf(d1, c1f, p1 * a1[q]); m1 = n1;
              if (uw < (1 << 30)) continue;                  // not a user's last column (unused columns: uw = -1)
              logit = (m0 > -1.0e29f ? d0 / z0 : 0.f) + (m1 > -1.0e29f ? d1 / z1 : 0.f) + bfv;
              if (a.residual_ca) logit += twv + s.meta[st][MP_CW * CT + c0 + e];
              m0 = m1 = kMasked; z0 = z1 = d0 = d1 = 0.f;
            } else {
              if (uw < 0) continue;
              logit = v0[e] + k0[q];
            }
            const float y = 1.0f / (1.0f + expf(-logit));
            const float yp = s.meta[st][MP_YPOS * CT + c0 + e];
            const int pi = __float_as_int(s.meta[st][MP_PITEM * CT + c0 + e]);
            const bool before = live && item != pi && (y > yp || (y == yp && item < pi));
            const unsigned bal = __ballot_sync(kFull, before);
            if (lane == 0 && bal) atomicAdd(a.counts + (uw & 0x3fffffff), __popc(bal));
          }
        }
      }
      umma::fence_before_sync();
      __syncwarp();
      if (lane == 0) rows::mbar_arrive(&s.acc_empty[st]);
      if (lane == 0 && quarter == 0) mark(a, 4 + wg, it + 1);
    }
  }
  umma::fence_before_sync();
  if (threadIdx.x % 32 == 0) mark(a, 6, 0);
  __syncthreads();
  if (warp == 1) umma::tmem_free(tmem0, 512);
  if (threadIdx.x == 32) mark(a, 7, 1);
}

}  // namespace cat
}  // namespace carca
#endif
