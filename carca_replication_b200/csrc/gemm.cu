// Host launcher for the general fp32 GEMM (see gemm.cuh).
#include "gemm.cuh"
#include "gemm_tc.cuh"

#include <cstdlib>

namespace carca {

static char g_err[512] = {0};
char* err_buf() { return g_err; }

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

static const unsigned long long* g_seed_dev = nullptr;
const unsigned long long* seed_source() { return g_seed_dev; }
void set_seed_source(const unsigned long long* p) { g_seed_dev = p; }

static long long g_launches = 0;
long long launch_count() { return g_launches; }

int check_launch(const char* what) {
  ++g_launches;
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(-3, "%s: CUDA launch failed: %s", what, cudaGetErrorString(e));
  return 0;
}

GemmArgs gemm_defaults(const float* A, const float* B, float* C, int M, int N, int K) {
  GemmArgs g;
  memset(&g, 0, sizeof(g));
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.lda = K; g.ldb = K; g.ldc = N;
  g.transA = 0; g.transB = 1;
  g.alpha = 1.0f;
  g.drop = make_drop(0.f, 0ull, 0u);
  g.ldr = N;
  return g;
}

#ifndef CARCA_EMU
// tcgen05 path (gemm_tc.cuh); CARCA_GEMM=ffma in the environment forces the CUDA-core kernels
static bool tc_enabled() {
  static const bool on = [] {
    const char* e = std::getenv("CARCA_GEMM");
    return !(e && std::strcmp(e, "ffma") == 0);
  }();
  return on;
}

template <int BN>
static int launch_tc(const GemmArgs& g, int tn, int tm, int splits, int vec_a, int vec_b, int vec_c,
                     cudaStream_t stream) {
  const size_t smem = sizeof(GemmTcSmem<BN>);
  auto k = gemm_tc_kernel<BN>;
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return fail(-3, "gemm_tc: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
    configured = true;
  }
  // rows known on the device only (m_dev): at most ~8 CTAs per SM, each loops over its row tiles
  if (g.m_dev) tm = min(tm, max(1, 148 * 8 / max(1, tn * splits)));
  CARCA_LAUNCH(k, dim3(tn, tm, splits), dim3(GT_THREADS), smem, stream, g, vec_a, vec_b, vec_c);
  return check_launch("gemm_tc");
}

static int launch_gemm_tc(GemmArgs g, cudaStream_t stream) {
  constexpr int BK = GT_BK;
  const int BN = g.N > 128 ? 256 : (g.N > 64 ? 128 : 64);   // N <= 256: one column tile, A is read once
  const int tm = ceil_div(g.M, GT_BM), tn = ceil_div(g.N, BN);
  int splits = 1;
  if (g.transA && g.K > 8 * BK) {
    const int want = max(1, (2 * 148) / (tm * tn));
    splits = min(want, ceil_div(g.K, 8 * BK));
  }
  const int kps = ceil_div(ceil_div(max(g.K, 1), splits), BK) * BK;
  splits = ceil_div(max(g.K, 1), kps);
  g.k_per_split = kps;
  if (splits > 1) {
    if (g.bias || g.act || g.R || g.row_mask || g.drop.p > 0.f)
      return fail(-2, "gemm: epilogue options are not available with split-K");
    if (!g.accumulate) {
      if (g.ldc != g.N) return fail(-2, "gemm: split-K needs a dense C");
      cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.M * g.N, stream);
    }
  }
  auto aligned = [](const void* p, long long ld) { return ((uintptr_t)p % 16 == 0) && (ld % 4 == 0); };
  const int vec_a = aligned(g.A, g.lda), vec_b = aligned(g.B, g.ldb), vec_c = aligned(g.C, g.ldc);
  if (BN == 64) return launch_tc<64>(g, tn, tm, splits, vec_a, vec_b, vec_c, stream);
  if (BN == 128) return launch_tc<128>(g, tn, tm, splits, vec_a, vec_b, vec_c, stream);
  return launch_tc<256>(g, tn, tm, splits, vec_a, vec_b, vec_c, stream);
}
#endif

int launch_gemm(GemmArgs g, cudaStream_t stream) {
  if (g.M <= 0 || g.N <= 0) return 0;
#ifndef CARCA_EMU
  if (tc_enabled() && g.M >= 32 && (long long)g.M * g.N * g.K >= (1ll << 21)) return launch_gemm_tc(g, stream);
#endif
  constexpr int BK = 16;
  int BM, BN;
  if (g.N > 64 && g.M > 64) { BM = 128; BN = 128; }
  else if (g.M >= 128 * 148) { BM = 128; BN = 64; }
  else { BM = 64; BN = 64; }
  const int tm = ceil_div(g.M, BM), tn = ceil_div(g.N, BN);
  int splits = 1;
  if (g.transA && g.K > 8 * BK) {
    // weight gradients: tiny [N',K'] output, reduction over every position -> split-K
    const int want = max(1, (2 * 148) / (tm * tn));
    splits = min(want, ceil_div(g.K, 8 * BK));
  }
  int kps = ceil_div(ceil_div(max(g.K, 1), splits), BK) * BK;
  splits = ceil_div(max(g.K, 1), kps);
  g.k_per_split = kps;
  if (splits > 1) {
    if (g.bias || g.act || g.R || g.row_mask || g.drop.p > 0.f)
      return fail(-2, "gemm: epilogue options are not available with split-K");
    if (!g.accumulate) {
      if (g.ldc != g.N) return fail(-2, "gemm: split-K needs a dense C");
      cudaMemsetAsync(g.C, 0, sizeof(float) * (size_t)g.M * g.N, stream);
    }
  }
  dim3 grid(tn, tm, splits), block(256);
  if (BM == 128 && BN == 128) {
    auto k = gemm_kernel<128, 128, 8, 8>;
    CARCA_LAUNCH(k, grid, block, 0, stream, g);
  } else if (BM == 128) {
    auto k = gemm_kernel<128, 64, 8, 4>;
    CARCA_LAUNCH(k, grid, block, 0, stream, g);
  } else {
    auto k = gemm_kernel<64, 64, 4, 4>;
    CARCA_LAUNCH(k, grid, block, 0, stream, g);
  }
  return check_launch("gemm");
}

}  // namespace carca
