// clipk_gemm_bf16: the generic entry of the tcgen05 engine (building block + pipeline unit test).
#include <stdlib.h>

#include "common.cuh"
#include "epilogues.cuh"

namespace clipk {

// engine selection of the generic entry: the CTA-pair engine for anything with more than one 128-row tile of M
// (CLIPK_GEMM_ENGINE=1|2 in the environment forces one, for the pipeline unit tests)
static int pick_engine(int M) {
  static const int forced = [] {
    const char* e = getenv("CLIPK_GEMM_ENGINE");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  return M > eng::BM ? 2 : 1;
}

template <int BN, bool A_MN, bool B_MN>
static int gemm2_dispatch_out(const OperandDesc& a, const OperandDesc& b, int ksteps, void* C, int64_t ldc,
                              int64_t strideC, int out_dtype, int M, int N, int batches, float alpha, int accumulate,
                              cudaStream_t st) {
  int ks[1] = {ksteps};
  int ksub[1] = {ksteps};
  if (out_dtype == CLIPK_BF16 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0 &&
      (batches == 1 || strideC % 8 == 0)) {
    epi::StoreTma::Params ep{{C, ldc, batches > 1 ? strideC : (int64_t)M * ldc, M, N, batches}, alpha};
    return launch_gemm2<BN, A_MN, B_MN, epi::StoreTma>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
  }
  if (out_dtype == CLIPK_BF16) {
    typename epi::Store<true>::Params ep{C, ldc, strideC, M, N, alpha, 0};
    return launch_gemm2<BN, A_MN, B_MN, epi::Store<true>>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
  }
  typename epi::Store<false>::Params ep{C, ldc, strideC, M, N, alpha, accumulate};
  return launch_gemm2<BN, A_MN, B_MN, epi::Store<false>>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
}

template <int BN>
static int gemm2_dispatch_major(const OperandDesc& a, const OperandDesc& b, int ksteps, void* C, int64_t ldc,
                                int64_t strideC, int out_dtype, int M, int N, int batches, float alpha,
                                int accumulate, cudaStream_t st) {
  if constexpr (BN % 128 == 0) {
    if (!a.mn_major && b.mn_major)
      return gemm2_dispatch_out<BN, false, true>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
    if (a.mn_major && b.mn_major)
      return gemm2_dispatch_out<BN, true, true>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  }
  if (!a.mn_major && !b.mn_major)
    return gemm2_dispatch_out<BN, false, false>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  if (a.mn_major && !b.mn_major)
    return gemm2_dispatch_out<BN, true, false>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  set_error("gemm: internal dispatch error (MN-major B needs a 128-multiple tile)");
  return CLIPK_ERR_INVALID;
}

template <int BN, bool A_MN, bool B_MN>
static int gemm_dispatch_out(const OperandDesc& a, const OperandDesc& b, int ksteps, void* C, int64_t ldc,
                             int64_t strideC, int out_dtype, int M, int N, int batches, float alpha, int accumulate,
                             cudaStream_t st) {
  int ks[1] = {ksteps};
  int ksub[1] = {ksteps};
  if (out_dtype == CLIPK_BF16 && (reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 8 == 0 &&
      (batches == 1 || strideC % 8 == 0)) {
    // 16-byte aligned rows: the tile leaves the SM through the engine's TMA-store path
    epi::StoreTma::Params ep{{C, ldc, batches > 1 ? strideC : (int64_t)M * ldc, M, N, batches}, alpha};
    return launch_gemm<BN, A_MN, B_MN, epi::StoreTma>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
  }
  if (out_dtype == CLIPK_BF16) {
    typename epi::Store<true>::Params ep{C, ldc, strideC, M, N, alpha, 0};
    return launch_gemm<BN, A_MN, B_MN, epi::Store<true>>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
  }
  typename epi::Store<false>::Params ep{C, ldc, strideC, M, N, alpha, accumulate};
  return launch_gemm<BN, A_MN, B_MN, epi::Store<false>>(&a, &b, 1, ks, ksub, M, N, batches, ep, st);
}

template <int BN>
static int gemm_dispatch_major(const OperandDesc& a, const OperandDesc& b, int ksteps, void* C, int64_t ldc,
                               int64_t strideC, int out_dtype, int M, int N, int batches, float alpha,
                               int accumulate, cudaStream_t st) {
  if (!a.mn_major && !b.mn_major)
    return gemm_dispatch_out<BN, false, false>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  if (!a.mn_major && b.mn_major)
    return gemm_dispatch_out<BN, false, true>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  if (a.mn_major && !b.mn_major)
    return gemm_dispatch_out<BN, true, false>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  return gemm_dispatch_out<BN, true, true>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
}

int gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn, int64_t ldb,
              int64_t strideB, void* C, int64_t ldc, int64_t strideC, int out_dtype, int M, int N, int K, int batches,
              float alpha, int accumulate, cudaStream_t st) {
  CLIPK_REQUIRE(M > 0 && N > 0 && K > 0 && batches > 0, "gemm: empty problem (M=%d N=%d K=%d batches=%d)", M, N, K, batches);
  CLIPK_REQUIRE(out_dtype == CLIPK_BF16 || out_dtype == CLIPK_F32, "gemm: bad out_dtype %d", out_dtype);
  CLIPK_REQUIRE(!(accumulate && out_dtype != CLIPK_F32), "gemm: accumulate needs fp32 output");
  OperandDesc a, b;
  a.ptr = A; a.mn_major = a_mn != 0; a.rows = M; a.k = K; a.ld = lda; a.batch = batches; a.batch_stride = strideA;
  a.bmul = strideA != 0 ? 1 : 0;
  if (strideA == 0) a.batch = 1;
  b.ptr = B; b.mn_major = b_mn != 0; b.rows = N; b.k = K; b.ld = ldb; b.batch = batches; b.batch_stride = strideB;
  b.bmul = strideB != 0 ? 1 : 0;
  if (strideB == 0) b.batch = 1;
  const int ksteps = (K + eng::BK - 1) / eng::BK;
  if (pick_engine(M) == 2) {
    if (b.mn_major) {       // each CTA's half of an MN-major B tile must be whole 64-wide swizzle groups
      if (N > 128)
        return gemm2_dispatch_major<256>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
      return gemm2_dispatch_major<128>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
    }
    if (N > 192)
      return gemm2_dispatch_major<256>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
    if (N > 128)
      return gemm2_dispatch_major<192>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
    if (N > 64)
      return gemm2_dispatch_major<128>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
    return gemm2_dispatch_major<64>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  }
  if (N > 192)
    return gemm_dispatch_major<256>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  if (N > 128)
    return gemm_dispatch_major<192>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  if (N > 64)
    return gemm_dispatch_major<128>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
  return gemm_dispatch_major<64>(a, b, ksteps, C, ldc, strideC, out_dtype, M, N, batches, alpha, accumulate, st);
}

// ------------------------------------------------------------------------------------------- fp32 on tensor cores
// fp32-accurate GEMM on the bf16 tensor cores: every fp32 operand value is split into three bf16 terms
// x = h + m + l (24 mantissa bits), and the six products hh, hm, mh, mm, hl, lh are folded into ONE engine GEMM by
// concatenating the terms along K:   A' = [h h m m h l] (M x 6Kp),  B' = [h m h m l h] (N x 6Kp),  both K-major.
// The dropped products (ml, lm, ll) are below 2^-24 relative; accumulation is fp32 in TMEM.
constexpr int kSplitTerms = 6;

// X(r, k) = src[r * sr + k * sk]  ->  out [R][6 * Kp] bf16 (K-major), segment order given by `pattern`
// (2 bits per segment: 0 = h, 1 = m, 2 = l).  Tiled 32 x 32 through shared memory so that both the read (along the
// source's contiguous dim) and the write (along k) are coalesced.  k in [K, Kp) is written as zero.
__global__ void split_f32_bf16x3_kernel(const float* __restrict__ src, int64_t sr, int64_t sk, int R, int K, int Kp,
                                        uint32_t pattern, __nv_bfloat16* __restrict__ out) {
  __shared__ float tile[32][33];
  const int r0 = blockIdx.y * 32, k0 = blockIdx.x * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;
  if (sk == 1 || sr != 1) {        // k contiguous (or generic): lanes along k
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + ty + 8 * j, k = k0 + tx;
      tile[ty + 8 * j][tx] = (r < R && k < K) ? src[(int64_t)r * sr + (int64_t)k * sk] : 0.f;
    }
  } else {                         // r contiguous: lanes along r
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = r0 + tx, k = k0 + ty + 8 * j;
      tile[tx][ty + 8 * j] = (r < R && k < K) ? src[(int64_t)r + (int64_t)k * sk] : 0.f;
    }
  }
  __syncthreads();
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int r = r0 + ty + 8 * j, k = k0 + tx;
    if (r >= R || k >= Kp) continue;
    const float x = tile[ty + 8 * j][tx];
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const float x1 = x - __bfloat162float(h);
    const __nv_bfloat16 m = __float2bfloat16_rn(x1);
    const __nv_bfloat16 l = __float2bfloat16_rn(x1 - __bfloat162float(m));
    __nv_bfloat16* o = out + (int64_t)r * (kSplitTerms * (int64_t)Kp) + k;
#pragma unroll
    for (int sgm = 0; sgm < kSplitTerms; ++sgm) {
      const uint32_t which = (pattern >> (2 * sgm)) & 3u;
      o[(int64_t)sgm * Kp] = which == 0 ? h : (which == 1 ? m : l);
    }
  }
}

static inline int kpad(int K) { return (K + eng::BK - 1) / eng::BK * eng::BK; }
static size_t split_ws_bytes(int M, int N, int K) {
  const size_t kp = (size_t)kpad(K);
  auto al = [](size_t b) { return (b + 1023) / 1024 * 1024; };
  return al((size_t)M * kSplitTerms * kp * 2) + al((size_t)N * kSplitTerms * kp * 2);
}

int gemm_f32_split(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                   int64_t ldc, int M, int N, int K, float alpha, int accumulate, void* ws, size_t ws_bytes,
                   cudaStream_t st) {
  CLIPK_REQUIRE(M > 0 && N > 0 && K > 0, "gemm_f32_split: empty problem (M=%d N=%d K=%d)", M, N, K);
  const size_t need = split_ws_bytes(M, N, K);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "gemm_f32_split: workspace too small (%zu < %zu)", ws_bytes, need);
  CLIPK_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 127) == 0, "gemm_f32_split: workspace must be 128-byte aligned");
  const int Kp = kpad(K);
  __nv_bfloat16* As = static_cast<__nv_bfloat16*>(ws);
  const size_t a_bytes = ((size_t)M * kSplitTerms * Kp * 2 + 1023) / 1024 * 1024;
  __nv_bfloat16* Bs = reinterpret_cast<__nv_bfloat16*>(static_cast<char*>(ws) + a_bytes);
  // segments:            0  1  2  3  4  5
  //   A' terms           h  h  m  m  h  l      -> 0 0 1 1 0 2
  //   B' terms           h  m  h  m  l  h      -> 0 1 0 1 2 0
  constexpr uint32_t patA = 0u | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (2u << 10);
  constexpr uint32_t patB = 0u | (1u << 2) | (0u << 4) | (1u << 6) | (2u << 8) | (0u << 10);
  const dim3 blk(32, 8);
  split_f32_bf16x3_kernel<<<dim3(Kp / 32, (M + 31) / 32), blk, 0, st>>>(A, sam, sak, M, K, Kp, patA, As);
  split_f32_bf16x3_kernel<<<dim3(Kp / 32, (N + 31) / 32), blk, 0, st>>>(B, sbn, sbk, N, K, Kp, patB, Bs);
  count_launches(2);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  const int64_t ld = (int64_t)kSplitTerms * Kp;
  return gemm_bf16(As, 0, ld, 0, Bs, 0, ld, 0, C, ldc, 0, CLIPK_F32, M, N, kSplitTerms * Kp, 1, alpha, accumulate, st);
}

}  // namespace clipk

extern "C" size_t clipk_gemm_f32_split_workspace_bytes(int M, int N, int K) {
  if (M <= 0 || N <= 0 || K <= 0) return 0;
  return clipk::split_ws_bytes(M, N, K);
}

extern "C" int clipk_gemm_f32_split(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn,
                                    float* C, int64_t ldc, int M, int N, int K, float alpha, int accumulate,
                                    void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::gemm_f32_split(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, alpha, accumulate, workspace, ws_bytes,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int clipk_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn,
                               int64_t ldb, int64_t strideB, void* C, int64_t ldc, int64_t strideC, int out_dtype,
                               int M, int N, int K, int batches, float alpha, int accumulate, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::gemm_bf16(A, a_mn, lda, strideA, B, b_mn, ldb, strideB, C, ldc, strideC, out_dtype, M, N, K, batches,
                          alpha, accumulate, static_cast<cudaStream_t>(stream));
}
