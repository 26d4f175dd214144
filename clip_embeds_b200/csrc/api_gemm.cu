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

}  // namespace clipk

extern "C" int clipk_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn,
                               int64_t ldb, int64_t strideB, void* C, int64_t ldc, int64_t strideC, int out_dtype,
                               int M, int N, int K, int batches, float alpha, int accumulate, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::gemm_bf16(A, a_mn, lda, strideA, B, b_mn, ldb, strideB, C, ldc, strideC, out_dtype, M, N, K, batches,
                          alpha, accumulate, static_cast<cudaStream_t>(stream));
}
