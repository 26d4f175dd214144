// Host-side helpers shared by the C-ABI translation units: error reporting, device gate, TMA tensor maps,
// engine launch.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <atomic>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>

#include "../../include/clipk.h"
#include "gemm_engine.cuh"
#include "gemm2_engine.cuh"

namespace clipk {

void set_error(const char* fmt, ...);          // thread-local message returned by clipk_last_error()
int check_device();                            // 0 if the current device is sm_100 (B200), else CLIPK_ERR_ARCH
int sm_count();                                // SM count of the current device
void count_launches(int n);
bool pdl_enabled();                            // CLIPK_PDL=0 disables programmatic dependent launch of the CTA-pair engine
// Scope guard: no programmatic dependent launch on this thread while it lives.  Used where several stream lanes feed
// the GPU: a dependent grid's CTAs would take freed SMs only to wait, ahead of the other lane's runnable CTAs
// (measured: 64-image groups on 2 lanes 8.0-8.4 ms without, 8.6 ms with).
struct PdlBlock {
  bool active;
  explicit PdlBlock(bool block);
  ~PdlBlock();
  PdlBlock(const PdlBlock&) = delete;
  PdlBlock& operator=(const PdlBlock&) = delete;
};
bool trace_enabled();                          // CLIPK_TRACE=1: record CUDA events around every engine launch
void trace_begin(const char* name, cudaStream_t st);
void trace_end(cudaStream_t st);                    // bump the library-wide kernel-launch counter (clipk_launch_count)

#define CLIPK_CHECK_CUDA(expr)                                                                       \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) {                                                                         \
      clipk::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__);  \
      return CLIPK_ERR_CUDA;                                                                         \
    }                                                                                                \
  } while (0)

#define CLIPK_REQUIRE(cond, ...)            \
  do {                                      \
    if (!(cond)) {                          \
      clipk::set_error(__VA_ARGS__);        \
      return CLIPK_ERR_INVALID;             \
    }                                       \
  } while (0)

#define CLIPK_TRY(...)          \
  do {                          \
    int _r = (__VA_ARGS__);     \
    if (_r != 0) return _r;     \
  } while (0)

// 3-D bf16 tensor map with SWIZZLE_128B and a (64, box_rows, 1) box.
//   dims    = (inner, rows, batch) extents in elements (exact: out-of-bounds box parts are zero-filled)
//   strides = byte strides of dims 1 and 2 (multiples of 16)
int make_tmap_bf16(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                   uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_rows);

// same with an explicit inner box extent: 64 elements (SWIZZLE_128B) or 32 elements (SWIZZLE_64B)
int make_tmap_bf16_box(CUtensorMap* out, const void* base, uint64_t inner, uint64_t rows, uint64_t batch,
                       uint64_t row_stride_bytes, uint64_t batch_stride_bytes, uint32_t box_inner, uint32_t box_rows);

struct OperandDesc {
  const void* ptr = nullptr;
  bool mn_major = false;     // false: [rows][k] (k contiguous); true: [k][rows] (rows contiguous)
  int64_t rows = 0;          // M (for A) or N (for B) extent
  int64_t k = 0;             // reduction extent of ONE sub-batch
  int64_t ld = 0;            // leading dimension in elements (stride of the non-contiguous matrix dim)
  int64_t batch = 1;         // extent of the 3rd (batch) dim of the tensor
  int64_t batch_stride = 0;  // elements
  int bmul = 0;              // batch coordinate = b * bmul + sub * smul
  int smul = 0;
  int sub_per_batch = 0;     // split-K: sub = b * sub_per_batch + (kstep / ksub)
  int sub_total = 0;         // >0: uneven split, batch b covers sub-batches [b*spb, min((b+1)*spb, sub_total))
  int reverse = 0;           // (read from operand A of pair 0) walk batches downwards
  int64_t k_batch_offset = 0;  // (operand A of each pair; CTA-pair engine) batch b covers k in [b * off, b * off + ksteps * 64)
                               // of ONE long reduction dim: describe the operands with their true k extent, batch = 1
};

template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm(const OperandDesc* a, const OperandDesc* b, int num_pairs, const int* ksteps, const int* ksub,
                int M, int N, int batches, const typename Epi::Params& ep, cudaStream_t stream) {
  constexpr bool kDual = eng::epi_dual<Epi>::value;
  constexpr bool kTmaOut = eng::epi_tma_out<Epi>::value;
  using L = eng::SmemLayout<BN, kDual, kTmaOut>;
  eng::OperandMaps maps;
  memset(&maps, 0, sizeof(maps));
  eng::Problem pb;
  memset(&pb, 0, sizeof(pb));
  pb.M = M;
  pb.N = N;
  pb.batches = batches;
  pb.tiles_m = (M + eng::BM - 1) / eng::BM;
  pb.tiles_n = (N + BN - 1) / BN;
  pb.num_pairs = num_pairs;
  pb.reverse = a[0].reverse;
  for (int q = 0; q < num_pairs; ++q) {
    pb.ksteps[q] = ksteps[q];
    pb.ksub[q] = ksub[q] > 0 ? ksub[q] : (ksteps[q] > 0 ? ksteps[q] : 1);
    pb.a_bmul[q] = a[q].bmul; pb.a_smul[q] = a[q].smul;
    pb.b_bmul[q] = b[q].bmul; pb.b_smul[q] = b[q].smul;
    pb.sub_per_batch[q] = a[q].sub_per_batch;
    pb.sub_total[q] = a[q].sub_total;
    if (!A_MN) {
      CLIPK_TRY(make_tmap_bf16(&maps.a[q], a[q].ptr, a[q].k, a[q].rows, a[q].batch, a[q].ld * 2, a[q].batch_stride * 2, eng::BM));
    } else {
      CLIPK_TRY(make_tmap_bf16(&maps.a[q], a[q].ptr, a[q].rows, a[q].k, a[q].batch, a[q].ld * 2, a[q].batch_stride * 2, 64));
    }
    if (!B_MN) {
      CLIPK_TRY(make_tmap_bf16(&maps.b[q], b[q].ptr, b[q].k, b[q].rows, b[q].batch, b[q].ld * 2, b[q].batch_stride * 2, BN));
    } else {
      CLIPK_TRY(make_tmap_bf16(&maps.b[q], b[q].ptr, b[q].rows, b[q].k, b[q].batch, b[q].ld * 2, b[q].batch_stride * 2, 64));
    }
  }
  if constexpr (kDual) {      // second A operand (shares pair 0's B): a[1] describes it, num_pairs stays 1
    pb.a_bmul[1] = a[1].bmul;
    CLIPK_TRY(make_tmap_bf16(&maps.a[1], a[1].ptr, a[1].k, a[1].rows, a[1].batch, a[1].ld * 2, a[1].batch_stride * 2, eng::BM));
  }
  if constexpr (kTmaOut) {
    const eng::OutDesc& o = ep.out;
    CLIPK_TRY(make_tmap_bf16(&maps.out, o.ptr, (uint64_t)o.cols, (uint64_t)o.rows, (uint64_t)o.batches, (uint64_t)o.ld * 2,
                             (uint64_t)o.stride * 2, eng::BM));
  }
  const int total = pb.batches * pb.tiles_m * pb.tiles_n;
  if (total <= 0) return 0;
  auto kern = eng::gemm_kernel<BN, A_MN, B_MN, Epi>;
  static std::atomic<uint64_t> attr_mask{0};   // per-device: the attribute is per context
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const int grid = total < sm_count() ? total : sm_count();
  const bool tr = trace_enabled();
  if (tr) trace_begin(__PRETTY_FUNCTION__, stream);
  kern<<<grid, eng::kThreads, L::kTotal, stream>>>(maps, pb, ep);
  if (tr) trace_end(stream);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// CTA-pair engine (gemm2_engine.cuh): same operand / epilogue description as launch_gemm; a cluster of two CTAs
// computes one 256 x BN tile (each CTA: 128 rows of A, BN/2 rows of B).
template <int BN, bool A_MN, bool B_MN, class Epi>
int launch_gemm2(const OperandDesc* a, const OperandDesc* b, int num_pairs, const int* ksteps, const int* ksub,
                 int M, int N, int batches, const typename Epi::Params& ep, cudaStream_t stream) {
  constexpr bool kDual = eng::epi_dual<Epi>::value;
  constexpr bool kTmaOut = eng::epi_tma_out<Epi>::value;
  constexpr bool kTmaOut2 = eng::epi_tma_out2<Epi>::value;
  constexpr bool kChunkIn = eng::epi_chunk_in<Epi>::value;
  constexpr int kEpiWarps = eng::epi_warps<Epi>::value;
  using L = eng2::SmemLayout<BN, kDual, kTmaOut, kTmaOut2, kChunkIn, kEpiWarps>;
  eng::OperandMaps maps;
  memset(&maps, 0, sizeof(maps));
  eng::Problem pb;
  memset(&pb, 0, sizeof(pb));
  pb.M = M;
  pb.N = N;
  pb.batches = batches;
  pb.tiles_m = (M + 2 * eng2::BM - 1) / (2 * eng2::BM);
  pb.tiles_n = (N + BN - 1) / BN;
  pb.num_pairs = num_pairs;
  pb.reverse = a[0].reverse;
  for (int q = 0; q < num_pairs; ++q) {
    pb.ksteps[q] = ksteps[q];
    pb.ksub[q] = ksub[q] > 0 ? ksub[q] : (ksteps[q] > 0 ? ksteps[q] : 1);
    pb.a_bmul[q] = a[q].bmul; pb.a_smul[q] = a[q].smul;
    pb.b_bmul[q] = b[q].bmul; pb.b_smul[q] = b[q].smul;
    pb.sub_per_batch[q] = a[q].sub_per_batch;
    pb.sub_total[q] = a[q].sub_total;
    pb.k_boff[q] = (int)a[q].k_batch_offset;
    if (!A_MN) {
      CLIPK_TRY(make_tmap_bf16(&maps.a[q], a[q].ptr, a[q].k, a[q].rows, a[q].batch, a[q].ld * 2, a[q].batch_stride * 2, eng2::BM));
    } else {
      CLIPK_TRY(make_tmap_bf16(&maps.a[q], a[q].ptr, a[q].rows, a[q].k, a[q].batch, a[q].ld * 2, a[q].batch_stride * 2, 64));
    }
    if (!B_MN) {
      CLIPK_TRY(make_tmap_bf16(&maps.b[q], b[q].ptr, b[q].k, b[q].rows, b[q].batch, b[q].ld * 2, b[q].batch_stride * 2, BN / 2));
    } else {
      CLIPK_TRY(make_tmap_bf16(&maps.b[q], b[q].ptr, b[q].rows, b[q].k, b[q].batch, b[q].ld * 2, b[q].batch_stride * 2, 64));
    }
  }
  if constexpr (kDual) {
    pb.a_bmul[1] = a[1].bmul;
    CLIPK_TRY(make_tmap_bf16(&maps.a[1], a[1].ptr, a[1].k, a[1].rows, a[1].batch, a[1].ld * 2, a[1].batch_stride * 2, eng2::BM));
  }
  if constexpr (kTmaOut) {
    const eng::OutDesc& o = ep.out;
    CLIPK_TRY(make_tmap_bf16_box(&maps.out, o.ptr, (uint64_t)o.cols, (uint64_t)o.rows, (uint64_t)o.batches,
                                 (uint64_t)o.ld * 2, (uint64_t)o.stride * 2, 32, 32));   // one [32 x 32] box per epilogue warp
  }
  if constexpr (kTmaOut2) {
    const eng::OutDesc& o = ep.out2;
    CLIPK_TRY(make_tmap_bf16_box(&maps.out2, o.ptr, (uint64_t)o.cols, (uint64_t)o.rows, (uint64_t)o.batches,
                                 (uint64_t)o.ld * 2, (uint64_t)o.stride * 2, 32, 32));
  }
  if constexpr (kChunkIn) {
    const eng::OutDesc& o = ep.in;
    CLIPK_TRY(make_tmap_bf16_box(&maps.in, o.ptr, (uint64_t)o.cols, (uint64_t)o.rows, (uint64_t)o.batches,
                                 (uint64_t)o.ld * 2, (uint64_t)o.stride * 2, 32, 32));
  }
  const int total = pb.batches * pb.tiles_m * pb.tiles_n;
  if (total <= 0) return 0;
  auto kern = eng2::gemm2_kernel<BN, A_MN, B_MN, Epi>;
  static std::atomic<uint64_t> attr_mask{0};
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::kTotal));
    attr_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const int pairs = sm_count() / 2;
  const int grid = 2 * (total < pairs ? total : pairs);
  const bool tr = trace_enabled();
  if (tr) trace_begin(__PRETTY_FUNCTION__, stream);
  if (pdl_enabled()) {
    // programmatic dependent launch: the next engine kernel's prologue overlaps this kernel's tail (see the kernel)
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(eng2::threads_for(kEpiWarps));
    cfg.dynamicSmemBytes = L::kTotal;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    CLIPK_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, maps, pb, ep));
  } else {
    kern<<<grid, eng2::threads_for(kEpiWarps), L::kTotal, stream>>>(maps, pb, ep);
  }
  if (tr) trace_end(stream);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace clipk
