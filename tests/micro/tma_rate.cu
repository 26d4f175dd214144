// Microbenchmark (diagnostic): TMA load throughput / latency per SM as a function of the bytes in flight.
// Every SM runs one CTA with a ring of `slots` 16 KB shared-memory slots; a producer thread keeps the ring full
// with [128 rows x 128 B] SWIZZLE_128B boxes (the K-major operand slices of the GEMM kernels), a consumer thread
// releases each slot `hold` cycles after it has landed.  mode 0: all CTAs re-read one 1.5 MB matrix (L2 hits);
// mode 1: every CTA streams its own region (HBM).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_rate tma_rate.cu
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda.h>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t n) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(n) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t par) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(s32(b)), "r"(par) : "memory");
}
__device__ __forceinline__ void tma3(void* dst, const CUtensorMap* t, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
               ::"r"(s32(dst)), "l"((uint64_t)t), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}

__global__ void __launch_bounds__(64, 1) k(const __grid_constant__ CUtensorMap map, int slots, int loads, int hold, int mode,
                                           int rows_total, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(sm + 12 * 16384);
  uint64_t* empty = full + 12;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int ksteps = 12;                     // 768 columns = 12 boxes of 64
  const int row_tiles = rows_total / 128;
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int i = 0; i < loads; ++i) {
      mbar_wait(&empty[slot], ph ^ 1);
      mbar_expect(&full[slot], 16384);
      int tile = mode == 0 ? (i / ksteps) % 8 : ((blockIdx.x * 64 + i / ksteps) % row_tiles);
      tma3(sm + slot * 16384, &map, &full[slot], (i % ksteps) * 64, tile * 128, 0);
      if (++slot == slots) { slot = 0; ph ^= 1; }
    }
  } else if (threadIdx.x == 32) {
    int slot = 0; uint32_t ph = 0;
    for (int i = 0; i < loads; ++i) {
      mbar_wait(&full[slot], ph);
      if (hold > 0) { long long t = clock64(); while (clock64() - t < hold) {} }
      mbar_arrive(&empty[slot]);
      if (++slot == slots) { slot = 0; ph ^= 1; }
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

// mode 2: the access pattern of the fused forward's phase 1 without any MMA: a "k-step" = one [128 x 64] slice of a text
// tile (1024-row matrix shared by all CTAs, tile = CTA % 8) + one [96 x 64] slice of an image chunk (image = CTA / 8 +
// 18 * item, 576 rows per image); the consumer waits for both slots, holds them `hold` cycles, releases both.
__global__ void __launch_bounds__(64, 1) k2(const __grid_constant__ CUtensorMap mapT, const __grid_constant__ CUtensorMap mapV,
                                            int slots, int items, int hold, int images, long long* cyc) {
  extern __shared__ uint8_t raw[];
  uint8_t* sm = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  uint64_t* full = (uint64_t*)(sm + 12 * 16384);
  uint64_t* empty = full + 12;
  if (threadIdx.x == 0) {
    for (int i = 0; i < 12; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  const int b = blockIdx.x;
  if (threadIdx.x == 0) {
    int slot = 0; uint32_t ph = 0;
    for (int it = 0; it < items; ++it) {
      const int img = (b / 8 + 18 * it) % images;
      for (int sc = 0; sc < 3; ++sc)
        for (int ks = 0; ks < 12; ++ks) {
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect(&full[slot], 16384);
          tma3(sm + slot * 16384, &mapT, &full[slot], ks * 64, (b % 8) * 128, 0);
          if (++slot == slots) { slot = 0; ph ^= 1; }
          mbar_wait(&empty[slot], ph ^ 1);
          mbar_expect(&full[slot], 96 * 128);
          tma3(sm + slot * 16384, &mapV, &full[slot], ks * 64, sc * 192 + (b & 1) * 96, img);
          if (++slot == slots) { slot = 0; ph ^= 1; }
        }
    }
  } else if (threadIdx.x == 32) {
    int slot = 0; uint32_t ph = 0;
    for (int i = 0; i < items * 36; ++i) {
      mbar_wait(&full[slot], ph);
      const int s0 = slot;
      if (++slot == slots) { slot = 0; ph ^= 1; }
      mbar_wait(&full[slot], ph);
      const int s1 = slot;
      if (++slot == slots) { slot = 0; ph ^= 1; }
      if (hold > 0) { long long t = clock64(); while (clock64() - t < hold) {} }
      mbar_arrive(&empty[s0]);
      mbar_arrive(&empty[s1]);
    }
    cyc[blockIdx.x] = clock64() - t0;
  }
}

int main() {
  const int rows = 148 * 64 * 128;          // 1.2 M rows x 768 bf16 = 1.86 GB
  const size_t bytes = (size_t)rows * 768 * 2;
  void* buf; CK(cudaMalloc(&buf, bytes)); CK(cudaMemset(buf, 1, bytes));
  long long* cyc; CK(cudaMalloc(&cyc, 148 * 8));
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  auto enc = (CUresult(*)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                          const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill))fn;
  CUtensorMap map;
  cuuint64_t dims[3] = {768, (cuuint64_t)rows, 1}, strides[2] = {1536, (cuuint64_t)bytes};
  cuuint32_t box[3] = {64, 128, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 1; }
  const int smem = 12 * 16384 + 256 + 1024;
  CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const int loads = 12 * 64;
  for (int mode = 0; mode < 2; ++mode)
    for (int hold : {0, 192})
      for (int slots : {1, 2, 3, 4, 5, 6, 8, 10, 12}) {
        k<<<148, 64, smem>>>(map, slots, loads, hold, mode, rows, cyc);
        k<<<148, 64, smem>>>(map, slots, loads, hold, mode, rows, cyc);
        CK(cudaDeviceSynchronize());
        long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
        double avg = 0; long long mx = 0;
        for (int i = 0; i < 148; ++i) { avg += h[i]; mx = h[i] > mx ? h[i] : mx; }
        avg /= 148;
        printf("mode %d (%s) hold %3d slots %2d: %.0f cyc per 16 KB load (max CTA %.0f) -> %.1f B/clk/SM; implied latency %.0f cyc\n", mode,
               mode == 0 ? "L2 hits" : "HBM stream", hold, slots, avg / loads, (double)mx / loads, 16384.0 * loads / avg,
               avg / loads * slots);
      }
  // ---- mode 2
  {
    const int images = 1024;
    CUtensorMap mT, mV;
    cuuint64_t dT[3] = {768, 1024, 1}, sT[2] = {1536, 1536 * 1024};
    cuuint32_t bT[3] = {64, 128, 1};
    enc(&mT, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, buf, dT, sT, bT, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    cuuint64_t dV[3] = {768, 576, (cuuint64_t)images}, sV[2] = {1536, 1536 * 576};
    cuuint32_t bV[3] = {64, 96, 1};
    enc(&mV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (char*)buf + (64 << 20), dV, sV, bV, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    CK(cudaFuncSetAttribute(k2, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    for (int imgs : {64, 1024})
      for (int hold : {0, 384})
        for (int slots : {4, 5, 6, 8, 10, 12}) {
          k2<<<148, 64, smem>>>(mT, mV, slots, 8, hold, imgs, cyc);
          k2<<<148, 64, smem>>>(mT, mV, slots, 8, hold, imgs, cyc);
          CK(cudaDeviceSynchronize());
          long long h[148]; CK(cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost));
          double avg = 0;
          for (int i = 0; i < 148; ++i) avg += h[i];
          avg /= 148;
          printf("fused-like phase 1, %4d images (%s), hold %3d, slots %2d: %.0f cyc per k-step (28 KB) -> %.1f B/clk/SM\n", imgs,
                 imgs <= 64 ? "L2-resident" : "HBM", hold, slots, avg / (8 * 36), 28672.0 * 8 * 36 / avg);
        }
  }
  return 0;
}
