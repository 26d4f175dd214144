// Fused epilogue functors for eng::gemm_kernel.  One thread owns one output row `m`; chunk() receives 32
// consecutive fp32 accumulator columns [n, n+32) of that row.
#pragma once
#include "ptx.cuh"

namespace epi {

__device__ __forceinline__ void store_bf16x32(__nv_bfloat16* dst, const float* v, int valid) {
  // dst points at 32 consecutive bf16 (64 B); vectorised when fully valid and 16-B aligned
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 p0 = __floats2bfloat162_rn(v[8 * j + 0], v[8 * j + 1]);
      __nv_bfloat162 p1 = __floats2bfloat162_rn(v[8 * j + 2], v[8 * j + 3]);
      __nv_bfloat162 p2 = __floats2bfloat162_rn(v[8 * j + 4], v[8 * j + 5]);
      __nv_bfloat162 p3 = __floats2bfloat162_rn(v[8 * j + 6], v[8 * j + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2);
      u.w = *reinterpret_cast<uint32_t*>(&p3);
      d4[j] = u;
    }
  } else {
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = __float2bfloat16(v[j]);
  }
}

__device__ __forceinline__ void load_bf16x32(const __nv_bfloat16* src, float* v, int valid) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u = s4[j];
      const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __bfloat1622float2(p[t]);
        v[8 * j + 2 * t] = f.x;
        v[8 * j + 2 * t + 1] = f.y;
      }
    }
  } else {
    for (int j = 0; j < 32; ++j) v[j] = (j < valid) ? __bfloat162float(src[j]) : 0.f;
  }
}

__device__ __forceinline__ void store_f32x32(float* dst, const float* v, int valid) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    float4* d4 = reinterpret_cast<float4*>(dst);
#pragma unroll
    for (int j = 0; j < 8; ++j) d4[j] = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  } else {
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = v[j];
  }
}

// ------------------------------------------------------------------------------------------------ plain store
// C[b][m][n] (+)= alpha * acc    (fp32 or bf16 output)
template <bool OUT_BF16>
struct Store {
  struct Params {
    void* C;
    int64_t ldc, strideC;
    int M, N;
    float alpha;
    int accumulate;
  };
  Params p;
  __device__ explicit Store(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ void chunk(int b, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
    const int valid = min(32, p.N - n);
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
    if constexpr (OUT_BF16) {
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(p.C) + b * p.strideC + (int64_t)m * p.ldc + n;
      store_bf16x32(dst, v, valid);
    } else {
      float* dst = reinterpret_cast<float*>(p.C) + b * p.strideC + (int64_t)m * p.ldc + n;
      if (p.accumulate) {
        for (int j = 0; j < 32; ++j)
          if (j < valid) dst[j] += v[j];
      } else {
        store_f32x32(dst, v, valid);
      }
    }
  }
  __device__ void tile_end(int, int, int, int) {}
};

}  // namespace epi
