// Fused epilogue functors for eng::gemm_kernel.  One thread owns one output row `m`; chunk() receives 32
// consecutive fp32 accumulator columns [n, n+32) of that row.
#pragma once
#include "../../include/clipk.h"
#include "gemm_engine.cuh"
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
    const float* alpha_dev;   // nullable: alpha read from device memory (no host copy of a device scalar)
  };
  Params p;
  __device__ explicit Store(const Params& pp) : p(pp) {
    if (p.alpha_dev != nullptr) p.alpha = __ldg(p.alpha_dev);
  }
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
        if (valid >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
          float4* d4 = reinterpret_cast<float4*>(dst);
          float4 o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = d4[j];            // all loads in flight before the first store
#pragma unroll
          for (int j = 0; j < 8; ++j)
            d4[j] = make_float4(o[j].x + v[4 * j], o[j].y + v[4 * j + 1], o[j].z + v[4 * j + 2], o[j].w + v[4 * j + 3]);
        } else {
          for (int j = 0; j < 32; ++j)
            if (j < valid) dst[j] += v[j];
        }
      } else {
        store_f32x32(dst, v, valid);
      }
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

}  // namespace epi

namespace epi {

__device__ __forceinline__ void store_f16x32(__half* dst, const float* v, int valid) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __half2 p0 = __floats2half2_rn(v[8 * j + 0], v[8 * j + 1]);
      __half2 p1 = __floats2half2_rn(v[8 * j + 2], v[8 * j + 3]);
      __half2 p2 = __floats2half2_rn(v[8 * j + 4], v[8 * j + 5]);
      __half2 p3 = __floats2half2_rn(v[8 * j + 6], v[8 * j + 7]);
      uint4 u;
      u.x = *reinterpret_cast<uint32_t*>(&p0);
      u.y = *reinterpret_cast<uint32_t*>(&p1);
      u.z = *reinterpret_cast<uint32_t*>(&p2);
      u.w = *reinterpret_cast<uint32_t*>(&p3);
      d4[j] = u;
    }
  } else {
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = __float2half(v[j]);
  }
}
__device__ __forceinline__ void load_f16x32(const __half* src, float* v, int valid) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(src) & 15) == 0) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u = s4[j];
      const __half2* p = reinterpret_cast<const __half2*>(&u);
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        float2 f = __half22float2(p[t]);
        v[8 * j + 2 * t] = f.x;
        v[8 * j + 2 * t + 1] = f.y;
      }
    }
  } else {
    for (int j = 0; j < 32; ++j) v[j] = (j < valid) ? __half2float(src[j]) : 0.f;
  }
}

// sigmoid(10 s) = 0.5 tanh(5 s) + 0.5 : one MUFU (tanh.approx.f32, rel. error ~2^-11, far below the bf16 rounding
// the activation receives before it is used as an MMA operand)
__device__ __forceinline__ float fast_sigmoid10(float s) {
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(5.f * s));
  return fmaf(0.5f, t, 0.5f);
}
// round-to-nearest-even to bf16 precision with integer ALU ops (finite inputs), result kept as fp32
__device__ __forceinline__ float bf16_round(float x) {
  uint32_t u = __float_as_uint(x);
  u += 0x7FFFu + ((u >> 16) & 1u);
  return __uint_as_float(u & 0xFFFF0000u);
}
// store 32 fp32 values that are ALREADY bf16-representable: pack the high halves with PRMT (no conversion)
__device__ __forceinline__ void store_bf16x32_exact(__nv_bfloat16* dst, const float* v, int valid) {
  if (valid >= 32 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
    uint4* d4 = reinterpret_cast<uint4*>(dst);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      uint4 u;
      u.x = __byte_perm(__float_as_uint(v[8 * j + 0]), __float_as_uint(v[8 * j + 1]), 0x7632);
      u.y = __byte_perm(__float_as_uint(v[8 * j + 2]), __float_as_uint(v[8 * j + 3]), 0x7632);
      u.z = __byte_perm(__float_as_uint(v[8 * j + 4]), __float_as_uint(v[8 * j + 5]), 0x7632);
      u.w = __byte_perm(__float_as_uint(v[8 * j + 6]), __float_as_uint(v[8 * j + 7]), 0x7632);
      d4[j] = u;
    }
  } else {
    for (int j = 0; j < 32; ++j)
      if (j < valid) dst[j] = __float2bfloat16(v[j]);
  }
}

// ------------------------------------------------------------------------------------ plain bf16 store via TMA
// C[b][m][n] = bf16(alpha * acc), written by the engine's TMA-store path (needs 16-byte aligned rows)
struct StoreTma {
  static constexpr bool kTmaOut = true;
  struct Params {
    eng::OutDesc out;
    float alpha;
    const float* alpha_dev;   // nullable
  };
  Params p;
  __device__ explicit StoreTma(const Params& pp) : p(pp) {
    if (p.alpha_dev != nullptr) p.alpha = __ldg(p.alpha_dev);
  }
  __device__ void tile_begin(int, int, int) {}
  __device__ void chunk(int, int, int, float* v) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= p.alpha;
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM1
// acc[m=text k][n=patch p] = <T_k, V_ip> (raw).  s = acc * rnT[k] * rnV[i,p];  a = sigmoid(10 s)  (pacl.py:133)
// output (TMA store): A bf16 [batch][M][Ppad]; columns >= P are never written (the output map's extent is P) and
// never read (the operand maps of the consumers have extent P: TMA zero-fills beyond it);
// side output: num[i,k] += sum_p a * <t^_k, V_ip> = <u_ik, t^_k>.
// Side data: lane l holds rnV[i, n + l] of the chunk (one coalesced load issued a chunk ahead, broadcast by shuffle).
__device__ __forceinline__ void bf16_round_pair(float& a0, float& a1) {
  const uint32_t pk = ptx::pack_bf16x2(a0, a1);       // one cvt.rn.bf16x2 for two values
  a0 = __uint_as_float(pk << 16);
  a1 = __uint_as_float(pk & 0xFFFF0000u);
}
struct PaclAct {
  static constexpr bool kTmaOut = true;
  using Side = float;
  struct Params {
    eng::OutDesc out;   // A
    const float* rnV;   // [batch][P]
    const float* rnT;   // [M], or nullptr when the A operand is already T^ = T rnT
    float* num;         // [batch][M] or nullptr
    int M, P, Ppad, act;
  };
  Params p;
  float rt, rt5, acc;
  __device__ explicit PaclAct(const Params& pp) : p(pp), rt(0.f), rt5(0.f), acc(0.f) {}
  __device__ void tile_begin(int, int m, int) {
    acc = 0.f;
    rt = (m < p.M) ? (p.rnT != nullptr ? __ldg(p.rnT + m) : 1.f) : 0.f;
    rt5 = 5.f * rt;
  }
  __device__ Side pre(int b, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (n + lane < p.P) ? __ldg(p.rnV + (int64_t)b * p.P + n + lane) : 0.f;
  }
  __device__ void chunk(int, int, int n, float* v, const Side& rn_l) {
    if (p.act == CLIPK_ACT_ONES) {                 // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        acc += v[j];
        v[j] = 1.f;
      }
    } else if (p.act == CLIPK_ACT_SOFTMAX10) {     // a = exp(10 (s - 1)) = 2^(10 log2(e) (s - 1))
      const float k2 = 2.f * 1.4426950408889634f * rt5;          // 10 log2(e) rnT
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        float r[8], a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __shfl_sync(0xffffffffu, rn_l, j0 + i);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = ptx::ex2_approx(fmaf(v[j0 + i] * k2, r[i], -14.426950408889634f));
#pragma unroll
        for (int i = 0; i < 8; i += 2) bf16_round_pair(a[i], a[i + 1]);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc = fmaf(a[i], v[j0 + i], acc);
          v[j0 + i] = a[i];
        }
      }
    } else {
      // blocks of 8 columns, each stage over the whole block: 8 independent shuffle -> MUFU -> FMA chains in flight
#pragma unroll
      for (int j0 = 0; j0 < 32; j0 += 8) {
        float r[8], a[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = __shfl_sync(0xffffffffu, rn_l, j0 + i);   // rnV of the column (rt5 is per row)
#pragma unroll
        for (int i = 0; i < 8; ++i) r[i] = v[j0 + i] * rt5 * r[i];
#pragma unroll
        for (int i = 0; i < 8; ++i) asm("tanh.approx.f32 %0, %1;" : "=f"(a[i]) : "f"(r[i]));   // sigmoid(10 s) = 0.5 tanh(5 s) + 0.5
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] = fmaf(0.5f, a[i], 0.5f);
#pragma unroll
        for (int i = 0; i < 8; i += 2) bf16_round_pair(a[i], a[i + 1]);               // the value the pooling GEMM will see
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          acc = fmaf(a[i], v[j0 + i], acc);        // sum_p a <T_k, V_p>   (scaled by rnT at tile end)
          v[j0 + i] = a[i];
        }
      }
    }
  }
  __device__ void tile_end(int b, int m, int, int, int) {
    if (p.num != nullptr && m < p.M) atomicAdd(p.num + (int64_t)b * p.M + m, acc * rt);
  }
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM2 (fwd)
// acc = u_ik[d];  usq[i,k] += sum_d u^2   (the pooled vector itself never leaves the SM)
struct Usq {
  struct Params {
    float* usq;   // [batch][M]
    int M, N;
  };
  Params p;
  float acc;
  __device__ explicit Usq(const Params& pp) : p(pp), acc(0.f) {}
  __device__ void tile_begin(int, int, int) { acc = 0.f; }
  __device__ void chunk(int, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (n + j < p.N) acc = fmaf(v[j], v[j], acc);
  }
  __device__ void tile_end(int b, int m, int, int, int) {
    if (m < p.M) atomicAdd(p.usq + (int64_t)b * p.M + m, acc);
  }
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM2 (fwd, saving u)
// acc = u_ik[d];  usq[i,k] += sum_d u^2 (from the fp32 accumulator) and the pooled vector itself is stored as bf16
// (TMA store) for the backward pass, which then needs neither the recompute of the activations' pooling GEMM nor a
// per-group G scratch: the gradient factors (alpha, beta) are applied where u is consumed.
struct UsqStore {
  static constexpr bool kTmaOut = true;
  struct Params {
    eng::OutDesc out;   // U bf16 [batch][M][N]
    float* usq;         // [batch][M]
    int M, N;
  };
  Params p;
  float acc;
  __device__ explicit UsqStore(const Params& pp) : p(pp), acc(0.f) {}
  __device__ void tile_begin(int, int, int) { acc = 0.f; }
  __device__ void chunk(int, int, int n, float* v) {
    if (n + 32 <= p.N) {
#pragma unroll
      for (int j = 0; j < 32; ++j) acc = fmaf(v[j], v[j], acc);
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (n + j < p.N) acc = fmaf(v[j], v[j], acc);
    }
  }
  __device__ void tile_end(int b, int m, int, int, int) {
    if (m < p.M) atomicAdd(p.usq + (int64_t)b * p.M + m, acc);
  }
};

// ------------------------------------------------------------------------------------ PACL all-pairs, dual GEMM (bwd)
// Two accumulators over the same V tile:  x = <T^_k, V_ip>  (A0 = T^ = bf16(T rnT))  and  d = <u_ik, V_ip>  (A1 = the
// pooled vectors saved by the forward).  One pass produces everything the two gradient GEMMs need:
//   s  = x rnV,  a = bf16(act(s))
//   da = <G_ik, V_ip> = alpha x - beta d          (G = alpha t^ - beta u: gradient w.r.t. the un-normalised pooled vector)
//   ds = da * act'(s)
//   E  = ds rnV + alpha a        (out : operand of dt^ += E V  and of dV += E^T T^)
//   A2 = -beta a                 (out2: operand of dV += A2^T U, the pooling term sum_k a G without its t^ part)
//   dsdot[i,p] += sum_k ds * s   (= <v^_ip, dv^_ip>, the normalise-Jacobian projection)
struct DsDual {
  static constexpr bool kDual = true;
  static constexpr bool kTmaOut = true;
  static constexpr bool kTmaOut2 = true;
  using Side = float;          // lane l holds rnV[i, n + l]
  struct Params {
    eng::OutDesc out;        // E  [batch][M][Ppad]
    eng::OutDesc out2;       // A2 [batch][M][Ppad]
    const float* rnV;        // [batch][P]
    const float* alpha;      // [batch][M]
    const float* beta;       // [batch][M]
    float* dsdot;            // [batch][P]
    int M, P, act;
  };
  Params p;
  float al, nb;
  __device__ explicit DsDual(const Params& pp) : p(pp), al(0.f), nb(0.f) {}
  __device__ void tile_begin(int b, int m, int) {
    al = (m < p.M) ? __ldg(p.alpha + (int64_t)b * p.M + m) : 0.f;
    nb = (m < p.M) ? -__ldg(p.beta + (int64_t)b * p.M + m) : 0.f;
  }
  __device__ Side pre(int b, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (n + lane < p.P) ? __ldg(p.rnV + (int64_t)b * p.P + n + lane) : 0.f;
  }
  // in: x[] = <T^,V>, d[] = <u,V>;  out: x[] = E, d[] = A2
  __device__ void chunk2(int b, int, int n, float* x, float* d, const Side& rn_l) {
    const int lane = (int)ptx::lane_id();
    const int col = n + lane;
#ifdef CLIPK_EPI_STUB
    if (p.act == 77) return;      // timing experiment: no epilogue math (garbage results)
#endif
    if (p.act == CLIPK_ACT_ONES) {                       // warp-uniform: a = 1, ds = 0
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = al;
        d[j] = nb;
      }
      return;
    }
    float dss[32];
    if (p.act == CLIPK_ACT_SOFTMAX10) {
      // a = exp(10 (s - 1)),  ds = da * 10 a
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float r0 = __shfl_sync(0xffffffffu, rn_l, j);        // 0 for p >= P
        const float r1 = __shfl_sync(0xffffffffu, rn_l, j + 1);
        const float s0 = x[j] * r0, s1 = x[j + 1] * r1;
        float a0 = ptx::ex2_approx(14.426950408889634f * (s0 - 1.f));
        float a1 = ptx::ex2_approx(14.426950408889634f * (s1 - 1.f));
        bf16_round_pair(a0, a1);
        const float ds0 = fmaf(al, x[j], nb * d[j]) * 10.f * a0;
        const float ds1 = fmaf(al, x[j + 1], nb * d[j + 1]) * 10.f * a1;
        x[j] = fmaf(ds0, r0, al * a0);
        x[j + 1] = fmaf(ds1, r1, al * a1);
        d[j] = nb * a0;
        d[j + 1] = nb * a1;
        dss[j] = ds0 * s0;
        dss[j + 1] = ds1 * s1;
      }
    } else {
      // factors of 5 folded as in DsIn: r5 = 5 rnV, s5 = 5 s, ds5 = ds / 5 = da * 2 a (1 - a); E = alpha a + ds5 r5
      const float r5_l = 5.f * rn_l;
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float r0 = __shfl_sync(0xffffffffu, r5_l, j);
        const float r1 = __shfl_sync(0xffffffffu, r5_l, j + 1);
        const float s0 = x[j] * r0, s1 = x[j + 1] * r1;
        float t0, t1;
        asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(s0));
        asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(s1));
        float a0 = fmaf(0.5f, t0, 0.5f);
        float a1 = fmaf(0.5f, t1, 0.5f);
        bf16_round_pair(a0, a1);
        const float u0 = a0 + a0, u1 = a1 + a1;
        const float ds0 = fmaf(al, x[j], nb * d[j]) * fmaf(-u0, a0, u0);        // da * 2 a (1 - a)
        const float ds1 = fmaf(al, x[j + 1], nb * d[j + 1]) * fmaf(-u1, a1, u1);
        x[j] = fmaf(ds0, r0, al * a0);                                         // E
        x[j + 1] = fmaf(ds1, r1, al * a1);
        d[j] = nb * a0;                                                        // A2
        d[j + 1] = nb * a1;
        dss[j] = ds0 * s0;
        dss[j + 1] = ds1 * s1;
      }
    }
    const float cs = ptx::warp_colsum32(dss);   // lane j: sum over this warp's 32 rows of column n + j
    if (col < p.P && cs != 0.f) atomicAdd(p.dsdot + (int64_t)b * p.P + col, cs);
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM2 (bwd)
// acc = u_ik[d].  The gradient w.r.t. the un-normalised pooled vector is G_ik = alpha_ik t^_k - beta_ik u_ik; only its
// u-part  Gn_ik = -beta_ik u_ik  is materialised (bf16, TMA store): the t^ part is folded analytically into the
// consumers (DsIn adds alpha <t^,V>; the dV GEMM uses E^T T^ which already contains alpha a t^).
struct GNeg {
  static constexpr bool kTmaOut = true;
  struct Params {
    eng::OutDesc out;          // Gn [batch][M][N]
    const float* beta;         // [batch][M]
    int M;
  };
  Params p;
  float nb;
  __device__ explicit GNeg(const Params& pp) : p(pp), nb(0.f) {}
  __device__ void tile_begin(int b, int m, int) { nb = (m < p.M) ? -__ldg(p.beta + (int64_t)b * p.M + m) : 0.f; }
  __device__ void chunk(int, int, int, float* v) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] *= nb;
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM1 (bwd)
// The backward's recompute of GEMM1 (CTA-pair engine): besides the activations A it stores the text-normalised raw
// scores  X = bf16(<t^_k, V_ip>) = bf16(acc * rnT[k]),  which the dS kernel (DsIn) reads back chunk by chunk through
// TMA instead of recomputing the T V^T GEMM a second time.
struct PaclActS {
  static constexpr bool kTmaOut = true;
  static constexpr bool kTmaOut2 = true;
  using Side = float;          // lane l holds rnV[i, n + l]
  struct Params {
    eng::OutDesc out;    // A [batch][M][Ppad]
    eng::OutDesc out2;   // X [batch][M][Ppad]
    const float* rnV;    // [batch][P]
    const float* rnT;    // [M]
    int M, P, Ppad, act;
  };
  Params p;
  float rt, rt5;
  __device__ explicit PaclActS(const Params& pp) : p(pp), rt(0.f), rt5(0.f) {}
  __device__ void tile_begin(int, int m, int) {
    rt = (m < p.M) ? __ldg(p.rnT + m) : 0.f;
    rt5 = 5.f * rt;
  }
  __device__ Side pre(int b, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (n + lane < p.P) ? __ldg(p.rnV + (int64_t)b * p.P + n + lane) : 0.f;
  }
  __device__ void chunk(int, int, int, float* v, const Side& rn_l, const uint32_t*, float* x) {
    if (p.act == CLIPK_ACT_ONES) {                       // warp-uniform
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        x[j] = v[j] * rt;
        v[j] = 1.f;
      }
      return;
    }
    if (p.act == CLIPK_ACT_SOFTMAX10) {
      const float k2 = 2.f * 1.4426950408889634f * rt5;          // 10 log2(e) rnT
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const float r0 = __shfl_sync(0xffffffffu, rn_l, j);
        x[j] = v[j] * rt;
        v[j] = ptx::ex2_approx(fmaf(v[j] * k2, r0, -14.426950408889634f));
      }
      return;
    }
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float r0 = __shfl_sync(0xffffffffu, rn_l, j);
      const float r1 = __shfl_sync(0xffffffffu, rn_l, j + 1);
      x[j] = v[j] * rt;                                    // X = <t^_k, V_ip>
      x[j + 1] = v[j + 1] * rt;
      float t0, t1;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(v[j] * rt5 * r0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(v[j + 1] * rt5 * r1));
      v[j] = fmaf(0.5f, t0, 0.5f);
      v[j + 1] = fmaf(0.5f, t1, 0.5f);
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, GEMM3 (bwd)
// acc[m=text k][n=patch p] = d = <Gn_ik, V_ip>;  input chunk X = <t^_k, V_ip> (bf16, written by PaclActS).
//   s  = X rnV,  a = bf16(sigmoid(10 s))
//   da = alpha_ik X + d   ( = <G_ik, V_ip>,  G = alpha t^ - beta u ),   ds = da * 10 a (1 - a)
//   E[i,k,p] = ds * rnV[i,p] + alpha_ik a     (TMA store; operand of dt^ += E V  and of dV += E^T T^)
//   dsdot[i,p] += sum_k ds * s                (= <v^_ip, dv^_ip>, the normalise-Jacobian projection)
struct DsIn {
  static constexpr bool kTmaOut = true;
  static constexpr bool kChunkIn = true;
  using Side = float;          // lane l holds rnV[i, n + l]
  struct Params {
    eng::OutDesc out;        // E [batch][M][Ppad]
    eng::OutDesc in;         // X [batch][M][Ppad]
    const float* rnV;        // [batch][P]
    const float* alpha;      // [batch][M]
    float* dsdot;            // [batch][P]
    int M, P, Ppad, act;
  };
  Params p;
  float al;
  __device__ explicit DsIn(const Params& pp) : p(pp), al(0.f) {}
  __device__ void tile_begin(int b, int m, int) { al = (m < p.M) ? __ldg(p.alpha + (int64_t)b * p.M + m) : 0.f; }
  __device__ Side pre(int b, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (n + lane < p.P) ? __ldg(p.rnV + (int64_t)b * p.P + n + lane) : 0.f;
  }
  __device__ void chunk(int b, int m, int n, float* d, const Side& rn_l, const uint32_t* in, float*) {
    const int lane = (int)ptx::lane_id();
    const int col = n + lane;
    if (p.act == CLIPK_ACT_ONES) {                       // warp-uniform: a = 1, ds = 0  ->  E = alpha, dsdot untouched
#pragma unroll
      for (int j = 0; j < 32; ++j) d[j] = al;
      return;
    }
    if (p.act == CLIPK_ACT_SOFTMAX10) {
      // a = exp(10 (s - 1)),  ds = da * 10 a;   E = alpha a + ds rnV,   ds * s -> dsdot
      const float g10 = m < p.M ? 10.f : 0.f;
      float dss[32];
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const float r0 = __shfl_sync(0xffffffffu, rn_l, j);        // 0 for p >= P
        const float r1 = __shfl_sync(0xffffffffu, rn_l, j + 1);
        const float x0 = __uint_as_float(in[j >> 1] << 16);
        const float x1 = __uint_as_float(in[j >> 1] & 0xFFFF0000u);
        const float s0 = x0 * r0, s1 = x1 * r1;
        float a0 = ptx::ex2_approx(14.426950408889634f * (s0 - 1.f));
        float a1 = ptx::ex2_approx(14.426950408889634f * (s1 - 1.f));
        bf16_round_pair(a0, a1);
        const float ds0 = fmaf(al, x0, d[j]) * g10 * a0;
        const float ds1 = fmaf(al, x1, d[j + 1]) * g10 * a1;
        d[j] = fmaf(ds0, r0, al * a0);
        d[j + 1] = fmaf(ds1, r1, al * a1);
        dss[j] = ds0 * s0;
        dss[j + 1] = ds1 * s1;
      }
      const float cs = ptx::warp_colsum32(dss);
      if (col < p.P && cs != 0.f) atomicAdd(p.dsdot + (int64_t)b * p.P + col, cs);
      return;
    }
    // All factors of 5 are folded:  r5 = 5 rnV,  s5 = 5 s = X r5,  ds5 = ds / 5 = da * 2 a (1 - a)
    //   E = alpha a + ds rnV = alpha a + ds5 r5          ds * s = ds5 * s5
    const float r5_l = 5.f * rn_l;                       // 0 for p >= P: masks the pad columns
    const float g2 = m < p.M ? 2.f : 0.f;
    float dss[32];
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      const float r0 = __shfl_sync(0xffffffffu, r5_l, j);
      const float r1 = __shfl_sync(0xffffffffu, r5_l, j + 1);
      const float x0 = __uint_as_float(in[j >> 1] << 16);
      const float x1 = __uint_as_float(in[j >> 1] & 0xFFFF0000u);
      const float s0 = x0 * r0, s1 = x1 * r1;
      float t0, t1;
      asm("tanh.approx.f32 %0, %1;" : "=f"(t0) : "f"(s0));
      asm("tanh.approx.f32 %0, %1;" : "=f"(t1) : "f"(s1));
      float a0 = fmaf(0.5f, t0, 0.5f);
      float a1 = fmaf(0.5f, t1, 0.5f);
      bf16_round_pair(a0, a1);
      const float u0 = a0 * g2, u1 = a1 * g2;
      const float ds0 = fmaf(al, x0, d[j]) * fmaf(-u0, a0, u0);        // da * 2 a (1 - a)
      const float ds1 = fmaf(al, x1, d[j + 1]) * fmaf(-u1, a1, u1);
      d[j] = fmaf(ds0, r0, al * a0);                                   // E
      d[j + 1] = fmaf(ds1, r1, al * a1);
      dss[j] = ds0 * s0;                                               // ds * s  (0 in pad columns / invalid rows)
      dss[j + 1] = ds1 * s1;
    }
    const float cs = ptx::warp_colsum32(dss);   // lane j: sum over this warp's 32 rows of column n + j
    if (col < p.P && cs != 0.f) atomicAdd(p.dsdot + (int64_t)b * p.P + col, cs);
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, dV (bwd)
// acc[m=patch p][n=d] = sum_k a_ikp Gn_ik[d] + sum_k E_ikp T^_k[d];  dV_ip = acc - rnV_ip^2 dsdot_ip V_ip  (TMA store)
struct DvOut {
  static constexpr bool kTmaOut = true;
  struct Side {
    uint4 q[4];     // 32 bf16 of V[b][m][n .. n+32)
  };
  struct Params {
    eng::OutDesc out;         // dV [batch][M][N]
    const __nv_bfloat16* V;   // [batch][M][N]   (N % 8 == 0: 16-byte aligned rows)
    const float* rnV;         // [batch][M]
    const float* dsdot;       // [batch][M]
    int M, N;
  };
  Params p;
  float coef;
  __device__ explicit DvOut(const Params& pp) : p(pp), coef(0.f) {}
  __device__ void tile_begin(int b, int m, int) {
    coef = 0.f;
    if (m < p.M) {
      const float r = __ldg(p.rnV + (int64_t)b * p.M + m);
      coef = r * r * __ldcg(p.dsdot + (int64_t)b * p.M + m);   // L2: may have been written earlier in this kernel
    }
  }
  __device__ Side pre(int b, int m, int n) const {
    Side s;
    const uint4* src = reinterpret_cast<const uint4*>(p.V + ((int64_t)b * p.M + m) * p.N + n);
#pragma unroll
    for (int j = 0; j < 4; ++j)
      s.q[j] = (m < p.M && n + 8 * j + 8 <= p.N) ? __ldg(src + j) : make_uint4(0u, 0u, 0u, 0u);
    return s;
  }
  __device__ void chunk(int, int, int, float* v, const Side& s) {
    const float nc = -coef;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t w[4] = {s.q[j].x, s.q[j].y, s.q[j].z, s.q[j].w};
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        v[8 * j + 2 * t] = fmaf(nc, __uint_as_float(w[t] << 16), v[8 * j + 2 * t]);
        v[8 * j + 2 * t + 1] = fmaf(nc, __uint_as_float(w[t] & 0xFFFF0000u), v[8 * j + 2 * t + 1]);
      }
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ------------------------------------------------------------------------------------ PACL all-pairs, dV transposed
// Same result as DvOut computed as dV_i^T: acc[m = d][n = patch p] = sum_k Gn_ik[d] a_ikp + sum_k T^_k[d] E_ikp, so
// that M = D (a multiple of 256 for the CLIP widths) instead of M = P = 576 (which wastes a third of a 768-row cover).
//   dV[i][p][d] = acc[d][p] - rnV_ip^2 dsdot_ip V[i][p][d]
// The V chunk arrives by TMA as [32 p][32 d] (the output's layout); this lane owns column d of it and of the staged
// output chunk.  Side: lane l holds coef = rnV^2 dsdot of patch n + l.
struct DvOutT {
  static constexpr bool kTmaOut = true;
  static constexpr bool kChunkIn = true;
  static constexpr bool kTransposed = true;
  using Side = float;
  struct Params {
    eng::OutDesc out;         // dV [batch][P][D]
    eng::OutDesc in;          // V  [batch][P][D]
    const float* rnV;         // [batch][P]
    const float* dsdot;       // [batch][P]
    int P;
  };
  Params p;
  __device__ explicit DvOutT(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ Side pre(int b, int, int n) const {
    const int col = n + (int)ptx::lane_id();
    if (col >= p.P) return 0.f;
    const float r = __ldg(p.rnV + (int64_t)b * p.P + col);
    return r * r * __ldcg(p.dsdot + (int64_t)b * p.P + col);     // dsdot was written earlier in this stream / kernel
  }
  __device__ void chunk_t(int, int, int, float* v, const Side& coef_l, uint32_t in_addr, uint32_t out_addr, int lane) {
    // element (row j = patch, col = lane = d): 64-byte rows, 16-byte unit u of row j sits at u ^ ((j >> 1) & 3)
    const uint32_t unit = static_cast<uint32_t>(lane) >> 3, within = (static_cast<uint32_t>(lane) & 7u) * 2u;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const uint32_t off = j * 64 + ((unit ^ ((static_cast<uint32_t>(j) >> 1) & 3u)) << 4) + within;
      const float cj = __shfl_sync(0xffffffffu, coef_l, j);
      const float x = __uint_as_float(ptx::ld_shared_u16(in_addr + off) << 16);
      const float o = fmaf(-cj, x, v[j]);
      ptx::st_shared_u16(out_addr + off, ptx::pack_bf16x2(o, 0.f));
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

}  // namespace epi
