// PACL projection heads: the producer of the patch tensor V the scorer reads.
//
//   visual_projection   PACL/model/pacl.py:70-74   LayerNorm(Din) -> Dropout(0.1) -> Patch_Projection(Din, Dout)
//   Patch_Projection    PACL/model/pacl.py:35-48   y = W1 x + b1 + W3 gelu(W2 x + b2) + b3      (nn.GELU(): erf form)
//   text_projection     PACL/model/pacl.py:75-79   LayerNorm(D) -> Dropout(0.1) -> Linear(D, D)
//
// Layout: tokens are rows.  xn = LN(x) bf16 [R, Din]; weights bf16 [out, in] (nn.Linear layout), biases fp32.
//   fwd   z = xn W2^T + b2, H = gelu(z), G' = gelu'(z)   one CTA-pair GEMM, two TMA outputs (G' is kept for the backward)
//         Y = xn W1^T + H W3^T + (b1 + b3)   one CTA-pair GEMM over two operand pairs (K = Din, then K = Dout)
//   bwd   dZ = (dY W3) * G'                  GEMM with the G' chunks arriving by TMA in the epilogue
//         dxn = dY W1 + dZ W2                two operand pairs
//         dW1 = dY^T xn, dW2 = dZ^T xn, dW3 = dY^T H    split-K over token blocks into fp32 slabs + a reduction
//         db1 = db3 = colsum(dY), db2 = colsum(dZ)
//   LayerNorm forward / backward (d gamma, d beta through deterministic two-stage column sums, optional dx).
#include "common.cuh"
#include "epilogues.cuh"
#include "simt_util.cuh"

namespace epi {

// erf by Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7): one rcp and one ex2 on the MUFU pipe, 6 FMAs.
// Returns erf(x) and e = exp(-x^2) (the backward needs it for the Gaussian density).
__device__ __forceinline__ float erf_as(float x, float& e) {
  const float ax = fabsf(x);
  float t, q = ax * ax * -1.4426950408889634f;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(t) : "f"(fmaf(0.3275911f, ax, 1.f)));      // one MUFU each: the IEEE
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(q));                               // forms add fix-up code
  float p = fmaf(1.061405429f, t, -1.453152027f);
  p = fmaf(p, t, 1.421413741f);
  p = fmaf(p, t, -0.284496736f);
  p = fmaf(p, t, 0.254829592f);
  const float r = fmaf(-p * t, e, 1.f);
  return copysignf(r, x);
}

// acc = xn W2^T; z = acc + b2; first output h = gelu(z), second output g = gelu'(z) (both bf16):
//   gelu(z) = 0.5 z (1 + erf(z / sqrt 2)),   gelu'(z) = 0.5 (1 + erf(z / sqrt 2)) + z exp(-z^2 / 2) / sqrt(2 pi)
// erf and the Gaussian share one evaluation, so the derivative costs three more instructions here and turns the
// backward's epilogue into a single multiply (the pre-activation itself is not needed again).
struct BiasGelu2 {
  static constexpr bool kTmaOut = true;
  static constexpr bool kTmaOut2 = true;
  using Side = float;          // lane l holds bias[n + l]
  struct Params {
    eng::OutDesc out;    // H  [1][R][N]
    eng::OutDesc out2;   // G' [1][R][N]
    const float* bias;   // [N]
    int N;
  };
  Params p;
  __device__ explicit BiasGelu2(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ Side pre(int, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (n + lane < p.N) ? __ldg(p.bias + n + lane) : 0.f;
  }
  __device__ void chunk(int, int, int, float* v, const Side& b_l, const uint32_t*, float* g) {
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float z = v[j] + __shfl_sync(0xffffffffu, b_l, j);
      float e;
      const float f = erf_as(z * 0.70710678118654752f, e);
      const float phi = fmaf(0.5f, f, 0.5f);                  // Phi(z)
      v[j] = z * phi;
      g[j] = fmaf(z * 0.3989422804014327f, e, phi);
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// out = bf16(acc + bias[n]);  optionally rowsq[m] += sum_n out[m][n]^2 over the bf16-ROUNDED outputs: the squared row norms
// the PACL scorer needs of its patch tensor (F.normalize, pacl.py:122), so that it does not have to read V once more
struct BiasTma {
  static constexpr bool kTmaOut = true;
  using Side = float;
  struct Params {
    eng::OutDesc out;
    const float* bias;   // [N] or nullptr
    int N;
    float* rowsq;        // [M] or nullptr (zeroed by the caller)
    int M;
  };
  Params p;
  float sq;
  __device__ explicit BiasTma(const Params& pp) : p(pp), sq(0.f) {}
  __device__ void tile_begin(int, int, int) { sq = 0.f; }
  __device__ Side pre(int, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (p.bias != nullptr && n + lane < p.N) ? __ldg(p.bias + n + lane) : 0.f;
  }
  __device__ void chunk(int, int, int n, float* v, const Side& b_l) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, b_l, j);
    if (p.rowsq != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; j += 2) {
        const uint32_t pk = ptx::pack_bf16x2(v[j], v[j + 1]);          // the values the store will write
        const float a0 = __uint_as_float(pk << 16), a1 = __uint_as_float(pk & 0xFFFF0000u);
        if (n + j < p.N) sq = fmaf(a0, a0, sq);
        if (n + j + 1 < p.N) sq = fmaf(a1, a1, sq);
      }
    }
  }
  __device__ void tile_end(int, int m, int, int, int) {
    if (p.rowsq != nullptr && m < p.M) atomicAdd(p.rowsq + m, sq);
  }
};

// acc = dY W3 (= dH);  dZ = dH * gelu'(z): the derivative chunks (stored by the forward) arrive through TMA
// (kChunkIn), dZ leaves through TMA.
struct GeluBwdIn {
  static constexpr bool kTmaOut = true;
  static constexpr bool kChunkIn = true;
  using Side = float;          // unused (the chunk-in path needs a side type)
  struct Params {
    eng::OutDesc out;    // dZ
    eng::OutDesc in;     // G' = gelu'(z)
  };
  Params p;
  __device__ explicit GeluBwdIn(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ Side pre(int, int, int) const { return 0.f; }
  __device__ void chunk(int, int, int, float* d, const Side&, const uint32_t* in, float*) {
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
      d[j] *= __uint_as_float(in[j >> 1] << 16);
      d[j + 1] *= __uint_as_float(in[j >> 1] & 0xFFFF0000u);
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

}  // namespace epi

namespace clipk {

static inline int64_t rup64(int64_t x, int64_t m) { return (x + m - 1) / m * m; }

template <typename T>
__device__ __forceinline__ void load8(const T* p, float* f);
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float* f) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(w[i] << 16);
    f[2 * i + 1] = __uint_as_float(w[i] & 0xFFFF0000u);
  }
}
template <>
__device__ __forceinline__ void load8<float>(const float* p, float* f) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p)), b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float* f) {
  *reinterpret_cast<uint4*>(p) = make_uint4(ptx::pack_bf16x2(f[0], f[1]), ptx::pack_bf16x2(f[2], f[3]),
                                            ptx::pack_bf16x2(f[4], f[5]), ptx::pack_bf16x2(f[6], f[7]));
}
__device__ __forceinline__ void store8(float* p, const float* f) {
  reinterpret_cast<float4*>(p)[0] = make_float4(f[0], f[1], f[2], f[3]);
  reinterpret_cast<float4*>(p)[1] = make_float4(f[4], f[5], f[6], f[7]);
}

// ---------------------------------------------------------------------------------------------- LayerNorm
// One warp per row (D % 8 == 0): mean, biased variance (two passes, fp32), xn = (x - mean) rstd gamma + beta as bf16.
// Dropout fused into the LayerNorm pass (pacl.py:72,77: Sequential(LayerNorm, Dropout(0.1), ...)).  Counter-based: the
// keep decision of element (row, d) depends only on (seed, row * D + d), 16 random bits per element from a 32-bit
// integer mixer, keep iff bits >= thr16 (thr16 = round(p * 65536)); one byte of keep bits per 8-element vector is saved
// for the backward (1 bit per element instead of re-deriving or storing a bool tensor).
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t dropout_keep8(uint64_t seed, uint64_t vec_index, uint32_t thr16) {
  const uint32_t lo = static_cast<uint32_t>(vec_index), hi = static_cast<uint32_t>(vec_index >> 32);
  const uint32_t k = mix32(static_cast<uint32_t>(seed) ^ mix32(hi ^ static_cast<uint32_t>(seed >> 32)));
  uint32_t bits = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t h = mix32((lo * 4u + j) ^ k);
    bits |= ((h & 0xFFFFu) >= thr16 ? 1u : 0u) << (2 * j);
    bits |= ((h >> 16) >= thr16 ? 1u : 0u) << (2 * j + 1);
  }
  return bits;
}
__host__ __device__ inline uint32_t dropout_thr16(float p) {
  const float t = p * 65536.f + 0.5f;
  return t <= 0.f ? 0u : (t >= 65535.f ? 65535u : static_cast<uint32_t>(t));
}
__host__ __device__ inline float dropout_scale(uint32_t thr16) { return 65536.f / (65536.f - static_cast<float>(thr16)); }

template <typename T>
__global__ void ln_fwd_kernel(const T* __restrict__ X, int64_t rows, int D, const float* __restrict__ gamma,
                              const float* __restrict__ beta, float eps, __nv_bfloat16* __restrict__ Y,
                              float* __restrict__ mean, float* __restrict__ rstd, uint8_t* __restrict__ keep,
                              uint32_t thr16, uint64_t seed) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + row * D;
  const int nv = D >> 3;
  float s = 0.f;
  for (int v = lane; v < nv; v += 32) {
    float f[8];
    load8<T>(x + 8 * v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) s += f[i];
  }
  const float mu = ptx::warp_sum(s) / (float)D;
  float q = 0.f;
  for (int v = lane; v < nv; v += 32) {
    float f[8];
    load8<T>(x + 8 * v, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) q = fmaf(f[i] - mu, f[i] - mu, q);
  }
  const float rs = rsqrtf(ptx::warp_sum(q) / (float)D + eps);
  for (int v = lane; v < nv; v += 32) {
    float f[8], g[8], b[8];
    load8<T>(x + 8 * v, f);
    load8<float>(gamma + 8 * v, g);
    load8<float>(beta + 8 * v, b);
#pragma unroll
    for (int i = 0; i < 8; ++i) f[i] = fmaf((f[i] - mu) * rs, g[i], b[i]);
    if (keep != nullptr) {
      const uint64_t vi = (uint64_t)row * nv + v;
      const uint32_t bits = dropout_keep8(seed, vi, thr16);
      const float sc = dropout_scale(thr16);
#pragma unroll
      for (int i = 0; i < 8; ++i) f[i] = ((bits >> i) & 1u) ? f[i] * sc : 0.f;
      keep[vi] = static_cast<uint8_t>(bits);
    }
    store8(Y + row * D + 8 * v, f);
  }
  if (lane == 0) {
    mean[row] = mu;
    rstd[row] = rs;
  }
}

// Column sums for d gamma / d beta (and plain column sums when X == nullptr), stage 1.  Block (bx, by): columns
// [256 bx, 256 bx + 256) of rows [by * rows_per_block, ...); every lane owns 8 consecutive columns (one 16-byte load
// per row), a warp reads 512 contiguous bytes per row, eight rows in flight per block; part[by][{0,1}][D] receives
// the block's sums.      dg[d] = sum_r G[r,d] * (x[r,d] - mean_r) rstd_r      db[d] = sum_r G[r,d]      (D % 8 == 0)
template <typename T>
__global__ void __launch_bounds__(256) colsum_part_kernel(const __nv_bfloat16* __restrict__ G, const T* __restrict__ X,
                                                          const float* __restrict__ mean, const float* __restrict__ rstd,
                                                          int64_t rows, int D, int rows_per_block,
                                                          float* __restrict__ part,
                                                          const uint8_t* __restrict__ keep = nullptr, float keep_scale = 1.f) {
  __shared__ float red[8][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int d = blockIdx.x * 256 + 8 * lane;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  float sg[8], sb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) sg[i] = sb[i] = 0.f;
  if (d < D) {
#pragma unroll 4
    for (int64_t r = r0 + warp; r < r1; r += 8) {
      float g[8];
      load8<__nv_bfloat16>(G + r * D + d, g);
      if (keep != nullptr) {
        const uint32_t bits = keep[(r * D + d) >> 3];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = ((bits >> i) & 1u) ? g[i] * keep_scale : 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sb[i] += g[i];
      if (X != nullptr) {
        float x[8];
        load8<T>(X + r * D + d, x);
        const float mu = __ldg(mean + r), rs = __ldg(rstd + r);
#pragma unroll
        for (int i = 0; i < 8; ++i) sg[i] = fmaf(g[i], (x[i] - mu) * rs, sg[i]);
      }
    }
  }
  float* out = part + (int64_t)blockIdx.y * 2 * D;
  for (int pass = 0; pass < 2; ++pass) {        // pass 0: dg, pass 1: db  (cross-warp sums in a fixed order)
    if (pass == 0 && X == nullptr) {
      if (blockIdx.x * 256 + threadIdx.x < D) out[blockIdx.x * 256 + threadIdx.x] = 0.f;
      continue;
    }
    __syncthreads();
#pragma unroll
    for (int i = 0; i < 8; ++i) red[warp][8 * lane + i] = pass == 0 ? sg[i] : sb[i];
    __syncthreads();
    const int c = blockIdx.x * 256 + threadIdx.x;
    if (c < D) {
      float t = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) t += red[j][threadIdx.x];
      out[pass * D + c] = t;
    }
  }
}
// stage 2: out_g[d] = sum_by part[by][0][d], out_b[d] = sum_by part[by][1][d]  (fixed order: deterministic)
__global__ void colsum_reduce_kernel(const float* __restrict__ part, int nblk, int D, float* __restrict__ out_g,
                                     float* __restrict__ out_b) {
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  float sg = 0.f, sb = 0.f;
  for (int b = 0; b < nblk; ++b) {
    sg += part[(int64_t)b * 2 * D + d];
    sb += part[(int64_t)b * 2 * D + D + d];
  }
  if (out_g != nullptr) out_g[d] = sg;
  if (out_b != nullptr) out_b[d] = sb;
}

// dx = rstd (g - mean_d(g) - xh mean_d(g xh)),  g = dxn * gamma,  xh = (x - mean) rstd     (one warp per row)
template <typename T>
__global__ void ln_bwd_dx_kernel(const __nv_bfloat16* __restrict__ G, const T* __restrict__ X,
                                 const float* __restrict__ gamma, const float* __restrict__ mean,
                                 const float* __restrict__ rstd, int64_t rows, int D, T* __restrict__ dX,
                                 const uint8_t* __restrict__ keep, float keep_scale) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const int nv = D >> 3;
  const float mu = mean[row], rs = rstd[row];
  float s1 = 0.f, s2 = 0.f;
  for (int v = lane; v < nv; v += 32) {
    float g[8], x[8], ga[8];
    load8<__nv_bfloat16>(G + row * D + 8 * v, g);
    load8<T>(X + row * D + 8 * v, x);
    load8<float>(gamma + 8 * v, ga);
    if (keep != nullptr) {
      const uint32_t bits = keep[(uint64_t)row * nv + v];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = ((bits >> i) & 1u) ? g[i] * keep_scale : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float gg = g[i] * ga[i];
      s1 += gg;
      s2 = fmaf(gg, (x[i] - mu) * rs, s2);
    }
  }
  s1 = ptx::warp_sum(s1) / (float)D;
  s2 = ptx::warp_sum(s2) / (float)D;
  for (int v = lane; v < nv; v += 32) {
    float g[8], x[8], ga[8];
    load8<__nv_bfloat16>(G + row * D + 8 * v, g);
    load8<T>(X + row * D + 8 * v, x);
    load8<float>(gamma + 8 * v, ga);
    if (keep != nullptr) {
      const uint32_t bits = keep[(uint64_t)row * nv + v];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = ((bits >> i) & 1u) ? g[i] * keep_scale : 0.f;
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) g[i] = rs * (g[i] * ga[i] - s1 - (x[i] - mu) * rs * s2);
    store8(dX + row * D + 8 * v, g);
  }
}

// out[i] = sum_s slabs[s][i]   (n % 4 == 0; fixed order)
__global__ void slab_reduce_kernel(const float4* __restrict__ slabs, int nslab, int64_t n4, float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = slabs[i];
  for (int s = 1; s < nslab; ++s) {
    const float4 b = slabs[(int64_t)s * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}

// ---------------------------------------------------------------------------------------------- RoPE
// apply_rope (PACL/model/pacl.py:147-181): position p = row % S, pair j = (x[2j], x[2j+1]), angle table [S][D/2]:
//   out[j] = x[2j] cos - x[2j+1] sin,   out[D/2 + j] = x[2j] sin + x[2j+1] cos        (de-interleave, then concat halves)
// INVERSE != 0 applies the transpose (the gradient): dx[2j] = g[j] cos + g[D/2+j] sin, dx[2j+1] = -g[j] sin + g[D/2+j] cos.
template <typename TI, typename TO, bool INVERSE>
__global__ void rope_kernel(const TI* __restrict__ X, int64_t rows, int S, int D, const float* __restrict__ sn,
                            const float* __restrict__ cs, TO* __restrict__ Y) {
  const int half = D >> 1;
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;       // one (row, pair) per thread
  if (idx >= rows * half) return;
  const int64_t row = idx / half;
  const int j = (int)(idx - row * half);
  const int p = (int)(row % S);
  const float s = __ldg(sn + (int64_t)p * half + j), c = __ldg(cs + (int64_t)p * half + j);
  auto ld = [&](int64_t i) -> float {
    if constexpr (sizeof(TI) == 2) return __bfloat162float(X[i]);
    else return X[i];
  };
  auto st = [&](int64_t i, float v) {
    if constexpr (sizeof(TO) == 2) Y[i] = __float2bfloat16(v);
    else Y[i] = v;
  };
  if constexpr (!INVERSE) {
    const float x1 = ld(row * D + 2 * j), x2 = ld(row * D + 2 * j + 1);
    st(row * D + j, x1 * c - x2 * s);
    st(row * D + half + j, x1 * s + x2 * c);
  } else {
    const float g1 = ld(row * D + j), g2 = ld(row * D + half + j);
    st(row * D + 2 * j, g1 * c + g2 * s);
    st(row * D + 2 * j + 1, g2 * c - g1 * s);
  }
}

// ---------------------------------------------------------------------------------------------- host side
constexpr int kColsumRowsPerBlock = 4096;
static inline int colsum_blocks(int64_t rows) { return (int)((rows + kColsumRowsPerBlock - 1) / kColsumRowsPerBlock); }

template <typename T>
static int colsum_launch(const __nv_bfloat16* G, const T* X, const float* mean, const float* rstd, int64_t rows, int D,
                         float* part, float* out_g, float* out_b, cudaStream_t st, const uint8_t* keep = nullptr,
                         float keep_scale = 1.f) {
  const int nblk = colsum_blocks(rows);
  dim3 grid((D + 255) / 256, nblk);
  colsum_part_kernel<T><<<grid, 256, 0, st>>>(G, X, mean, rstd, rows, D, kColsumRowsPerBlock, part, keep, keep_scale);
  colsum_reduce_kernel<<<(D + 255) / 256, 256, 0, st>>>(part, nblk, D, out_g, out_b);
  count_launches(2);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// dW [out, in] = G^T X over `rows` tokens (G bf16 [rows, out], X bf16 [rows, in]); split-K over token blocks into
// fp32 slabs, summed in a fixed order.
struct WgradPlan {
  int nsplit;        // splits = batches of the one launch
  int64_t ksplit;    // tokens per split (multiple of 64); the last split's tail is zero-filled by TMA
};
static WgradPlan wgrad_plan(int64_t rows, int out, int in) {
  const int tiles = ((out + 255) / 256) * ((in + 255) / 256);
  const int clusters = sm_count() / 2;
  int want = (6 * clusters + tiles - 1) / tiles;                 // about six waves of CTA pairs
  const int64_t max_split = rows / 1024 > 0 ? rows / 1024 : 1;   // at least 1024 tokens per split
  if (want > max_split) want = (int)max_split;
  if (want < 1) want = 1;
  WgradPlan p;
  p.ksplit = rup64((rows + want - 1) / want, 64);
  p.nsplit = (int)((rows + p.ksplit - 1) / p.ksplit);
  return p;
}
static size_t wgrad_ws_bytes(int64_t rows, int out, int in) {
  const WgradPlan p = wgrad_plan(rows, out, in);
  return (size_t)rup64((int64_t)p.nsplit * out * in * 4, 1024);
}
static int wgrad_splitk(const __nv_bfloat16* G, int out, const __nv_bfloat16* X, int in, int64_t rows, float* slabs,
                        float* dW, cudaStream_t st) {
  const WgradPlan pl = wgrad_plan(rows, out, in);
  CLIPK_REQUIRE(out % 8 == 0 && in % 8 == 0, "wgrad: feature dims must be multiples of 8 (out=%d in=%d)", out, in);
  {
    // one launch: batch s covers tokens [s * ksplit, (s + 1) * ksplit) of the single long K (= tokens); the maps keep
    // the true extent `rows`
    OperandDesc a, b;
    a.ptr = G; a.mn_major = true; a.rows = out; a.k = rows; a.ld = out; a.k_batch_offset = pl.ksplit;
    b.ptr = X; b.mn_major = true; b.rows = in; b.k = rows; b.ld = in;
    const int ks[1] = {(int)(pl.ksplit / 64)};
    epi::Store<false>::Params ep{slabs, in, (int64_t)out * in, out, in, 1.f, 0};
    if (in > 128) CLIPK_TRY((launch_gemm2<256, true, true, epi::Store<false>>(&a, &b, 1, ks, ks, out, in, pl.nsplit, ep, st)));
    else CLIPK_TRY((launch_gemm2<128, true, true, epi::Store<false>>(&a, &b, 1, ks, ks, out, in, pl.nsplit, ep, st)));
  }
  const int64_t n4 = (int64_t)out * in / 4;
  slab_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(slabs), pl.nsplit, n4,
                                                                   reinterpret_cast<float4*>(dW));
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

struct ProjBwdWs {
  __nv_bfloat16* dZ;   // [R, Dout]
  float* slabs;        // split-K slabs (largest of the three weight gradients)
  float* part;         // column-sum partials [nblk][2][max(Din, Dout)]
};
static size_t proj_bwd_carve(ProjBwdWs* w, void* base, int64_t R, int Din, int Dout) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  w->dZ = static_cast<__nv_bfloat16*>(take((size_t)R * Dout * 2));
  size_t sl = wgrad_ws_bytes(R, Dout, Din);
  const size_t sl3 = wgrad_ws_bytes(R, Dout, Dout);
  if (sl3 > sl) sl = sl3;
  w->slabs = static_cast<float*>(take(sl));
  const int dmax = Din > Dout ? Din : Dout;
  w->part = static_cast<float*>(take((size_t)colsum_blocks(R) * 2 * dmax * 4));
  return off;
}

static int check_rows(int64_t R) {
  CLIPK_REQUIRE(R > 0 && R < (int64_t)1 << 31, "proj: row count %lld out of range", (long long)R);
  return 0;
}

}  // namespace clipk

extern "C" {

// xn = LayerNorm(x) * gamma + beta as bf16; mean / rstd [rows] fp32 are saved for the backward.  x: bf16 | fp32.
int clipk_ln_fwd(const void* x, int dtype, int64_t rows, int D, const float* gamma, const float* beta, float eps,
                 void* xn, float* mean, float* rstd, float drop_p, uint64_t seed, void* keep_bits, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(rows >= 0 && D > 0 && D % 8 == 0, "ln_fwd: bad shape rows=%lld D=%d (D %% 8 == 0)", (long long)rows, D);
  CLIPK_REQUIRE(dtype == CLIPK_BF16 || dtype == CLIPK_F32, "ln_fwd: dtype must be bf16 or fp32");
  if (rows == 0) return 0;
  CLIPK_REQUIRE(drop_p >= 0.f && drop_p < 1.f, "ln_fwd: dropout probability %f out of [0, 1)", drop_p);
  CLIPK_REQUIRE(drop_p == 0.f || keep_bits != nullptr, "ln_fwd: dropout needs the keep-bit buffer (rows * D / 8 bytes)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  uint8_t* keep = drop_p > 0.f ? static_cast<uint8_t*>(keep_bits) : nullptr;
  const uint32_t thr = dropout_thr16(drop_p);
  if (dtype == CLIPK_BF16)
    ln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), rows, D, gamma, beta, eps,
                                                       static_cast<__nv_bfloat16*>(xn), mean, rstd, keep, thr, seed);
  else
    ln_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), rows, D, gamma, beta, eps,
                                               static_cast<__nv_bfloat16*>(xn), mean, rstd, keep, thr, seed);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t clipk_ln_bwd_workspace_bytes(int64_t rows, int D) {
  return (size_t)clipk::colsum_blocks(rows) * 2 * D * 4 + 1024;
}

// dgamma, dbeta [D] fp32 from dxn (bf16 [rows, D]); dx (nullable, same dtype as x) = LayerNorm input gradient.
int clipk_ln_bwd(const void* x, int dtype, int64_t rows, int D, const float* gamma, const float* mean,
                 const float* rstd, const void* dxn, float* dgamma, float* dbeta, void* dx, float drop_p,
                 const void* keep_bits, void* workspace, size_t ws_bytes, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(rows > 0 && D > 0 && D % 8 == 0, "ln_bwd: bad shape rows=%lld D=%d", (long long)rows, D);
  CLIPK_REQUIRE(dtype == CLIPK_BF16 || dtype == CLIPK_F32, "ln_bwd: dtype must be bf16 or fp32");
  CLIPK_REQUIRE(workspace != nullptr && ws_bytes >= clipk_ln_bwd_workspace_bytes(rows, D), "ln_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(dxn);
  float* part = static_cast<float*>(workspace);
  const unsigned grid = (unsigned)((rows + 7) / 8);
  CLIPK_REQUIRE(drop_p == 0.f || keep_bits != nullptr, "ln_bwd: dropout needs the keep bits of the forward");
  const uint8_t* keep = drop_p > 0.f ? static_cast<const uint8_t*>(keep_bits) : nullptr;
  const float ksc = dropout_scale(dropout_thr16(drop_p));
  if (dtype == CLIPK_BF16) {
    CLIPK_TRY(colsum_launch<__nv_bfloat16>(g, static_cast<const __nv_bfloat16*>(x), mean, rstd, rows, D, part, dgamma, dbeta, st,
                                           keep, ksc));
    if (dx != nullptr)
      ln_bwd_dx_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(g, static_cast<const __nv_bfloat16*>(x), gamma, mean, rstd,
                                                            rows, D, static_cast<__nv_bfloat16*>(dx), keep, ksc);
  } else {
    CLIPK_TRY(colsum_launch<float>(g, static_cast<const float*>(x), mean, rstd, rows, D, part, dgamma, dbeta, st, keep, ksc));
    if (dx != nullptr)
      ln_bwd_dx_kernel<float><<<grid, 256, 0, st>>>(g, static_cast<const float*>(x), gamma, mean, rstd, rows, D,
                                                    static_cast<float*>(dx), keep, ksc);
  }
  if (dx != nullptr) count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// apply_rope (pacl.py:147-181) on rows of [B*S, D]; sin / cos: fp32 [S, D/2] (built by the caller exactly as the
// reference does).  dtype_in / dtype_out: bf16 | fp32.  inverse != 0: the transposed rotation (gradient).
int clipk_rope(const void* x, int dtype_in, int64_t rows, int S, int D, const float* sin_t, const float* cos_t, void* y,
               int dtype_out, int inverse, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(rows >= 0 && S > 0 && D > 0 && D % 2 == 0, "rope: bad shape rows=%lld S=%d D=%d (D even)", (long long)rows, S, D);
  CLIPK_REQUIRE((dtype_in == CLIPK_BF16 || dtype_in == CLIPK_F32) && (dtype_out == CLIPK_BF16 || dtype_out == CLIPK_F32),
                "rope: dtypes must be bf16 or fp32");
  if (rows == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int64_t n = rows * (D / 2);
  const unsigned grid = (unsigned)((n + 255) / 256);
#define CLIPK_ROPE(TI, TO, INV) \
  rope_kernel<TI, TO, INV><<<grid, 256, 0, st>>>(static_cast<const TI*>(x), rows, S, D, sin_t, cos_t, static_cast<TO*>(y))
  const bool ib = dtype_in == CLIPK_BF16, ob = dtype_out == CLIPK_BF16;
  if (!inverse) {
    if (ib && ob) CLIPK_ROPE(__nv_bfloat16, __nv_bfloat16, false);
    else if (ib) CLIPK_ROPE(__nv_bfloat16, float, false);
    else if (ob) CLIPK_ROPE(float, __nv_bfloat16, false);
    else CLIPK_ROPE(float, float, false);
  } else {
    if (ib && ob) CLIPK_ROPE(__nv_bfloat16, __nv_bfloat16, true);
    else if (ib) CLIPK_ROPE(__nv_bfloat16, float, true);
    else if (ob) CLIPK_ROPE(float, __nv_bfloat16, true);
    else CLIPK_ROPE(float, float, true);
  }
#undef CLIPK_ROPE
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Patch_Projection forward (pacl.py:35-48).  xn bf16 [R, Din]; W1, W2 bf16 [Dout, Din]; W3 bf16 [Dout, Dout];
// b13 = b1 + b3, b2: fp32 [Dout].  Outputs (bf16 [R, Dout]): Gp = gelu'(z), H = gelu(z) (both saved for the
// backward), Y.
int clipk_patch_proj_fwd(const void* xn, int64_t R, int Din, int Dout, const void* W1, const void* W2, const void* W3,
                         const float* b13, const float* b2, void* Gp, void* H, void* Y, float* ysq, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_TRY(check_rows(R));
  CLIPK_REQUIRE(Din > 0 && Dout > 0 && Din % 8 == 0 && Dout % 8 == 0, "patch_proj_fwd: dims must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = (int)R;
  {
    OperandDesc a, b;
    a.ptr = xn; a.rows = M; a.k = Din; a.ld = Din;
    b.ptr = W2; b.rows = Dout; b.k = Din; b.ld = Din;
    const int ks[1] = {(Din + 63) / 64};
    const eng::OutDesc oh{H, Dout, (int64_t)M * Dout, M, Dout, 1}, oz{Gp, Dout, (int64_t)M * Dout, M, Dout, 1};
    epi::BiasGelu2::Params ep{oh, oz, b2, Dout};
    CLIPK_TRY((launch_gemm2<256, false, false, epi::BiasGelu2>(&a, &b, 1, ks, ks, M, Dout, 1, ep, st)));
  }
  {
    OperandDesc a[2], b[2];
    a[0].ptr = xn; a[0].rows = M; a[0].k = Din; a[0].ld = Din;
    b[0].ptr = W1; b[0].rows = Dout; b[0].k = Din; b[0].ld = Din;
    a[1].ptr = H; a[1].rows = M; a[1].k = Dout; a[1].ld = Dout;
    b[1].ptr = W3; b[1].rows = Dout; b[1].k = Dout; b[1].ld = Dout;
    const int ks[2] = {(Din + 63) / 64, (Dout + 63) / 64};
    if (ysq != nullptr) CLIPK_CHECK_CUDA(cudaMemsetAsync(ysq, 0, (size_t)M * 4, st));
    epi::BiasTma::Params ep{{Y, Dout, (int64_t)M * Dout, M, Dout, 1}, b13, Dout, ysq, M};
    CLIPK_TRY((launch_gemm2<256, false, false, epi::BiasTma>(a, b, 2, ks, ks, M, Dout, 1, ep, st)));
  }
  return 0;
}

size_t clipk_patch_proj_bwd_workspace_bytes(int64_t R, int Din, int Dout) {
  clipk::ProjBwdWs w{};
  return clipk::proj_bwd_carve(&w, nullptr, R, Din, Dout);
}

// Patch_Projection backward.  dY bf16 [R, Dout].  dxn (nullable) bf16 [R, Din]; dW1, dW2 [Dout, Din], dW3 [Dout, Dout],
// db13 (= db1 = db3), db2 [Dout]: fp32, overwritten.
int clipk_patch_proj_bwd(const void* xn, const void* Gp, const void* H, const void* dY, int64_t R, int Din, int Dout,
                         const void* W1, const void* W2, const void* W3, void* dxn, float* dW1, float* dW2, float* dW3,
                         float* db13, float* db2, void* workspace, size_t ws_bytes, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_TRY(check_rows(R));
  CLIPK_REQUIRE(Din > 0 && Dout > 0 && Din % 8 == 0 && Dout % 8 == 0, "patch_proj_bwd: dims must be multiples of 8");
  ProjBwdWs w{};
  const size_t need = proj_bwd_carve(&w, workspace, R, Din, Dout);
  CLIPK_REQUIRE(workspace != nullptr && ws_bytes >= need, "patch_proj_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = (int)R;
  const __nv_bfloat16* dy = static_cast<const __nv_bfloat16*>(dY);
  const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(xn);
  // dZ = (dY W3) * G'        B[n][k] = W3[k][n]: W3 read as an MN-major operand
  {
    OperandDesc a, b;
    a.ptr = dY; a.rows = M; a.k = Dout; a.ld = Dout;
    b.ptr = W3; b.mn_major = true; b.rows = Dout; b.k = Dout; b.ld = Dout;
    const int ks[1] = {(Dout + 63) / 64};
    const eng::OutDesc od{w.dZ, Dout, (int64_t)M * Dout, M, Dout, 1};
    const eng::OutDesc oz{const_cast<void*>(Gp), Dout, (int64_t)M * Dout, M, Dout, 1};
    epi::GeluBwdIn::Params ep{od, oz};
    if (Dout > 128) CLIPK_TRY((launch_gemm2<256, false, true, epi::GeluBwdIn>(&a, &b, 1, ks, ks, M, Dout, 1, ep, st)));
    else CLIPK_TRY((launch_gemm2<128, false, true, epi::GeluBwdIn>(&a, &b, 1, ks, ks, M, Dout, 1, ep, st)));
  }
  // dxn = dY W1 + dZ W2
  if (dxn != nullptr) {
    OperandDesc a[2], b[2];
    a[0].ptr = dY; a[0].rows = M; a[0].k = Dout; a[0].ld = Dout;
    b[0].ptr = W1; b[0].mn_major = true; b[0].rows = Din; b[0].k = Dout; b[0].ld = Din;
    a[1].ptr = w.dZ; a[1].rows = M; a[1].k = Dout; a[1].ld = Dout;
    b[1].ptr = W2; b[1].mn_major = true; b[1].rows = Din; b[1].k = Dout; b[1].ld = Din;
    const int ks[2] = {(Dout + 63) / 64, (Dout + 63) / 64};
    epi::StoreTma::Params ep{{dxn, Din, (int64_t)M * Din, M, Din, 1}, 1.f};
    if (Din > 128) CLIPK_TRY((launch_gemm2<256, false, true, epi::StoreTma>(a, b, 2, ks, ks, M, Din, 1, ep, st)));
    else CLIPK_TRY((launch_gemm2<128, false, true, epi::StoreTma>(a, b, 2, ks, ks, M, Din, 1, ep, st)));
  }
  // weight gradients (split-K over tokens)
  CLIPK_TRY(wgrad_splitk(dy, Dout, x, Din, R, w.slabs, dW1, st));
  CLIPK_TRY(wgrad_splitk(w.dZ, Dout, x, Din, R, w.slabs, dW2, st));
  CLIPK_TRY(wgrad_splitk(dy, Dout, static_cast<const __nv_bfloat16*>(H), Dout, R, w.slabs, dW3, st));
  // bias gradients
  CLIPK_TRY(colsum_launch<float>(dy, nullptr, nullptr, nullptr, R, Dout, w.part, nullptr, db13, st));
  CLIPK_TRY(colsum_launch<float>(w.dZ, nullptr, nullptr, nullptr, R, Dout, w.part, nullptr, db2, st));
  return 0;
}

// Linear forward / backward on bf16 rows (text_projection's nn.Linear, pacl.py:78): y = x W^T + b.
int clipk_linear_fwd(const void* x, int64_t R, int Din, int Dout, const void* W, const float* bias, void* y,
                     void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_TRY(check_rows(R));
  CLIPK_REQUIRE(Din > 0 && Dout > 0 && Din % 8 == 0 && Dout % 8 == 0, "linear_fwd: dims must be multiples of 8");
  OperandDesc a, b;
  a.ptr = x; a.rows = R; a.k = Din; a.ld = Din;
  b.ptr = W; b.rows = Dout; b.k = Din; b.ld = Din;
  const int ks[1] = {(Din + 63) / 64};
  const int M = (int)R;
  epi::BiasTma::Params ep{{y, Dout, (int64_t)M * Dout, M, Dout, 1}, bias, Dout};
  return launch_gemm2<256, false, false, epi::BiasTma>(&a, &b, 1, ks, ks, M, Dout, 1, ep, static_cast<cudaStream_t>(stream));
}

size_t clipk_linear_bwd_workspace_bytes(int64_t R, int Din, int Dout) {
  return clipk::wgrad_ws_bytes(R, Dout, Din) + (size_t)clipk::colsum_blocks(R) * 2 * Dout * 4 + 2048;
}

int clipk_linear_bwd(const void* x, const void* dY, int64_t R, int Din, int Dout, const void* W, void* dx, float* dW,
                     float* db, void* workspace, size_t ws_bytes, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_TRY(check_rows(R));
  CLIPK_REQUIRE(Din > 0 && Dout > 0 && Din % 8 == 0 && Dout % 8 == 0, "linear_bwd: dims must be multiples of 8");
  CLIPK_REQUIRE(workspace != nullptr && ws_bytes >= clipk_linear_bwd_workspace_bytes(R, Din, Dout), "linear_bwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int M = (int)R;
  float* slabs = static_cast<float*>(workspace);
  float* part = reinterpret_cast<float*>(static_cast<char*>(workspace) + (wgrad_ws_bytes(R, Dout, Din) + 1023) / 1024 * 1024);
  if (dx != nullptr) {
    OperandDesc a, b;
    a.ptr = dY; a.rows = M; a.k = Dout; a.ld = Dout;
    b.ptr = W; b.mn_major = true; b.rows = Din; b.k = Dout; b.ld = Din;
    const int ks[1] = {(Dout + 63) / 64};
    epi::StoreTma::Params ep{{dx, Din, (int64_t)M * Din, M, Din, 1}, 1.f};
    if (Din > 128) CLIPK_TRY((launch_gemm2<256, false, true, epi::StoreTma>(&a, &b, 1, ks, ks, M, Din, 1, ep, st)));
    else CLIPK_TRY((launch_gemm2<128, false, true, epi::StoreTma>(&a, &b, 1, ks, ks, M, Din, 1, ep, st)));
  }
  CLIPK_TRY(wgrad_splitk(static_cast<const __nv_bfloat16*>(dY), Dout, static_cast<const __nv_bfloat16*>(x), Din, R, slabs, dW, st));
  CLIPK_TRY(colsum_launch<float>(static_cast<const __nv_bfloat16*>(dY), nullptr, nullptr, nullptr, R, Dout, part, nullptr, db, st));
  return 0;
}

}  // extern "C"
