// PACL paired path (reference training / eval semantics, SURVEY §8 a1+a2): one (image, text) pair per sample,
// bandwidth-bound.  One CTA per sample streams V_b [P, D] exactly once in forward (norm, score, activation,
// weighted pooling and the final L2-normalisation fused) and once in backward (re-derives score/activation from
// the same registers, writes dV, reduces dT).  Algorithmic bytes: 3 * B*P*D * sizeof(V) (+ O(B*D)).
//
// Replaces PACL/model/pacl.py:120-133 (patch_alignment) and :140-145 (pooling + F.normalize) and their autograd.
#include "common.cuh"
#include "simt_util.cuh"

namespace clipk {

constexpr int kPairedThreads = 256;
constexpr int kPairedWarps = kPairedThreads / 32;

// NIT = ceil(D / 256): each lane owns 8 consecutive elements at offset lane*8 + it*256.
template <class T, int NIT>
__global__ void __launch_bounds__(kPairedThreads)
pacl_paired_fwd_kernel(const T* __restrict__ V, const T* __restrict__ Tx, int B, int v_div, int P, int D, int act,
                       float* __restrict__ act_out, float* __restrict__ img_feat, float* __restrict__ txt_feat,
                       float* __restrict__ cosine, float* __restrict__ stats) {
  extern __shared__ float sm[];
  float* th = sm;                 // [D] normalised text
  float* ured = sm + NIT * 256;   // [warps][NIT*256]
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int DP = NIT * 256;

  // ---- text: t^ = t / max(||t||, 1e-12)
  float tsq = 0.f;
  for (int d = threadIdx.x; d < DP; d += blockDim.x) {
    const float x = d < D ? simt::to_f(Tx[(int64_t)b * D + d]) : 0.f;
    th[d] = x;
    tsq += x * x;
  }
  tsq = simt::block_sum(tsq, red);
  const float tnorm = sqrtf(tsq);
  const float rt = 1.f / fmaxf(tnorm, 1e-12f);
  __syncthreads();
  for (int d = threadIdx.x; d < DP; d += blockDim.x) th[d] *= rt;
  __syncthreads();

  float tl[NIT][8];
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) tl[it][j] = th[it * 256 + lane * 8 + j];

  float u[NIT][8];
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) u[it][j] = 0.f;

  const T* Vb = V + (int64_t)(b / v_div) * P * D;
  auto load_row = [&](int p, float (&v)[NIT][8]) {
    const T* row = Vb + (int64_t)p * D;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int d0 = it * 256 + lane * 8;
      if (d0 < D && p < P) simt::load8<T>(row + d0, v[it]);
      else {
#pragma unroll
        for (int j = 0; j < 8; ++j) v[it][j] = 0.f;
      }
    }
  };
  auto consume_row = [&](int p, const float (&v)[NIT][8]) {
    float r = 0.f, nsq = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r = fmaf(tl[it][j], v[it][j], r);
        nsq = fmaf(v[it][j], v[it][j], nsq);
      }
    r = ptx::warp_sum(r);
    nsq = ptx::warp_sum(nsq);
    const float rn = 1.f / fmaxf(sqrtf(nsq), 1e-12f);
    const float s = r * rn;
    const float sig = act == CLIPK_ACT_SOFTMAX10 ? expf(10.f * (s - 1.f)) : 1.f / (1.f + expf(-10.f * s));
    if (act_out != nullptr && lane == 0) act_out[(int64_t)b * P + p] = sig;
    const float a = (act == CLIPK_ACT_ONES) ? 1.f : sig;
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) u[it][j] = fmaf(a, v[it][j], u[it][j]);
  };
  // two rows per warp in flight: the occupancy is register-bound (2 CTAs / SM), so the bytes in flight per warp are
  // what sets the achieved bandwidth
  for (int p = warp; p < P; p += 2 * kPairedWarps) {
    float v0[NIT][8], v1[NIT][8];
    load_row(p, v0);
    load_row(p + kPairedWarps, v1);
    consume_row(p, v0);
    if (p + kPairedWarps < P) consume_row(p + kPairedWarps, v1);
  }
  // ---- cross-warp reduction of the pooled vector
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) ured[warp * DP + it * 256 + lane * 8 + j] = u[it][j];
  __syncthreads();
  float usq = 0.f, dotp = 0.f;
  for (int d = threadIdx.x; d < DP; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kPairedWarps; ++w) acc += ured[w * DP + d];
    ured[d] = acc;      // warp 0's slot now holds the sum (each thread touches only its own d)
    usq += acc * acc;
    dotp += acc * th[d];
  }
  usq = simt::block_sum(usq, red);
  dotp = simt::block_sum(dotp, red);
  const float unorm = sqrtf(usq);
  const float ru = 1.f / fmaxf(unorm, 1e-12f);
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    img_feat[(int64_t)b * D + d] = ured[d] * ru;
    txt_feat[(int64_t)b * D + d] = th[d];
  }
  if (threadIdx.x == 0) {
    if (cosine != nullptr) cosine[b] = dotp * ru;
    stats[2 * b] = unorm;
    stats[2 * b + 1] = tnorm;
  }
}

template <class T, int NIT>
__global__ void __launch_bounds__(kPairedThreads, NIT <= 3 ? 2 : 1)
pacl_paired_bwd_kernel(const T* __restrict__ V, const T* __restrict__ Tx, int B, int P, int D, int act,
                       const float* __restrict__ img_feat, const float* __restrict__ stats,
                       const float* __restrict__ d_img, const float* __restrict__ d_txt, T* __restrict__ dV,
                       T* __restrict__ dT) {
  extern __shared__ float sm[];
  const int DP = NIT * 256;
  float* th = sm;              // [DP] t^
  float* gu = sm + DP;         // [DP] grad wrt the un-normalised pooled vector u
  float* tred = sm + 2 * DP;   // [warps][DP]
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const float unorm = stats[2 * b], tnorm = stats[2 * b + 1];
  const float ru = 1.f / fmaxf(unorm, 1e-12f), rt = 1.f / fmaxf(tnorm, 1e-12f);

  // g_u = (d_img - u^ <u^, d_img>) / ||u||      (Jacobian of F.normalize, pacl.py:145)
  float dotg = 0.f;
  for (int d = threadIdx.x; d < DP; d += blockDim.x) {
    const float uh = d < D ? img_feat[(int64_t)b * D + d] : 0.f;
    const float g = d < D ? d_img[(int64_t)b * D + d] : 0.f;
    th[d] = d < D ? simt::to_f(Tx[(int64_t)b * D + d]) * rt : 0.f;
    gu[d] = g;
    tred[d] = uh;
    dotg += uh * g;
  }
  dotg = simt::block_sum(dotg, red);
  // a clamped norm (||u|| < eps) has zero Jacobian projection in torch as well: x / eps is linear there
  const bool clamped_u = unorm < 1e-12f;
  __syncthreads();
  for (int d = threadIdx.x; d < DP; d += blockDim.x) gu[d] = (gu[d] - (clamped_u ? 0.f : tred[d] * dotg)) * ru;
  __syncthreads();

  float tl[NIT][8], gl[NIT][8], dth[NIT][8];
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      tl[it][j] = th[it * 256 + lane * 8 + j];
      gl[it][j] = gu[it * 256 + lane * 8 + j];
      dth[it][j] = 0.f;
    }

  const T* Vb = V + (int64_t)b * P * D;
  T* dVb = dV + (int64_t)b * P * D;
  // two rows per warp in flight: the next row's loads are issued (raw, unconverted) before this row is processed --
  // the occupancy is register-bound (2 CTAs / SM), so the bytes in flight per warp set the achieved bandwidth
  simt::Raw8<T> nxt[NIT];
  auto issue = [&](int p) {
    const T* row = Vb + (int64_t)p * D;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int d0 = it * 256 + lane * 8;
      nxt[it] = (d0 < D && p < P) ? simt::load_raw8<T>(row + d0) : simt::zero_raw8<T>();
    }
  };
  issue(warp);
  for (int p = warp; p < P; p += kPairedWarps) {
    float v[NIT][8];
#pragma unroll
    for (int it = 0; it < NIT; ++it) simt::unpack_raw8(nxt[it], v[it]);
    issue(p + kPairedWarps);
    float r = 0.f, nsq = 0.f, da = 0.f;
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        r = fmaf(tl[it][j], v[it][j], r);
        nsq = fmaf(v[it][j], v[it][j], nsq);
        da = fmaf(gl[it][j], v[it][j], da);
      }
    r = ptx::warp_sum(r);
    nsq = ptx::warp_sum(nsq);
    da = ptx::warp_sum(da);
    const float vn = sqrtf(nsq);
    const float rn = 1.f / fmaxf(vn, 1e-12f);
    const float s = r * rn;
    float a = 1.f, ds = 0.f;
    if (act == CLIPK_ACT_SOFTMAX10) {
      a = expf(10.f * (s - 1.f));
      ds = da * 10.f * a;
    } else if (act != CLIPK_ACT_ONES) {
      a = 1.f / (1.f + expf(-10.f * s));
      ds = da * 10.f * a * (1.f - a);
    }
    // dV_p = a g_u + rn ds (t^ - v^ s)  with v^ = rn v
    const float c_t = ds * rn;
    const float c_v = (vn < 1e-12f) ? 0.f : -ds * rn * rn * s;
    float o[NIT][8];
#pragma unroll
    for (int it = 0; it < NIT; ++it)
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        o[it][j] = fmaf(a, gl[it][j], fmaf(c_t, tl[it][j], c_v * v[it][j]));
        dth[it][j] = fmaf(c_t, v[it][j], dth[it][j]);
      }
    T* orow = dVb + (int64_t)p * D;
#pragma unroll
    for (int it = 0; it < NIT; ++it) {
      const int d0 = it * 256 + lane * 8;
      if (d0 < D) simt::store8<T>(orow + d0, o[it]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int it = 0; it < NIT; ++it)
#pragma unroll
    for (int j = 0; j < 8; ++j) tred[warp * DP + it * 256 + lane * 8 + j] = dth[it][j];
  __syncthreads();
  // dt^ (total) = sum_p ds_p v^_p + d_txt ;  dT = (dt^ - t^ <t^, dt^>) / ||t||
  float dott = 0.f;
  for (int d = threadIdx.x; d < DP; d += blockDim.x) {
    float acc = 0.f;
#pragma unroll
    for (int w = 0; w < kPairedWarps; ++w) acc += tred[w * DP + d];
    if (d < D && d_txt != nullptr) acc += d_txt[(int64_t)b * D + d];
    tred[d] = acc;
    dott += acc * th[d];
  }
  dott = simt::block_sum(dott, red);
  const bool clamped_t = tnorm < 1e-12f;
  for (int d = threadIdx.x; d < D; d += blockDim.x)
    dT[(int64_t)b * D + d] = simt::from_f<T>((tred[d] - (clamped_t ? 0.f : th[d] * dott)) * rt);
}

template <class T>
static int paired_fwd_t(const void* V, const void* Tx, int B, int v_div, int P, int D, int act, float* act_out,
                        float* img, float* txt, float* cosine, float* stats, cudaStream_t st) {
  const int nit = (D + 255) / 256;
  const size_t smem = (size_t)(nit * 256) * (1 + kPairedWarps) * sizeof(float);
#define LAUNCH(N)                                                                                              \
  {                                                                                                            \
    auto k = pacl_paired_fwd_kernel<T, N>;                                                                     \
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    k<<<B, kPairedThreads, smem, st>>>((const T*)V, (const T*)Tx, B, v_div, P, D, act, act_out, img, txt, cosine, stats); \
    clipk::count_launches(1); \
  }
  switch (nit) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    case 4: LAUNCH(4) break;
    default: set_error("pacl_paired: D=%d unsupported (D <= 1024)", D); return CLIPK_ERR_INVALID;
  }
#undef LAUNCH
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

template <class T>
static int paired_bwd_t(const void* V, const void* Tx, int B, int P, int D, int act, const float* img,
                        const float* stats, const float* d_img, const float* d_txt, void* dV, void* dT,
                        cudaStream_t st) {
  const int nit = (D + 255) / 256;
  const size_t smem = (size_t)(nit * 256) * (2 + kPairedWarps) * sizeof(float);
#define LAUNCH(N)                                                                                              \
  {                                                                                                            \
    auto k = pacl_paired_bwd_kernel<T, N>;                                                                     \
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    k<<<B, kPairedThreads, smem, st>>>((const T*)V, (const T*)Tx, B, P, D, act, img, stats, d_img, d_txt, (T*)dV, (T*)dT); \
    clipk::count_launches(1); \
  }
  switch (nit) {
    case 1: LAUNCH(1) break;
    case 2: LAUNCH(2) break;
    case 3: LAUNCH(3) break;
    case 4: LAUNCH(4) break;
    default: set_error("pacl_paired: D=%d unsupported (D <= 1024)", D); return CLIPK_ERR_INVALID;
  }
#undef LAUNCH
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace clipk

extern "C" {

int clipk_pacl_paired_fwd(const void* V, const void* T, int dtype, int B, int v_div, int P, int D, int act,
                          float* act_out, float* img_feat, float* txt_feat, float* cosine, float* stats,
                          void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(B >= 0 && P > 0 && D > 0 && v_div >= 1, "pacl_paired_fwd: bad shape B=%d P=%d D=%d v_div=%d", B, P, D, v_div);
  CLIPK_REQUIRE(D % 8 == 0, "pacl_paired_fwd: D=%d must be a multiple of 8", D);
  CLIPK_REQUIRE(act >= CLIPK_ACT_SIGMOID10 && act <= CLIPK_ACT_SOFTMAX10, "pacl_paired_fwd: bad activation %d", act);
  if (B == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CLIPK_BF16)
    return clipk::paired_fwd_t<__nv_bfloat16>(V, T, B, v_div, P, D, act, act_out, img_feat, txt_feat, cosine, stats, st);
  if (dtype == CLIPK_F32)
    return clipk::paired_fwd_t<float>(V, T, B, v_div, P, D, act, act_out, img_feat, txt_feat, cosine, stats, st);
  if (dtype == CLIPK_F16)
    return clipk::paired_fwd_t<__half>(V, T, B, v_div, P, D, act, act_out, img_feat, txt_feat, cosine, stats, st);
  clipk::set_error("pacl_paired_fwd: bad dtype %d", dtype);
  return CLIPK_ERR_INVALID;
}

int clipk_pacl_paired_bwd(const void* V, const void* T, int dtype, int B, int P, int D, int act,
                          const float* img_feat, const float* stats, const float* d_img, const float* d_txt,
                          void* dV, void* dT, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(B >= 0 && P > 0 && D > 0, "pacl_paired_bwd: bad shape B=%d P=%d D=%d", B, P, D);
  CLIPK_REQUIRE(D % 8 == 0, "pacl_paired_bwd: D=%d must be a multiple of 8", D);
  CLIPK_REQUIRE(act >= CLIPK_ACT_SIGMOID10 && act <= CLIPK_ACT_SOFTMAX10, "pacl_paired_bwd: bad activation %d", act);
  if (B == 0) return 0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CLIPK_BF16)
    return clipk::paired_bwd_t<__nv_bfloat16>(V, T, B, P, D, act, img_feat, stats, d_img, d_txt, dV, dT, st);
  if (dtype == CLIPK_F32)
    return clipk::paired_bwd_t<float>(V, T, B, P, D, act, img_feat, stats, d_img, d_txt, dV, dT, st);
  if (dtype == CLIPK_F16)
    return clipk::paired_bwd_t<__half>(V, T, B, P, D, act, img_feat, stats, d_img, d_txt, dV, dT, st);
  clipk::set_error("pacl_paired_bwd: bad dtype %d", dtype);
  return CLIPK_ERR_INVALID;
}
}
