// SPARC token-to-patch alignment and the SparcLoss local term.
//
//   sparc.forward        PACL/model/pacl.py:453-478   S = L V^T (raw) -> min-max -> threshold sigma -> row-normalise
//                                                      -> G = W V -> n(L), n(G)
//   masked pairwise CE   PACL/model/pacl.py:522-556   per-sample T x T logits, -1e8 column mask, mask-weighted mean
//
// The two contractions per direction run on the tcgen05 engine (tokens padded to the 128-row MMA tile by TMA
// zero-fill); the row statistics / Jacobians are small CUDA-core kernels.  Backward recomputes S and W
// (nothing of size [B,T,P] is saved between forward and backward).
#include "common.cuh"
#include "epilogues.cuh"
#include "simt_util.cuh"
#include "sparc_fused.cuh"

namespace epi {
// dV[b][m][n] = acc + gadd[b][n]      (gadd: broadcast gradient of the mean-pooled global feature, may be null)
struct SparcDv {
  struct Params {
    const float* gadd;   // [batch][N] or nullptr
    void* dV;            // [batch][M][N], bf16 or fp32
    int M, N, out_bf16;
  };
  Params p;
  __device__ explicit SparcDv(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ void chunk(int b, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
    const int valid = min(32, p.N - n);
    if (p.gadd != nullptr) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < valid) v[j] += __ldg(p.gadd + (int64_t)b * p.N + n + j);
    }
    const int64_t off = ((int64_t)b * p.M + m) * p.N + n;
    if (p.out_bf16) store_bf16x32(reinterpret_cast<__nv_bfloat16*>(p.dV) + off, v, valid);
    else store_f32x32(reinterpret_cast<float*>(p.dV) + off, v, valid);
  }
  __device__ void tile_end(int, int, int, int, int) {}
};
// The same as a TMA-store functor (bf16 dV with 16-byte aligned rows, CTA-pair engine): lane l holds gadd[b][n + l] of
// the chunk (loaded one chunk ahead), broadcast by shuffle.
struct SparcDvTma {
  static constexpr bool kTmaOut = true;
  using Side = float;
  struct Params {
    eng::OutDesc out;    // dV [batch][M][N] bf16
    const float* gadd;   // [batch][N] or nullptr
    int N;
  };
  Params p;
  __device__ explicit SparcDvTma(const Params& pp) : p(pp) {}
  __device__ void tile_begin(int, int, int) {}
  __device__ Side pre(int b, int, int n) const {
    const int lane = (int)ptx::lane_id();
    return (p.gadd != nullptr && n + lane < p.N) ? __ldg(p.gadd + (int64_t)b * p.N + n + lane) : 0.f;
  }
  __device__ void chunk(int, int, int, float* v, const Side& g_l) {
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] += __shfl_sync(0xffffffffu, g_l, j);
  }
  __device__ void tile_end(int, int, int, int, int) {}
};
}  // namespace epi

namespace clipk {

static inline int rup(int x, int m) { return (x + m - 1) / m * m; }

// ---------------------------------------------------------------------------------------------- row kernels
// One warp per (b,t) row of S [rows, ld] (P valid columns).  Forward: min / max / threshold / row-normalise.
//   stats[row] = (min, R = max - min + 1e-8, Zeps = sum + 1e-8, _), arg[row] = (argmin, argmax)  (first occurrence)
__global__ void sparc_rows_fwd_kernel(const float* __restrict__ S, int64_t rows, int P, int ld, float sigma,
                                           __nv_bfloat16* __restrict__ W, float4* __restrict__ stats,
                                           int2* __restrict__ arg) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* s = S + row * ld;
  float mn = INFINITY, mx = -INFINITY;
  int imn = 0x7fffffff, imx = 0x7fffffff;
  for (int p = lane; p < P; p += 32) {
    const float v = s[p];
    if (v < mn) { mn = v; imn = p; }
    if (v > mx) { mx = v; imx = p; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float omn = __shfl_xor_sync(0xffffffffu, mn, o);
    const int oimn = __shfl_xor_sync(0xffffffffu, imn, o);
    if (omn < mn || (omn == mn && oimn < imn)) { mn = omn; imn = oimn; }
    const float omx = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oimx = __shfl_xor_sync(0xffffffffu, imx, o);
    if (omx > mx || (omx == mx && oimx < imx)) { mx = omx; imx = oimx; }
  }
  const float R = mx - mn + 1e-8f;
  float z = 0.f;
  for (int p = lane; p < P; p += 32) {
    float h = (s[p] - mn) / R;
    h = (h < sigma) ? 0.f : h;
    z += h;
  }
  z = ptx::warp_sum(z);
  const float zeps = z + 1e-8f;
  __nv_bfloat16* w = W + row * ld;
  for (int p = lane; p < ld; p += 32) {
    float h = 0.f;
    if (p < P) {
      h = (s[p] - mn) / R;
      h = (h < sigma) ? 0.f : h / zeps;
    }
    w[p] = __float2bfloat16(h);
  }
  if (lane == 0) {
    stats[row] = make_float4(mn, R, zeps, 0.f);
    arg[row] = make_int2(imn, imx);
  }
}

// Backward through row-normalise / threshold / min-max:  dW [rows, ld] fp32 -> dS [rows, ld] bf16.
//   c = sum_q dW_q W_q ;  dh_p = [h_p >= sigma] (dW_p - c) / zeps ;  dS_p = dh_p / R ;
//   dmin = -sum dh/R + sum dh (S-mn)/R^2 ; dmax = -sum dh (S-mn)/R^2  (added at the arg positions)
__global__ void sparc_rows_bwd_kernel(const float* __restrict__ S, const float* __restrict__ dW, int64_t rows, int P,
                                           int ld, float sigma, const float4* __restrict__ stats,
                                           const int2* __restrict__ arg, __nv_bfloat16* __restrict__ dS) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const float* s = S + row * ld;
  const float* dw = dW + row * ld;
  const float4 st = stats[row];
  const float mn = st.x, R = st.y, zeps = st.z;
  const int2 ag = arg[row];
  float c = 0.f;
  for (int p = lane; p < P; p += 32) {
    float h = (s[p] - mn) / R;
    h = (h < sigma) ? 0.f : h;
    c = fmaf(dw[p], h / zeps, c);
  }
  c = ptx::warp_sum(c);
  float a1 = 0.f, a2 = 0.f;
  for (int p = lane; p < P; p += 32) {
    const float h = (s[p] - mn) / R;
    const float dh = (h < sigma) ? 0.f : (dw[p] - c) / zeps;
    a1 += dh;
    a2 = fmaf(dh, h, a2);
  }
  a1 = ptx::warp_sum(a1);
  a2 = ptx::warp_sum(a2);
  const float dmin = (-a1 + a2) / R;
  const float dmax = -a2 / R;
  __nv_bfloat16* o = dS + row * ld;
  for (int p = lane; p < ld; p += 32) {
    float g = 0.f;
    if (p < P) {
      const float h = (s[p] - mn) / R;
      g = (h < sigma) ? 0.f : (dw[p] - c) / (zeps * R);
      if (p == ag.x) g += dmin;
      if (p == ag.y) g += dmax;
    }
    o[p] = __float2bfloat16(g);
  }
}

// x^ = x / max(|x|, 1e-12) for rows of [rows, D]; writes fp32 and (optionally) bf16 copies and the norms.
template <class T>
__global__ void normalize_rows_kernel(const T* __restrict__ X, int64_t rows, int D, float* __restrict__ out,
                                      __nv_bfloat16* __restrict__ out_bf16, float* __restrict__ norm) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  const T* x = X + row * D;
  float acc = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float v = simt::to_f(x[d]);
    acc = fmaf(v, v, acc);
  }
  acc = ptx::warp_sum(acc);
  const float nr = sqrtf(acc);
  const float r = 1.f / fmaxf(nr, 1e-12f);
  for (int d = lane; d < D; d += 32) {
    const float v = simt::to_f(x[d]) * r;
    if (out != nullptr) out[row * D + d] = v;
    if (out_bf16 != nullptr) out_bf16[row * D + d] = __float2bfloat16(v);
  }
  if (lane == 0 && norm != nullptr) norm[row] = nr;
}

// dx = (g - x^ <x^, g>) / |x|   (x^ fp32, g fp32) ; writes fp32 (out, optional, += if accumulate) and bf16 copies
__global__ void normalize_rows_bwd_kernel(const float* __restrict__ xh, const float* __restrict__ g,
                                          const float* __restrict__ norm, int64_t rows, int D, float* __restrict__ out,
                                          int accumulate, __nv_bfloat16* __restrict__ out_bf16) {
  const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const int lane = threadIdx.x & 31;
  float dot = 0.f;
  for (int d = lane; d < D; d += 32) dot = fmaf(xh[row * D + d], g[row * D + d], dot);
  dot = ptx::warp_sum(dot);
  const float nr = norm[row];
  const float r = 1.f / fmaxf(nr, 1e-12f);
  const float proj = nr < 1e-12f ? 0.f : dot;
  for (int d = lane; d < D; d += 32) {
    const float v = (g[row * D + d] - xh[row * D + d] * proj) * r;
    if (out != nullptr) out[row * D + d] = accumulate ? out[row * D + d] + v : v;
    if (out_bf16 != nullptr) out_bf16[row * D + d] = __float2bfloat16(v);
  }
}

// mean over dim 1 of X [B, R, D] -> [B, D]   (one block per (b, 256-wide d slice))
template <class T>
__global__ void mean_dim1_kernel(const T* __restrict__ X, int R, int D, float* __restrict__ out) {
  const int b = blockIdx.y;
  const int d = blockIdx.x * blockDim.x + threadIdx.x;
  if (d >= D) return;
  const T* x = X + (int64_t)b * R * D + d;
  float acc = 0.f;
  for (int r = 0; r < R; ++r) acc += simt::to_f(x[(int64_t)r * D]);
  out[(int64_t)b * D + d] = acc / (float)R;
}

// ---------------------------------------------------------------------------------------------- local loss
// One block per sample.  Z [T, ldz] fp32 = scale * <a_m, b_n> (already scaled).  mask [T].
//   loss_sum[b] = sum_m mask_m (lse_row_m - Z_mm) + sum_n mask_n (lse_col_n - Z_nn)       (columns / rows masked by -1e8)
//   dZ[m,n] = wgt * [ mask_m (softmax_row(m)[n] - d_mn) + mask_n (softmax_col(n)[m] - d_mn) ]   (bf16, when wgt given)
__global__ void sparc_local_kernel(const float* __restrict__ Z, int T, int ldz, const float* __restrict__ mask,
                                   float* __restrict__ loss_sum, const float* __restrict__ wgt_ptr,
                                   __nv_bfloat16* __restrict__ dZ, int lddz) {
  extern __shared__ float sm[];
  float* z = sm;                       // [T][T+1]
  float* rl = z + T * (T + 1);         // row lse [T]
  float* cl = rl + T;                  // col lse [T]
  float* mk = cl + T;                  // mask [T]
  __shared__ float red[32];
  const int b = blockIdx.x;
  const float* zb = Z + (int64_t)b * T * ldz;
  for (int i = threadIdx.x; i < T; i += blockDim.x) mk[i] = mask[(int64_t)b * T + i];
  __syncthreads();
  for (int e = threadIdx.x; e < T * T; e += blockDim.x) {
    const int m = e / T, n = e % T;
    z[m * (T + 1) + n] = zb[(int64_t)m * ldz + n];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  // rows: logits + (1 - mask_n) * (-1e8)   (pacl.py:537,549)
  for (int m = warp; m < T; m += nw) {
    float mx = -INFINITY;
    for (int n = lane; n < T; n += 32) mx = fmaxf(mx, z[m * (T + 1) + n] + (1.f - mk[n]) * (-1e8f));
    mx = ptx::warp_max(mx);
    float s = 0.f;
    for (int n = lane; n < T; n += 32) s += expf(z[m * (T + 1) + n] + (1.f - mk[n]) * (-1e8f) - mx);
    s = ptx::warp_sum(s);
    if (lane == 0) rl[m] = mx + logf(s);
  }
  // columns: the transposed problem (a <-> b), mask on m
  for (int n = warp; n < T; n += nw) {
    float mx = -INFINITY;
    for (int m = lane; m < T; m += 32) mx = fmaxf(mx, z[m * (T + 1) + n] + (1.f - mk[m]) * (-1e8f));
    mx = ptx::warp_max(mx);
    float s = 0.f;
    for (int m = lane; m < T; m += 32) s += expf(z[m * (T + 1) + n] + (1.f - mk[m]) * (-1e8f) - mx);
    s = ptx::warp_sum(s);
    if (lane == 0) cl[n] = mx + logf(s);
  }
  __syncthreads();
  float acc = 0.f;
  for (int i = threadIdx.x; i < T; i += blockDim.x) {
    const float zi = z[i * (T + 1) + i] + (1.f - mk[i]) * (-1e8f);
    acc += mk[i] * ((rl[i] - zi) + (cl[i] - zi));
  }
  acc = simt::block_sum(acc, red);
  if (threadIdx.x == 0) loss_sum[b] = acc;
  if (dZ != nullptr) {
    const float wgt = *wgt_ptr;
    __nv_bfloat16* o = dZ + (int64_t)b * T * lddz;
    for (int e = threadIdx.x; e < T * lddz; e += blockDim.x) {
      const int m = e / lddz, n = e % lddz;
      float g = 0.f;
      if (n < T) {
        const float hit = (m == n) ? 1.f : 0.f;
        const float zr = z[m * (T + 1) + n] + (1.f - mk[n]) * (-1e8f);
        const float zc = z[m * (T + 1) + n] + (1.f - mk[m]) * (-1e8f);
        g = wgt * (mk[m] * (expf(zr - rl[m]) - hit) + mk[n] * (expf(zc - cl[n]) - hit));
      }
      o[e] = __float2bfloat16(g);
    }
  }
}

// ---------------------------------------------------------------------------------------------- orchestration
struct SparcWs {
  float* S;              // [B][T][Ppad] fp32
  __nv_bfloat16* W;      // [B][T][Ppad]
  float* Graw;           // [B][T][D] fp32
  float4* stats;         // [B][T]
  int2* arg;             // [B][T]
  float* dW;             // [B][T][Ppad] fp32        (bwd)
  __nv_bfloat16* dS;     // [B][T][Ppad]             (bwd)
  __nv_bfloat16* dG;     // [B][T][D]                (bwd)
};
static size_t sparc_carve(SparcWs* w, void* base, int B, int T, int P, int D, int backward) {
  const int Ppad = rup(P, 64);
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  const size_t bt = (size_t)B * T;
  w->S = static_cast<float*>(take(bt * Ppad * 4));
  w->W = static_cast<__nv_bfloat16*>(take(bt * Ppad * 2));
  w->Graw = static_cast<float*>(take(bt * D * 4));
  w->stats = static_cast<float4*>(take(bt * sizeof(float4)));
  w->arg = static_cast<int2*>(take(bt * sizeof(int2)));
  if (backward) {
    w->dW = static_cast<float*>(take(bt * Ppad * 4));
    w->dS = static_cast<__nv_bfloat16*>(take(bt * Ppad * 2));
    w->dG = static_cast<__nv_bfloat16*>(take(bt * D * 2));
  }
  return off;
}

template <class Epi, bool A_MN, bool B_MN>
static int launch_bn(int n, const OperandDesc* a, const OperandDesc* b, int npairs, const int* ks, int M, int N,
                     int batches, const typename Epi::Params& ep, cudaStream_t st) {
  if (n > 192) return launch_gemm<256, A_MN, B_MN, Epi>(a, b, npairs, ks, ks, M, N, batches, ep, st);
  if (n > 128) return launch_gemm<192, A_MN, B_MN, Epi>(a, b, npairs, ks, ks, M, N, batches, ep, st);
  if (n > 64) return launch_gemm<128, A_MN, B_MN, Epi>(a, b, npairs, ks, ks, M, N, batches, ep, st);
  return launch_gemm<64, A_MN, B_MN, Epi>(a, b, npairs, ks, ks, M, N, batches, ep, st);
}

// S = L V^T, row statistics, W
static int sparc_scores_and_weights(const __nv_bfloat16* V, const __nv_bfloat16* L, int B, int T, int P, int D,
                                    float sigma, const SparcWs& w, cudaStream_t st) {
  const int Ppad = rup(P, 64);
  OperandDesc a, b;
  a.ptr = L; a.rows = T; a.k = D; a.ld = D; a.batch = B; a.batch_stride = (int64_t)T * D; a.bmul = 1;
  b.ptr = V; b.rows = P; b.k = D; b.ld = D; b.batch = B; b.batch_stride = (int64_t)P * D; b.bmul = 1;
  const int ks[1] = {(D + 63) / 64};
  epi::Store<false>::Params ep{w.S, Ppad, (int64_t)T * Ppad, T, P, 1.f, 0};
  const int bn = Ppad % 256 == 0 ? 256 : (Ppad % 192 == 0 ? 192 : (Ppad % 128 == 0 ? 128 : 64));
  CLIPK_TRY(launch_bn<epi::Store<false>, false, false>(bn, &a, &b, 1, ks, T, P, B, ep, st));
  const int64_t rows = (int64_t)B * T;
  sparc_rows_fwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.S, rows, P, Ppad, sigma, w.W, w.stats, w.arg);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// The fused per-sample forward (sparc_fused.cuh) covers T <= 79 tokens, P <= 640 patches, D <= 768 in multiples of 128;
// other shapes (and CLIPK_SPARC_FUSED=0) take the staged path below.
static bool sparc_fused_eligible(int T, int P, int D) {
  static const bool on = [] {
    const char* e = getenv("CLIPK_SPARC_FUSED");
    return e == nullptr || atoi(e) != 0;
  }();
  return on && T < sfz::kTn && P <= 128 * sfz::kMaxMT && D <= 64 * sfz::kMaxKsD && D % 128 == 0;
}

static int sparc_fused_fwd(const __nv_bfloat16* V, const __nv_bfloat16* L, int B, int T, int P, int D, float sigma,
                           float* l_hat, float* g_hat, float* lnorm, float* gnorm, float4* stats, int2* arg, float* pooled,
                           cudaStream_t st) {
  sfz::Params pr;
  memset(&pr, 0, sizeof(pr));
  pr.B = B; pr.T = T; pr.P = P; pr.D = D;
  pr.nMT = (P + 127) / 128; pr.ksD = (D + 63) / 64; pr.kbP = (P + 63) / 64; pr.nU = D / 128;
  pr.sigma = sigma;
  pr.L = L; pr.l_hat = l_hat; pr.g_hat = g_hat; pr.lnorm = lnorm; pr.gnorm = gnorm; pr.stats = stats; pr.arg = arg;
  pr.pooled = pooled;
  sfz::Maps mp;
  memset(&mp, 0, sizeof(mp));
  const uint64_t ldD = (uint64_t)D * 2;
  CLIPK_TRY(make_tmap_bf16(&mp.L, L, D, T, B, ldD, (uint64_t)T * ldD, sfz::kTn));
  CLIPK_TRY(make_tmap_bf16(&mp.Vk, V, D, P, B, ldD, (uint64_t)P * ldD, 128));
  CLIPK_TRY(make_tmap_bf16(&mp.Vmn, V, D, P, B, ldD, (uint64_t)P * ldD, 64));
  static std::atomic<uint64_t> attr_mask{0};
  int dev = 0;
  CLIPK_CHECK_CUDA(cudaGetDevice(&dev));
  if (!(attr_mask.load(std::memory_order_acquire) & (1ull << (dev & 63)))) {
    CLIPK_CHECK_CUDA(cudaFuncSetAttribute(sfz::sparc_fused_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, sfz::kSmemTotal));
    attr_mask.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  const int grid = B < sm_count() ? B : sm_count();
  const bool tr = trace_enabled();
  if (tr) trace_begin("sparc_fused_fwd_kernel", st);
  sfz::sparc_fused_fwd_kernel<<<grid, sfz::kThreads, sfz::kSmemTotal, st>>>(mp, pr);
  if (tr) trace_end(st);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int sparc_align_fwd(const __nv_bfloat16* V, const __nv_bfloat16* L, int B, int T, int P, int D, float sigma,
                    float* l_hat, float* g_hat, float* lnorm, float* gnorm, float* pooled, void* ws, size_t ws_bytes,
                    cudaStream_t st) {
  CLIPK_REQUIRE(B > 0 && T > 0 && P > 0 && D > 0 && D % 8 == 0, "sparc_align_fwd: bad shape B=%d T=%d P=%d D=%d", B, T, P, D);
  if (sparc_fused_eligible(T, P, D)) {
    SparcWs wf{};
    const size_t needf = sparc_carve(&wf, ws, B, T, P, D, 0);
    CLIPK_REQUIRE(ws != nullptr && ws_bytes >= needf, "sparc_align_fwd: workspace too small (%zu < %zu)", ws_bytes, needf);
    return sparc_fused_fwd(V, L, B, T, P, D, sigma, l_hat, g_hat, lnorm, gnorm, wf.stats, wf.arg, pooled, st);
  }
  if (pooled != nullptr) {
    dim3 grid((D + 255) / 256, B);
    mean_dim1_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(V, P, D, pooled);
    clipk::count_launches(1);
  }
  CLIPK_REQUIRE(T <= 128, "sparc_align_fwd: T=%d > 128 tokens is not supported", T);
  SparcWs w{};
  const size_t need = sparc_carve(&w, ws, B, T, P, D, 0);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "sparc_align_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ppad = rup(P, 64);
  CLIPK_TRY(sparc_scores_and_weights(V, L, B, T, P, D, sigma, w, st));
  // G = W V
  {
    OperandDesc a, b;
    a.ptr = w.W; a.rows = T; a.k = Ppad; a.ld = Ppad; a.batch = B; a.batch_stride = (int64_t)T * Ppad; a.bmul = 1;
    b.ptr = V; b.mn_major = true; b.rows = D; b.k = P; b.ld = D; b.batch = B; b.batch_stride = (int64_t)P * D; b.bmul = 1;
    const int ks[1] = {Ppad / 64};
    epi::Store<false>::Params ep{w.Graw, D, (int64_t)T * D, T, D, 1.f, 0};
    CLIPK_TRY(launch_bn<epi::Store<false>, false, true>(D, &a, &b, 1, ks, T, D, B, ep, st));
  }
  const int64_t rows = (int64_t)B * T;
  normalize_rows_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.Graw, rows, D, g_hat, nullptr, gnorm);
  clipk::count_launches(1);
  normalize_rows_kernel<__nv_bfloat16><<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(L, rows, D, l_hat, nullptr, lnorm);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// d_g_hat, d_l_hat [B,T,D] fp32; g_add [B,D] fp32 (nullable: broadcast gradient of mean_p V); outputs dV (bf16/fp32), dL fp32
int sparc_align_bwd(const __nv_bfloat16* V, const __nv_bfloat16* L, int B, int T, int P, int D, float sigma,
                    const float* l_hat, const float* g_hat, const float* lnorm, const float* gnorm,
                    const float* d_g_hat, const float* d_l_hat, const float* g_add, void* dV, int dv_bf16, float* dL,
                    void* ws, size_t ws_bytes, cudaStream_t st) {
  CLIPK_REQUIRE(B > 0 && T > 0 && P > 0 && D > 0 && D % 8 == 0, "sparc_align_bwd: bad shape B=%d T=%d P=%d D=%d", B, T, P, D);
  CLIPK_REQUIRE(T <= 128, "sparc_align_bwd: T=%d > 128 tokens is not supported", T);
  SparcWs w{};
  const size_t need = sparc_carve(&w, ws, B, T, P, D, 1);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "sparc_align_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int Ppad = rup(P, 64);
  const int64_t rows = (int64_t)B * T;
  CLIPK_TRY(sparc_scores_and_weights(V, L, B, T, P, D, sigma, w, st));       // recompute S, W, stats
  // dG (raw) = Jacobian of n(G) ; dL_direct = Jacobian of n(L)
  normalize_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(g_hat, d_g_hat, gnorm, rows, D, nullptr, 0, w.dG);
  clipk::count_launches(1);
  normalize_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(l_hat, d_l_hat, lnorm, rows, D, dL, 0, nullptr);
  clipk::count_launches(1);
  // dW = dG V^T
  {
    OperandDesc a, b;
    a.ptr = w.dG; a.rows = T; a.k = D; a.ld = D; a.batch = B; a.batch_stride = (int64_t)T * D; a.bmul = 1;
    b.ptr = V; b.rows = P; b.k = D; b.ld = D; b.batch = B; b.batch_stride = (int64_t)P * D; b.bmul = 1;
    const int ks[1] = {(D + 63) / 64};
    epi::Store<false>::Params ep{w.dW, Ppad, (int64_t)T * Ppad, T, P, 1.f, 0};
    const int bn = Ppad % 256 == 0 ? 256 : (Ppad % 192 == 0 ? 192 : (Ppad % 128 == 0 ? 128 : 64));
    CLIPK_TRY(launch_bn<epi::Store<false>, false, false>(bn, &a, &b, 1, ks, T, P, B, ep, st));
  }
  sparc_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, st>>>(w.S, w.dW, rows, P, Ppad, sigma, w.stats, w.arg, w.dS);
  clipk::count_launches(1);
  // dL += dS V
  {
    OperandDesc a, b;
    a.ptr = w.dS; a.rows = T; a.k = Ppad; a.ld = Ppad; a.batch = B; a.batch_stride = (int64_t)T * Ppad; a.bmul = 1;
    b.ptr = V; b.mn_major = true; b.rows = D; b.k = P; b.ld = D; b.batch = B; b.batch_stride = (int64_t)P * D; b.bmul = 1;
    const int ks[1] = {Ppad / 64};
    epi::Store<false>::Params ep{dL, D, (int64_t)T * D, T, D, 1.f, 1};
    CLIPK_TRY(launch_bn<epi::Store<false>, false, true>(D, &a, &b, 1, ks, T, D, B, ep, st));
  }
  // dV = W^T dG + dS^T L (+ g_add broadcast)
  {
    OperandDesc a[2], b[2];
    a[0].ptr = w.W; a[0].mn_major = true; a[0].rows = Ppad; a[0].k = T; a[0].ld = Ppad; a[0].batch = B;
    a[0].batch_stride = (int64_t)T * Ppad; a[0].bmul = 1;
    b[0].ptr = w.dG; b[0].mn_major = true; b[0].rows = D; b[0].k = T; b[0].ld = D; b[0].batch = B;
    b[0].batch_stride = (int64_t)T * D; b[0].bmul = 1;
    a[1] = a[0]; a[1].ptr = w.dS;
    b[1] = b[0]; b[1].ptr = L;
    const int ks[2] = {(T + 63) / 64, (T + 63) / 64};
    if (dv_bf16 && D % 8 == 0 && (reinterpret_cast<uintptr_t>(dV) & 15) == 0 && P > eng::BM) {
      // bf16 dV: CTA-pair engine, every epilogue warp writes its [32 x 32] chunk with its own TMA store
      epi::SparcDvTma::Params ep{{dV, D, (int64_t)P * D, P, D, B}, g_add, D};
      if (D > 128) CLIPK_TRY((launch_gemm2<256, true, true, epi::SparcDvTma>(a, b, 2, ks, ks, P, D, B, ep, st)));
      else CLIPK_TRY((launch_gemm2<128, true, true, epi::SparcDvTma>(a, b, 2, ks, ks, P, D, B, ep, st)));
    } else {
      epi::SparcDv::Params ep{g_add, dV, P, D, dv_bf16};
      CLIPK_TRY(launch_bn<epi::SparcDv, true, true>(D, a, b, 2, ks, P, D, B, ep, st));
    }
  }
  return 0;
}

}  // namespace clipk

extern "C" {

size_t clipk_sparc_workspace_bytes(int B, int T, int P, int D, int backward) {
  clipk::SparcWs w{};
  return clipk::sparc_carve(&w, nullptr, B, T, P, D, backward);
}

int clipk_sparc_align_fwd(const void* V, const void* L, int B, int T, int P, int D, float sigma, float* l_hat,
                          float* g_hat, float* lnorm, float* gnorm, float* pooled, void* workspace, size_t ws_bytes,
                          void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::sparc_align_fwd(static_cast<const __nv_bfloat16*>(V), static_cast<const __nv_bfloat16*>(L), B, T, P, D,
                                sigma, l_hat, g_hat, lnorm, gnorm, pooled, workspace, ws_bytes,
                                static_cast<cudaStream_t>(stream));
}

int clipk_sparc_align_bwd(const void* V, const void* L, int B, int T, int P, int D, float sigma, const float* l_hat,
                          const float* g_hat, const float* lnorm, const float* gnorm, const float* d_g_hat,
                          const float* d_l_hat, const float* g_add, void* dV, int dv_bf16, float* dL, void* workspace,
                          size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::sparc_align_bwd(static_cast<const __nv_bfloat16*>(V), static_cast<const __nv_bfloat16*>(L), B, T, P, D,
                                sigma, l_hat, g_hat, lnorm, gnorm, d_g_hat, d_l_hat, g_add, dV, dv_bf16, dL, workspace,
                                ws_bytes, static_cast<cudaStream_t>(stream));
}

// mean over dim 1: X [B,R,D] (bf16 | fp32) -> out [B,D] fp32          (torch.mean(..., dim=1), pacl.py:561-562)
int clipk_mean_dim1(const void* X, int dtype, int B, int R, int D, float* out, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(B >= 0 && R > 0 && D > 0, "mean_dim1: bad shape");
  if (B == 0) return 0;
  dim3 grid((D + 255) / 256, B);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dtype == CLIPK_BF16) {
    clipk::mean_dim1_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(X), R, D, out);
    clipk::count_launches(1);
  } else if (dtype == CLIPK_F32) {
    clipk::mean_dim1_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(X), R, D, out);
    clipk::count_launches(1);
  } else if (dtype == CLIPK_F16) {
    clipk::mean_dim1_kernel<__half><<<grid, 256, 0, st>>>(static_cast<const __half*>(X), R, D, out);
    clipk::count_launches(1);
  }
  else { clipk::set_error("mean_dim1: bad dtype %d", dtype); return CLIPK_ERR_INVALID; }
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// x^ = F.normalize(x, dim=-1) for fp32 rows [rows, D]; norm [rows] saved for the backward
int clipk_normalize_rows_fwd(const float* X, int64_t rows, int D, float* out, float* norm, void* stream) {
  CLIPK_TRY(clipk::check_device());
  if (rows == 0) return 0;
  clipk::normalize_rows_kernel<float><<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      X, rows, D, out, nullptr, norm);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}
int clipk_normalize_rows_bwd(const float* xh, const float* g, const float* norm, int64_t rows, int D, float* dx,
                             void* stream) {
  CLIPK_TRY(clipk::check_device());
  if (rows == 0) return 0;
  clipk::normalize_rows_bwd_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      xh, g, norm, rows, D, dx, 0, nullptr);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// SparcLoss local term (pacl.py:522-556), both directions at once: a = g_hat, b = l_hat (fp32 [B,T,D], normalised).
//   Z_b = scale * a_b b_b^T (tcgen05, bf16 operands).  Direction 1 is the row-wise CE of Z (columns masked),
//   direction 2 the column-wise CE (rows masked); both share the same logits.
//   fwd: loss_sum [B] = sum_m mask_m CE_row_m + sum_n mask_n CE_col_n    (caller divides by 2 * sum(mask))
//   bwd: wgt = device scalar (upstream * 0.5 / sum(mask)) -> d_a, d_b fp32 [B,T,D]
size_t clipk_sparc_local_workspace_bytes(int B, int T, int D) {
  const size_t Tp = (size_t)clipk::rup(T, 8);
  auto r = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  return 2 * r((size_t)B * T * D * 2) + r((size_t)B * T * Tp * 4) + r((size_t)B * T * Tp * 2);
}
}  // extern "C"

namespace clipk {

__global__ void cast_bf16_kernel(const float* __restrict__ x, int64_t n, __nv_bfloat16* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = __float2bfloat16(x[i]);
}
// 8 values per thread: two 16-byte loads, one 16-byte store (n8 = n / 8; both pointers 16-byte aligned)
__global__ void cast_bf16x8_kernel(const float4* __restrict__ x, int64_t n8, uint4* __restrict__ y) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  const float4 a = __ldg(x + 2 * i), b = __ldg(x + 2 * i + 1);
  y[i] = make_uint4(ptx::pack_bf16x2(a.x, a.y), ptx::pack_bf16x2(a.z, a.w), ptx::pack_bf16x2(b.x, b.y),
                    ptx::pack_bf16x2(b.z, b.w));
}
static void launch_cast_bf16(const float* x, int64_t n, __nv_bfloat16* y, cudaStream_t st) {
  if (n % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0) {
    const int64_t n8 = n / 8;
    cast_bf16x8_kernel<<<(unsigned)((n8 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(x), n8,
                                                                     reinterpret_cast<uint4*>(y));
  } else {
    cast_bf16_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, n, y);
  }
  clipk::count_launches(1);
}

struct LocalWs {
  __nv_bfloat16 *a16, *b16, *dZ;
  float* Z;
};
static void local_carve(LocalWs* w, void* ws, int B, int T, int D) {
  const size_t Tp = (size_t)rup(T, 8);
  auto r = [](size_t x) { return (x + 1023) / 1024 * 1024; };
  char* base = static_cast<char*>(ws);
  w->a16 = reinterpret_cast<__nv_bfloat16*>(base);
  base += r((size_t)B * T * D * 2);
  w->b16 = reinterpret_cast<__nv_bfloat16*>(base);
  base += r((size_t)B * T * D * 2);
  w->Z = reinterpret_cast<float*>(base);
  base += r((size_t)B * T * Tp * 4);
  w->dZ = reinterpret_cast<__nv_bfloat16*>(base);
}

static int sparc_local_logits(const float* a, const float* b, int B, int T, int D, float scale, const LocalWs& w,
                              cudaStream_t st) {
  const int Tp = rup(T, 8);
  const int64_t n = (int64_t)B * T * D;
  launch_cast_bf16(a, n, w.a16, st);
  launch_cast_bf16(b, n, w.b16, st);
  OperandDesc oa, ob;
  oa.ptr = w.a16; oa.rows = T; oa.k = D; oa.ld = D; oa.batch = B; oa.batch_stride = (int64_t)T * D; oa.bmul = 1;
  ob.ptr = w.b16; ob.rows = T; ob.k = D; ob.ld = D; ob.batch = B; ob.batch_stride = (int64_t)T * D; ob.bmul = 1;
  const int ks[1] = {(D + 63) / 64};
  epi::Store<false>::Params ep{w.Z, Tp, (int64_t)T * Tp, T, T, scale, 0};
  return launch_bn<epi::Store<false>, false, false>(T, &oa, &ob, 1, ks, T, T, B, ep, st);
}

}  // namespace clipk

extern "C" {

int clipk_sparc_local_fwd(const float* a, const float* b, const float* mask, int B, int T, int D, float scale,
                          float* loss_sum, void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(B > 0 && T > 0 && T <= 128 && D > 0 && D % 8 == 0, "sparc_local_fwd: bad shape B=%d T=%d D=%d", B, T, D);
  CLIPK_REQUIRE(workspace != nullptr && ws_bytes >= clipk_sparc_local_workspace_bytes(B, T, D),
                "sparc_local_fwd: workspace too small");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  clipk::LocalWs w{};
  clipk::local_carve(&w, workspace, B, T, D);
  CLIPK_TRY(clipk::sparc_local_logits(a, b, B, T, D, scale, w, st));
  const int Tp = clipk::rup(T, 8);
  const size_t smem = (size_t)(T * (T + 1) + 3 * T) * sizeof(float);
  CLIPK_CHECK_CUDA(cudaFuncSetAttribute(clipk::sparc_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  clipk::sparc_local_kernel<<<B, 256, smem, st>>>(w.Z, T, Tp, mask, loss_sum, nullptr, nullptr, Tp);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_sparc_local_bwd(const float* a, const float* b, const float* mask, int B, int T, int D, float scale,
                          const float* wgt, float* d_a, float* d_b, float* loss_sum_scratch, void* workspace,
                          size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(B > 0 && T > 0 && T <= 128 && D > 0 && D % 8 == 0, "sparc_local_bwd: bad shape B=%d T=%d D=%d", B, T, D);
  CLIPK_REQUIRE(workspace != nullptr && ws_bytes >= clipk_sparc_local_workspace_bytes(B, T, D),
                "sparc_local_bwd: workspace too small");
  using namespace clipk;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  LocalWs w{};
  local_carve(&w, workspace, B, T, D);
  CLIPK_TRY(sparc_local_logits(a, b, B, T, D, scale, w, st));
  const int Tp = rup(T, 8);
  const size_t smem = (size_t)(T * (T + 1) + 3 * T) * sizeof(float);
  CLIPK_CHECK_CUDA(cudaFuncSetAttribute(sparc_local_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  sparc_local_kernel<<<B, 256, smem, st>>>(w.Z, T, Tp, mask, loss_sum_scratch, wgt, w.dZ, Tp);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  const int ks[1] = {(T + 63) / 64};
  {  // d_a = scale * dZ b
    OperandDesc oa, ob;
    oa.ptr = w.dZ; oa.rows = T; oa.k = T; oa.ld = Tp; oa.batch = B; oa.batch_stride = (int64_t)T * Tp; oa.bmul = 1;
    ob.ptr = w.b16; ob.mn_major = true; ob.rows = D; ob.k = T; ob.ld = D; ob.batch = B; ob.batch_stride = (int64_t)T * D; ob.bmul = 1;
    epi::Store<false>::Params ep{d_a, D, (int64_t)T * D, T, D, scale, 0};
    CLIPK_TRY(launch_bn<epi::Store<false>, false, true>(D, &oa, &ob, 1, ks, T, D, B, ep, st));
  }
  {  // d_b = scale * dZ^T a
    OperandDesc oa, ob;
    oa.ptr = w.dZ; oa.mn_major = true; oa.rows = T; oa.k = T; oa.ld = Tp; oa.batch = B; oa.batch_stride = (int64_t)T * Tp; oa.bmul = 1;
    ob.ptr = w.a16; ob.mn_major = true; ob.rows = D; ob.k = T; ob.ld = D; ob.batch = B; ob.batch_stride = (int64_t)T * D; ob.bmul = 1;
    epi::Store<false>::Params ep{d_b, D, (int64_t)T * D, T, D, scale, 0};
    CLIPK_TRY(launch_bn<epi::Store<false>, true, true>(D, &oa, &ob, 1, ks, T, D, B, ep, st));
  }
  return 0;
}
}
