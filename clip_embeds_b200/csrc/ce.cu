// Cross-entropy / InfoNCE kernels.
//
//  (1) CE over a materialised fp32 logit / score matrix L [M, N] (all-pairs PACL scores, fp32 feature path):
//      row LSE + per-row loss, column (max, sumexp) partials for cross-rank merging, and the gradient of
//      w_row * CE_rows + w_col * CE_cols.   Replaces F.cross_entropy at PACL/model/pacl.py:509-512.
//  (2) fp32 SIMT GEMM with generic strides: the fp32 feature path (reference training dtype, pacl.py:498-501 and
//      open_clip/src/open_clip/loss.py:156-164 when features are fp32).
//  (3) bf16 feature CE on the tcgen05 engine: logits = scale * X Y^T + bias are produced tile by tile in TMEM and
//      reduced by the epilogue (online log-sum-exp partials) without ever being written; backward re-computes the
//      tile, emits dL = w (softmax - onehot) as bf16 and feeds it to two engine GEMMs (dX = dL Y, dY = dL^T X).
//      Handles ignore_index rows (labels < 0, loss.py:131-134) and label offsets (loss.py:128-130).
#include "common.cuh"
#include "epilogues.cuh"
#include "simt_util.cuh"

namespace clipk {

// ------------------------------------------------------------------------------------------------ (1) matrix CE
__global__ void ce_rows_kernel(const float* __restrict__ L, int M, int N, int64_t ld, const int64_t* __restrict__ labels,
                               int64_t label_offset, float* __restrict__ row_lse, float* __restrict__ row_loss) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= M) return;
  const int lane = threadIdx.x & 31;
  const float* l = L + (int64_t)row * ld;
  float mx = -INFINITY;
  for (int n = lane; n < N; n += 32) mx = fmaxf(mx, l[n]);
  mx = ptx::warp_max(mx);
  float sm = 0.f;
  for (int n = lane; n < N; n += 32) sm += expf(l[n] - mx);
  sm = ptx::warp_sum(sm);
  if (lane == 0) {
    const float lse = mx + logf(sm);
    row_lse[row] = lse;
    if (row_loss != nullptr) {
      const int64_t lab = labels != nullptr ? labels[row] : (int64_t)row + label_offset;
      row_loss[row] = (lab >= 0 && lab < N) ? lse - l[lab] : 0.f;
    }
  }
}

// column-wise (max, sumexp) over the M local rows.  block = 32 columns x 32 row-lanes; two passes (max, then
// sum exp(v - max)) so that every load / exp of a thread is independent of the previous one (the one-pass online form is
// a serial dependency chain of M / lanes exps per thread).
__global__ void __launch_bounds__(1024) ce_cols_kernel(const float* __restrict__ L, int M, int N, int64_t ld,
                                                        float* __restrict__ col_max, float* __restrict__ col_sum) {
  __shared__ float red[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int n = blockIdx.x * 32 + tx;
  float mx = -INFINITY;
  if (n < N)
    for (int m = ty; m < M; m += 32) mx = fmaxf(mx, L[(int64_t)m * ld + n]);
  red[ty][tx] = mx;
  __syncthreads();
  float gm = -INFINITY;
#pragma unroll
  for (int j = 0; j < 32; ++j) gm = fmaxf(gm, red[j][tx]);
  __syncthreads();
  float sm = 0.f;
  if (n < N && gm > -INFINITY)
    for (int m = ty; m < M; m += 32) sm += expf(L[(int64_t)m * ld + n] - gm);
  red[ty][tx] = sm;
  __syncthreads();
  if (ty == 0 && n < N) {
    float gs = 0.f;
#pragma unroll
    for (int j = 0; j < 32; ++j) gs += red[j][tx];
    col_max[n] = gm;
    col_sum[n] = gs;
  }
}

// dL[i,k] = w_row (exp(L - rlse_i) - [k == lab_i]) + w_col (exp(L - clse_k) - [k == lab_i]),  lab_i = i + offset
__global__ void ce_scores_grad_kernel(const float* __restrict__ L, int M, int N, int64_t ld,
                                      const float* __restrict__ row_lse, const float* __restrict__ col_lse,
                                      int64_t label_offset, float w_row, float w_col, float* __restrict__ dL) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int i = (int)(idx / N), k = (int)(idx % N);
  const float v = L[(int64_t)i * ld + k];
  const float hit = ((int64_t)k == (int64_t)i + label_offset) ? 1.f : 0.f;
  float g = 0.f;
  if (w_row != 0.f) g += w_row * (expf(v - row_lse[i]) - hit);
  if (w_col != 0.f) g += w_col * (expf(v - col_lse[k]) - hit);
  dL[idx] = g;
}

// dL[i,k] = w[i] (exp(L - rlse_i) - [k == lab_i]) with explicit labels (ignore rows: w[i] = 0)
__global__ void ce_rows_grad_kernel(const float* __restrict__ L, int M, int N, int64_t ld,
                                    const float* __restrict__ row_lse, const int64_t* __restrict__ labels,
                                    int64_t label_offset, const float* __restrict__ w, float* __restrict__ dL) {
  const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (int64_t)M * N) return;
  const int i = (int)(idx / N), k = (int)(idx % N);
  const int64_t lab = labels != nullptr ? labels[i] : (int64_t)i + label_offset;
  const float wi = w[i];
  dL[idx] = wi == 0.f ? 0.f : wi * (expf(L[(int64_t)i * ld + k] - row_lse[i]) - ((int64_t)k == lab ? 1.f : 0.f));
}

// out2[0] = sum_i row_loss_i,  out2[1] = sum_i (row_lse_i - row_loss_i) = sum of the label logits (the local diagonal)
__global__ void ce_rowsums_kernel(const float* __restrict__ row_lse, const float* __restrict__ row_loss, int M,
                                  float* __restrict__ out2) {
  __shared__ float red[32];
  float a = 0.f, d = 0.f;
  for (int i = threadIdx.x; i < M; i += blockDim.x) {
    a += row_loss[i];
    d += row_lse[i] - row_loss[i];
  }
  a = simt::block_sum(a, red);
  d = simt::block_sum(d, red);
  if (threadIdx.x == 0) {
    out2[0] = a;
    out2[1] = d;
  }
}

// Merge of the per-rank payloads [W][2N + 2] = (col_max [N], col_sum [N], sum row_loss, sum diag):
//   col_lse[k] = log sum_r col_sum[r][k] exp(col_max[r][k] - max_r)  + max_r        (online-softmax merge)
//   loss = 0.5 / N * ( sum_r rowloss_r + sum_k col_lse[k] - sum_r diag_r )           (pacl.py:509-512 on the scores)
__global__ void ce_merge_kernel(const float* __restrict__ g, int W, int N, float* __restrict__ col_lse,
                                float* __restrict__ loss) {
  __shared__ float red[32];
  const int64_t stride = 2 * (int64_t)N + 2;
  float acc = 0.f;
  for (int k = threadIdx.x; k < N; k += blockDim.x) {
    float mx = -INFINITY;
    for (int r = 0; r < W; ++r) mx = fmaxf(mx, g[r * stride + k]);
    float sm = 0.f;
    for (int r = 0; r < W; ++r) {
      const float m = g[r * stride + k];
      if (m > -INFINITY) sm += g[r * stride + N + k] * expf(m - mx);
    }
    const float l = mx + logf(sm);
    col_lse[k] = l;
    acc += l;
  }
  acc = simt::block_sum(acc, red);
  if (threadIdx.x == 0) {
    float rl = 0.f, dg = 0.f;
    for (int r = 0; r < W; ++r) {
      rl += g[r * stride + 2 * (int64_t)N];
      dg += g[r * stride + 2 * (int64_t)N + 1];
    }
    *loss = 0.5f / (float)N * (rl + acc - dg);
  }
}

// ------------------------------------------------------------------------------------------------ (2) fp32 GEMM
// C[m,n] = alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] + beta * C[m*ldc + n]; 64x64 tile, 4x4 per thread.
__global__ void __launch_bounds__(256)
sgemm_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
                     int64_t sbn, float* __restrict__ C, int64_t ldc, int M, int N, int K, float alpha, float beta) {
  __shared__ float As[16][65], Bs[16][65];
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int m0 = blockIdx.y * 64, n0 = blockIdx.x * 64;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += 16) {
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int r, kk;
      if (sak == 1) { kk = e & 15; r = e >> 4; } else { r = e & 63; kk = e >> 6; }
      const int m = m0 + r, k = k0 + kk;
      As[kk][r] = (m < M && k < K) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.f;
    }
    for (int e = threadIdx.x; e < 64 * 16; e += 256) {
      int c, kk;
      if (sbk == 1) { kk = e & 15; c = e >> 4; } else { c = e & 63; kk = e >> 6; }
      const int n = n0 + c, k = k0 + kk;
      Bs[kk][c] = (n < N && k < K) ? B[(int64_t)k * sbk + (int64_t)n * sbn] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < 16; ++kk) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* c = C + (int64_t)m * ldc + n;
      *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}

// Larger tiling of the same GEMM for the shapes the fp32 feature path actually runs (B x B x D logits and their
// gradients): 128 x 128 tile, 8 x 8 outputs per thread (two 4-wide groups 64 apart in each direction, so shared-memory
// reads are conflict-free LDS.128), BK = 8, global loads of the next k-slab prefetched into registers while the
// current one is multiplied.  Same fp32 FMA arithmetic as above (only the order of the K loop blocks differs).
__global__ void __launch_bounds__(256)
sgemm128_strided_kernel(const float* __restrict__ A, int64_t sam, int64_t sak, const float* __restrict__ B, int64_t sbk,
                        int64_t sbn, float* __restrict__ C, int64_t ldc, int M, int N, int K, float alpha, float beta) {
  constexpr int BMN = 128, BKK = 8;
  __shared__ __align__(16) float As[2][BKK][BMN + 4];     // +4: the k-fastest stores of a K-contiguous operand hit 32 banks
  __shared__ __align__(16) float Bs[2][BKK][BMN + 4];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;              // 16 x 16 threads
  const int m0 = blockIdx.y * BMN, n0 = blockIdx.x * BMN;
  // each thread loads 4 elements of A and 4 of B per k-slab: element e = tid + 256 * i of the [BMN x BKK] slab
  float ra[4], rb[4];
  auto gload = [&](int k0) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int r, kk;
      if (sak == 1) { kk = e & 7; r = e >> 3; } else { r = e & 127; kk = e >> 7; }
      const int m = m0 + r, k = k0 + kk;
      ra[i] = (m < M && k < K) ? A[(int64_t)m * sam + (int64_t)k * sak] : 0.f;
      int c, kb;
      if (sbk == 1) { kb = e & 7; c = e >> 3; } else { c = e & 127; kb = e >> 7; }
      const int n = n0 + c, k2 = k0 + kb;
      rb[i] = (n < N && k2 < K) ? B[(int64_t)k2 * sbk + (int64_t)n * sbn] : 0.f;
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int e = tid + 256 * i;
      int r, kk;
      if (sak == 1) { kk = e & 7; r = e >> 3; } else { r = e & 127; kk = e >> 7; }
      As[buf][kk][r] = ra[i];
      int c, kb;
      if (sbk == 1) { kb = e & 7; c = e >> 3; } else { c = e & 127; kb = e >> 7; }
      Bs[buf][kb][c] = rb[i];
    }
  };
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  gload(0);
  sstore(0);
  __syncthreads();
  int buf = 0;
  for (int k0 = 0; k0 < K; k0 += BKK) {
    const bool more = k0 + BKK < K;
    if (more) gload(k0 + BKK);
#pragma unroll
    for (int kk = 0; kk < BKK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (more) {
      sstore(buf ^ 1);
      __syncthreads();
      buf ^= 1;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n >= N) continue;
      float* c = C + (int64_t)m * ldc + n;
      *c = alpha * acc[i][j] + (beta != 0.f ? beta * *c : 0.f);
    }
  }
}

}  // namespace clipk

// ------------------------------------------------------------------------------------------------ (3) engine CE
namespace epi {

// number of existing columns among [n, n + 32) (warp-uniform): the matrix edge N and, past slab_n0, the fill count of
// the hard-negative slab the chunk lies in (chunks do not straddle slabs: slab_n0 and slab_rows are multiples of 32)
__device__ __forceinline__ int chunk_valid_cols(int n, int N, const int* slab_counts, int slab_n0, int slab_rows) {
  int lim = N - n;
  lim = lim < 32 ? lim : 32;
  if (slab_counts != nullptr && n >= slab_n0) {
    const int off = n - slab_n0;
    const int r = off / slab_rows;
    const int left = __ldg(slab_counts + r) - (off - r * slab_rows);
    lim = left < lim ? left : lim;
  }
  return lim < 0 ? 0 : lim;
}

// per-(row, n-tile) online log-sum-exp partial of logits = scale * acc + bias; captures the label logit.
struct LsePart {
  struct Params {
    float2* part;            // [tiles_n][M] (max, sumexp); tiles_n counts HALF tiles (two epilogue warps per row)
    float* pos;              // [M] logit at the label column (written by the tile that owns it)
    const int64_t* labels;   // nullable
    int64_t label_offset;
    int M, N, tiles_n;
    float scale, bias;
    const float* scale_dev;  // nullable: logit scale read from device memory (open_clip passes logit_scale.exp() as a tensor)
    // Fixed-capacity hard-negative slabs (loss.py:67-87 without the host-side size exchange): columns >= slab_n0 are W
    // slabs of `slab_rows` rows each, of which only the first slab_counts[r] exist; the others are masked out
    // (logit = -inf).  slab_counts == nullptr: every column exists.
    const int* slab_counts;
    int slab_n0, slab_rows;
  };
  Params p;
  float mx, sm;
  int64_t lab;
  __device__ explicit LsePart(const Params& pp) : p(pp), mx(0.f), sm(0.f), lab(-1) {
    if (p.scale_dev != nullptr) p.scale = __ldg(p.scale_dev);
  }
  __device__ void tile_begin(int, int m, int) {
    mx = -INFINITY;
    sm = 0.f;
    lab = -1;
    if (m < p.M) lab = p.labels != nullptr ? p.labels[m] : (int64_t)m + p.label_offset;
  }
  __device__ void chunk(int, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
    const int lim = chunk_valid_cols(n, p.N, p.slab_counts, p.slab_n0, p.slab_rows);
    if (lim <= 0) return;
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = fmaf(v[j], p.scale, p.bias);
      if (j < lim) cm = fmaxf(cm, v[j]);
    }
    const float nm = fmaxf(mx, cm);
    float acc = sm * __expf(mx - nm);
#pragma unroll
    for (int j = 0; j < 32; ++j)
      if (j < lim) acc += __expf(v[j] - nm);
    sm = acc;
    mx = nm;
    if (lab >= n && lab < (int64_t)n + 32 && lab < p.N) {
      float pv = 0.f;
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if ((int64_t)(n + j) == lab) pv = v[j];
      p.pos[m] = pv;
    }
  }
  __device__ void tile_end(int, int m, int, int tn, int half) {
    if (m < p.M) p.part[(int64_t)(2 * tn + half) * p.M + m] = make_float2(mx, sm);     // [tiles_n][M]: lanes = rows, coalesced
  }
};

// dL[m][n] = w[m] (exp(logit - lse[m]) - [n == label_m])  as bf16, leading dim ldd
struct DlOut {
  struct Params {
    const float* lse;        // [M]
    const float* w;          // [M]
    const int64_t* labels;   // nullable
    int64_t label_offset;
    __nv_bfloat16* dL;       // [M][ldd]
    int64_t ldd;
    int M, N;
    float scale, bias;
    const float* scale_dev;  // nullable
    const int* slab_counts;  // see LsePart
    int slab_n0, slab_rows;
  };
  Params p;
  float lse, w;
  int64_t lab;
  __device__ explicit DlOut(const Params& pp) : p(pp), lse(0.f), w(0.f), lab(-1) {
    if (p.scale_dev != nullptr) p.scale = __ldg(p.scale_dev);
  }
  __device__ void tile_begin(int, int m, int) {
    if (m < p.M) {
      lse = p.lse[m];
      w = p.w[m];
      lab = p.labels != nullptr ? p.labels[m] : (int64_t)m + p.label_offset;
    }
  }
  __device__ void chunk(int, int m, int n, float* v) {
    if (m >= p.M || n >= p.N) return;
    const int lim = chunk_valid_cols(n, p.N, p.slab_counts, p.slab_n0, p.slab_rows);
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float l = fmaf(v[j], p.scale, p.bias);
      const float hit = ((int64_t)(n + j) == lab) ? 1.f : 0.f;
      v[j] = (w == 0.f || j >= lim) ? 0.f : w * (__expf(l - lse) - hit);
    }
    store_bf16x32(p.dL + (int64_t)m * p.ldd + n, v, min(32, p.N - n));
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// The same through the engine's TMA-store path (CTA-pair engine: every epilogue warp stages its [32 x 32] bf16 chunk and
// writes it with its own TMA store; rows >= M and columns >= N are clipped by the output map's extents).
struct DlOutTma {
  static constexpr bool kTmaOut = true;
  struct Params {
    eng::OutDesc out;        // dL [M][ldd] bf16, extents (M, N)
    const float* lse;        // [M]
    const float* w;          // [M]
    const int64_t* labels;   // nullable
    int64_t label_offset;
    int M, N;
    float scale, bias;
    const float* scale_dev;  // nullable
    const int* slab_counts;  // see LsePart
    int slab_n0, slab_rows;
  };
  Params p;
  float nlse, w;
  int lab;                   // label column of this row, or -1
  __device__ explicit DlOutTma(const Params& pp) : p(pp), nlse(0.f), w(0.f), lab(-1) {
    if (p.scale_dev != nullptr) p.scale = __ldg(p.scale_dev);
  }
  __device__ void tile_begin(int, int m, int) {
    w = 0.f;
    nlse = 0.f;
    lab = -1;
    if (m < p.M) {
      nlse = fmaf(p.bias, 1.f, -p.lse[m]);          // bias - lse: logit - lse = acc * scale + (bias - lse)
      w = p.w[m];
      const int64_t l = p.labels != nullptr ? p.labels[m] : (int64_t)m + p.label_offset;
      lab = (l >= 0 && l < p.N) ? (int)l : -1;
    }
  }
  __device__ void chunk(int, int, int n, float* v) {
    // w (exp(logit - lse) - [n + j == label]);  rows with w == 0 (ignored rows, rows past M) give exact zeros
    const float k2 = p.scale * 1.4426950408889634f, c2 = nlse * 1.4426950408889634f;
    const int lim = p.slab_counts != nullptr ? chunk_valid_cols(n, p.N, p.slab_counts, p.slab_n0, p.slab_rows) : 32;
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = (w == 0.f || j >= lim) ? 0.f : w * ptx::ex2_approx(fmaf(v[j], k2, c2));
    const int rel = lab - n;
    if (rel >= 0 && rel < 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == rel) v[j] -= w;
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

// ---------------------------------------------------------------------------------------- symmetric CE, one GEMM
// Row AND column cross-entropy from ONE logits GEMM (open_clip computes `logits_per_text` with a second GEMM,
// loss.py:156-157): besides the row partials of LsePart, every chunk adds its column sums
//     col_sum[j] += sum_i exp(l_ij - ref),   ref = scale / 2 + bias      (columns j < ncol: the original captions)
// with a FIXED reference instead of a running column maximum, so that partial sums -- of other row tiles and of other
// ranks -- combine by plain addition (one atomicAdd per column and warp, one all-reduce across ranks).  Unit-norm
// features bound the logits to [bias - scale, bias + scale]: l - ref <= scale / 2 <= 50 (open_clip clamps logit_scale
// at 100) cannot overflow, and a column underflows only if its largest cosine is below 1/2 - 87 / scale (-0.37 at
// scale 100) -- the column's own positive pair rules that out; col_lse[j] = ref + log(col_sum[j]).
struct LseRowCol {
  struct Params {
    float2* part;            // [tiles_n][M] (max, sumexp) row partials, as LsePart
    float* pos;              // [M] logit at the label column
    int64_t label_offset;    // label of row m = m + label_offset
    int M, N, tiles_n;
    float scale, bias;
    const float* scale_dev;  // nullable
    const int* slab_counts;  // hard-negative slabs, see LsePart
    int slab_n0, slab_rows;
    float* col_sum;          // [ncol], zeroed by the caller
    int ncol;
  };
  Params p;
  float mx, sm, ref;
  int64_t lab;
  __device__ explicit LseRowCol(const Params& pp) : p(pp), mx(0.f), sm(0.f), lab(-1) {
    if (p.scale_dev != nullptr) p.scale = __ldg(p.scale_dev);
    ref = fmaf(0.5f, p.scale, p.bias);
  }
  __device__ void tile_begin(int, int m, int) {
    mx = -INFINITY;
    sm = 0.f;
    lab = m < p.M ? (int64_t)m + p.label_offset : -1;
  }
  __device__ void chunk(int, int m, int n, float* v) {
    if (n >= p.N) return;                                        // (warp-uniform)
    const int lim = chunk_valid_cols(n, p.N, p.slab_counts, p.slab_n0, p.slab_rows);
    if (lim <= 0) return;
    const bool rowok = m < p.M;
    float cm = -INFINITY;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      v[j] = fmaf(v[j], p.scale, p.bias);
      if (j < lim) cm = fmaxf(cm, v[j]);
    }
    if (rowok) {
      const float nm = fmaxf(mx, cm);
      float acc = sm * __expf(mx - nm);
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < lim) acc += __expf(v[j] - nm);
      sm = acc;
      mx = nm;
      if (lab >= n && lab < (int64_t)n + 32 && lab < p.N) {
        float pv = 0.f;
#pragma unroll
        for (int j = 0; j < 32; ++j)
          if ((int64_t)(n + j) == lab) pv = v[j];
        p.pos[m] = pv;
      }
    }
    if (n < p.ncol) {                                            // column sums over this warp's 32 rows
      float e[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) e[j] = rowok ? __expf(v[j] - ref) : 0.f;
      const float cs = ptx::warp_colsum32(e);                    // lane j: sum over the rows of column n + j
      const int col = n + (int)ptx::lane_id();
      if (col < p.ncol && cs != 0.f) atomicAdd(p.col_sum + col, cs);
    }
  }
  __device__ void tile_end(int, int m, int, int tn, int half) {
    if (m < p.M) p.part[(int64_t)(2 * tn + half) * p.M + m] = make_float2(mx, sm);
  }
};

// dL[m][n] = w_m (exp(l - rlse_m) - [n == lab_m]) + (exp(l + cb_n) - cw_lab [n == lab_m])   as bf16 through TMA stores;
// cb_n = log(col_w_n) - col_lse_n for the columns that carry a column CE (n < ncol), -inf elsewhere (Side: lane l holds
// cb of column n + l, loaded one chunk ahead).
struct DlSymTma {
  static constexpr bool kTmaOut = true;
  using Side = float;
  struct Params {
    eng::OutDesc out;        // dL [M][ldd] bf16, extents (M, N)
    const float* lse;        // [M] row lse
    const float* w;          // [M] row weights
    int64_t label_offset;
    int M, N;
    float scale, bias;
    const float* scale_dev;  // nullable
    const int* slab_counts;
    int slab_n0, slab_rows;
    const float* col_bias;   // [ncol]
    const float* col_w;      // [ncol]
    int ncol;
  };
  Params p;
  float nlse, w, wdiag;
  int lab;
  __device__ explicit DlSymTma(const Params& pp) : p(pp), nlse(0.f), w(0.f), wdiag(0.f), lab(-1) {
    if (p.scale_dev != nullptr) p.scale = __ldg(p.scale_dev);
  }
  __device__ void tile_begin(int, int m, int) {
    w = 0.f;
    nlse = 0.f;
    wdiag = 0.f;
    lab = -1;
    if (m < p.M) {
      nlse = p.bias - p.lse[m];
      w = p.w[m];
      const int64_t l = (int64_t)m + p.label_offset;
      lab = (l >= 0 && l < p.N) ? (int)l : -1;
      wdiag = w + ((lab >= 0 && lab < p.ncol) ? __ldg(p.col_w + lab) : 0.f);
    }
  }
  __device__ Side pre(int, int, int n) const {
    const int col = n + (int)ptx::lane_id();
    return col < p.ncol ? __ldg(p.col_bias + col) : -INFINITY;
  }
  __device__ void chunk(int, int m, int n, float* v, const Side& cb_l) {
    const float L2E = 1.4426950408889634f;
    const float k2 = p.scale * L2E, c2 = nlse * L2E, b2 = p.bias * L2E;
    const int lim = p.slab_counts != nullptr ? chunk_valid_cols(n, p.N, p.slab_counts, p.slab_n0, p.slab_rows) : 32;
    const bool rowok = m < p.M;
    const float cbl2 = cb_l * L2E + b2;                       // -inf stays -inf
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const float cj = __shfl_sync(0xffffffffu, cbl2, j);
      const float x = v[j] * k2;
      const float r = w * ptx::ex2_approx(x + c2);
      const float c = ptx::ex2_approx(x + cj);                // 0 where the column carries no CE
      v[j] = (rowok && j < lim) ? r + c : 0.f;
    }
    const int rel = lab - n;
    if (rel >= 0 && rel < 32) {
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j == rel) v[j] -= wdiag;
    }
  }
  __device__ void tile_end(int, int, int, int, int) {}
};

}  // namespace epi

namespace clipk {

// partials are laid out [tiles_n][M] (written coalesced by the epilogue warps, lanes = rows): one thread per row
__global__ void lse_merge_kernel(const float2* __restrict__ part, const float* __restrict__ pos, int M, int tiles_n,
                                 const int64_t* __restrict__ labels, int64_t label_offset, int N,
                                 float* __restrict__ row_lse, float* __restrict__ row_loss) {
  const int m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= M) return;
  float mx = -INFINITY;
  for (int t = 0; t < tiles_n; ++t) mx = fmaxf(mx, part[(int64_t)t * M + m].x);
  float sm = 0.f;
  for (int t = 0; t < tiles_n; ++t) {
    const float2 q = part[(int64_t)t * M + m];
    if (q.x > -INFINITY) sm += q.y * expf(q.x - mx);
  }
  const float lse = mx + logf(sm);
  row_lse[m] = lse;
  const int64_t lab = labels != nullptr ? labels[m] : (int64_t)m + label_offset;
  row_loss[m] = (lab >= 0 && lab < N) ? lse - pos[m] : 0.f;
}

static inline int round_up_i(int x, int m) { return (x + m - 1) / m * m; }
constexpr int kCeChunkRows = 4096;

// The feature-CE GEMMs run on the CTA-pair engine (256-row tiles) whenever the problem has more than one 128-row
// tile; CLIPK_CE_ENGINE=1|2 forces one (A/B measurements, pipeline tests).
static int ce_engine(int M) {
  static const int forced = [] {
    const char* e = getenv("CLIPK_CE_ENGINE");
    return e != nullptr ? atoi(e) : 0;
  }();
  if (forced == 1 || forced == 2) return forced;
  return M > eng::BM ? 2 : 1;
}
static bool ce_dl_tma() {       // CLIPK_CE_DLTMA=0: per-thread stores of dL (A/B measurements)
  static const bool on = [] {
    const char* e = getenv("CLIPK_CE_DLTMA");
    return e == nullptr || atoi(e) != 0;
  }();
  return on;
}
template <int BN, bool A_MN, bool B_MN, class Epi>
static int ce_launch(int engine, const OperandDesc* a, const OperandDesc* b, const int* ks, int M, int N,
                     const typename Epi::Params& ep, cudaStream_t st) {
  if (engine == 2) return launch_gemm2<BN, A_MN, B_MN, Epi>(a, b, 1, ks, ks, M, N, 1, ep, st);
  return launch_gemm<BN, A_MN, B_MN, Epi>(a, b, 1, ks, ks, M, N, 1, ep, st);
}

// bf16 output variant (gradients wanted in bf16: no fp32 round trip + cast pass)
__global__ void ce_slab_reduce_bf16_kernel(const float4* __restrict__ slabs, int nslab, int64_t n4,
                                           uint2* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nslab; ++s) {
    const float4 b = slabs[(int64_t)s * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = make_uint2(ptx::pack_bf16x2(a.x, a.y), ptx::pack_bf16x2(a.z, a.w));
}

struct CeWorkspace {
  float2* part;
  float* pos;
  __nv_bfloat16* dL;
  float* slabs;      // split-K slabs of dX: [nsplit][mc][D] fp32 (only when the dX GEMM is split)
};

// dX = dL Y has only ceil(mc/256) * ceil(D/256) output tiles (48 for mc = 4096, D = 768) for 74 CTA pairs and a very
// long K (= N): split K so that the tile count fills whole waves.  Returns the split count (1 = no split) and the
// columns per split (a multiple of 64).
struct DxSplit {
  int nsplit;
  int ksplit;
};
static DxSplit dx_split(int mc, int N, int D) {
  DxSplit r{1, N};
  if (mc <= eng::BM || N < 8192) return r;
  const int tiles = ((mc + 255) / 256) * ((D + 255) / 256);
  const int clusters = sm_count() / 2;
  double best = (double)tiles / (((tiles + clusters - 1) / clusters) * (double)clusters);
  for (int s = 2; s <= 8; ++s) {
    if (N / s < 2048) break;
    const int t = tiles * s;
    const double eff = (double)t / (((t + clusters - 1) / clusters) * (double)clusters);
    if (eff > best + 0.08) {       // a split costs a slab round trip: take it only for a clear gain
      best = eff;
      r.nsplit = s;
    }
  }
  if (r.nsplit > 1) r.ksplit = ((N + r.nsplit - 1) / r.nsplit + 63) / 64 * 64;    // nsplit splits cover N; tail zero-filled
  return r;
}

// dX[i] = (accumulate ? dX[i] : 0) + sum_s slabs[s][i]      (n % 4 == 0, fixed order)
__global__ void ce_slab_reduce_kernel(const float4* __restrict__ slabs, int nslab, int64_t n4, int accumulate,
                                      float4* __restrict__ out) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 a = accumulate ? out[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  for (int s = 0; s < nslab; ++s) {
    const float4 b = slabs[(int64_t)s * n4 + i];
    a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
  }
  out[i] = a;
}
static size_t ce_carve(CeWorkspace* w, void* base, int M, int N, int D = 0) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    void* p = base ? static_cast<char*>(base) + off : nullptr;
    off += (bytes + 1023) / 1024 * 1024;
    return p;
  };
  const int tiles_n = 2 * ((N + 255) / 256);
  const int mc = M < kCeChunkRows ? M : kCeChunkRows;
  w->part = static_cast<float2*>(take((size_t)M * tiles_n * sizeof(float2)));
  w->pos = static_cast<float*>(take((size_t)M * 4));
  w->dL = static_cast<__nv_bfloat16*>(take((size_t)mc * round_up_i(N, 64) * 2));
  w->slabs = nullptr;
  if (D > 0) {
    const DxSplit sp = dx_split(mc, N, D);
    if (sp.nsplit > 1) w->slabs = static_cast<float*>(take((size_t)sp.nsplit * mc * D * 4));
  }
  return off;
}

int ce_feat_fwd(const __nv_bfloat16* X, const __nv_bfloat16* Y, int M, int N, int D, float scale, float bias,
                const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows, const int64_t* labels, int64_t label_offset, float* row_lse, float* row_loss, void* ws,
                size_t ws_bytes, cudaStream_t st) {
  CLIPK_REQUIRE(M > 0 && N > 0 && D > 0 && D % 8 == 0, "ce_feat_fwd: bad shape M=%d N=%d D=%d (D %% 8 == 0)", M, N, D);
  CLIPK_REQUIRE(slab_counts == nullptr || (slab_rows > 0 && slab_rows % 32 == 0 && slab_n0 % 32 == 0),
                "ce_feat: hard-negative slabs need slab_n0 (%d) and slab_rows (%d) to be multiples of 32", slab_n0, slab_rows);
  CeWorkspace w{};
  const size_t need = ce_carve(&w, ws, M, N);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "ce_feat_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  OperandDesc a, b;
  a.ptr = X; a.rows = M; a.k = D; a.ld = D;
  b.ptr = Y; b.rows = N; b.k = D; b.ld = D;
  const int ks[1] = {(D + 63) / 64};
  const int tiles_n = 2 * ((N + 255) / 256);   // one partial per (n-tile, epilogue-warp half)
  CLIPK_CHECK_CUDA(cudaMemsetAsync(w.pos, 0, (size_t)M * 4, st));
  epi::LsePart::Params ep{w.part, w.pos, labels, label_offset, M, N, tiles_n, scale, bias, scale_dev, slab_counts, slab_n0, slab_rows};
  CLIPK_TRY((ce_launch<256, false, false, epi::LsePart>(ce_engine(M), &a, &b, ks, M, N, ep, st)));
  lse_merge_kernel<<<(M + 127) / 128, 128, 0, st>>>(w.part, w.pos, M, tiles_n, labels, label_offset, N, row_lse, row_loss);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// Symmetric CE forward: row lse / loss as ce_feat_fwd (labels = m + label_offset) plus the column sums of the first
// `ncol` columns (col_sum, zeroed here).  CTA-pair engine only.
int ce_sym_fwd(const __nv_bfloat16* X, const __nv_bfloat16* Y, int M, int N, int D, float scale, float bias,
               const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows, int64_t label_offset,
               int ncol, float* row_lse, float* row_loss, float* col_sum, void* ws, size_t ws_bytes, cudaStream_t st) {
  CLIPK_REQUIRE(M > 0 && N > 0 && D > 0 && D % 8 == 0, "ce_sym_fwd: bad shape M=%d N=%d D=%d (D %% 8 == 0)", M, N, D);
  CLIPK_REQUIRE(ncol >= 0 && ncol <= N, "ce_sym_fwd: ncol=%d out of range (N=%d)", ncol, N);
  CLIPK_REQUIRE(slab_counts == nullptr || (slab_rows > 0 && slab_rows % 32 == 0 && slab_n0 % 32 == 0),
                "ce_sym: hard-negative slabs need slab_n0 (%d) and slab_rows (%d) to be multiples of 32", slab_n0, slab_rows);
  CeWorkspace w{};
  const size_t need = ce_carve(&w, ws, M, N);
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "ce_sym_fwd: workspace too small (%zu < %zu)", ws_bytes, need);
  OperandDesc a, b;
  a.ptr = X; a.rows = M; a.k = D; a.ld = D;
  b.ptr = Y; b.rows = N; b.k = D; b.ld = D;
  const int ks[1] = {(D + 63) / 64};
  const int tiles_n = 2 * ((N + 255) / 256);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(w.pos, 0, (size_t)M * 4, st));
  if (ncol > 0) CLIPK_CHECK_CUDA(cudaMemsetAsync(col_sum, 0, (size_t)ncol * 4, st));
  epi::LseRowCol::Params ep{w.part, w.pos, label_offset, M, N, tiles_n, scale, bias, scale_dev, slab_counts, slab_n0,
                            slab_rows, col_sum, ncol};
  CLIPK_TRY((launch_gemm2<256, false, false, epi::LseRowCol>(&a, &b, 1, ks, ks, M, N, 1, ep, st)));
  lse_merge_kernel<<<(M + 127) / 128, 128, 0, st>>>(w.part, w.pos, M, tiles_n, nullptr, label_offset, N, row_lse, row_loss);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

struct CeSymBwd {           // column part of the symmetric backward (null col_bias: plain row CE)
  const float* col_bias = nullptr;
  const float* col_w = nullptr;
  int ncol = 0;
  int phase = 0;            // 0: dL, dX, dY;  1: dL + dY only;  2: dX only from the dL phase 1 left in the workspace
};

int ce_feat_bwd(const __nv_bfloat16* X, const __nv_bfloat16* Y, int M, int N, int D, float scale, float bias,
                const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows, const int64_t* labels, int64_t label_offset, const float* row_lse, const float* row_w, void* dXv,
                int accX, void* dYv, int accY, int grads_bf16, void* ws, size_t ws_bytes, cudaStream_t st,
                const CeSymBwd& sym = CeSymBwd()) {
  CLIPK_REQUIRE(M > 0 && N > 0 && D > 0 && D % 8 == 0, "ce_feat_bwd: bad shape M=%d N=%d D=%d (D %% 8 == 0)", M, N, D);
  // bf16 gradients: written straight from the GEMM epilogues (TMA stores) / the slab sum; needs a single row chunk
  // (dY is not accumulated across chunks) and the CTA-pair engine
  CLIPK_REQUIRE(!grads_bf16 || (accX == 0 && accY == 0 && M <= kCeChunkRows && M > eng::BM && N > eng::BM),
                "ce_feat_bwd: bf16 gradients need %d < M <= %d, N > %d and no accumulation (M=%d N=%d)", eng::BM,
                kCeChunkRows, eng::BM, M, N);
  float* dX = static_cast<float*>(dXv);
  float* dY = static_cast<float*>(dYv);
  CeWorkspace w{};
  size_t need = ce_carve(&w, ws, M, N, D);
  if (ws_bytes < need) need = ce_carve(&w, ws, M, N);       // sized by the D-less query: run the dX GEMM unsplit
  CLIPK_REQUIRE(ws != nullptr && ws_bytes >= need, "ce_feat_bwd: workspace too small (%zu < %zu)", ws_bytes, need);
  const int64_t ldd = round_up_i(N, 64);
  const int ksD[1] = {(D + 63) / 64};
  for (int m0 = 0; m0 < M; m0 += kCeChunkRows) {
    const int mc = (M - m0) < kCeChunkRows ? (M - m0) : kCeChunkRows;
    // recompute logits tile -> dL (bf16)
    {
      OperandDesc a, b;
      a.ptr = X + (int64_t)m0 * D; a.rows = mc; a.k = D; a.ld = D;
      b.ptr = Y; b.rows = N; b.k = D; b.ld = D;
      if (sym.phase == 2) {
        // dL of this (single) row chunk is still in the workspace
      } else if (sym.col_bias != nullptr) {
        epi::DlSymTma::Params ep{{w.dL, ldd, (int64_t)mc * ldd, mc, N, 1}, row_lse + m0, row_w + m0, label_offset + m0, mc, N,
                                 scale, bias, scale_dev, slab_counts, slab_n0, slab_rows, sym.col_bias, sym.col_w, sym.ncol};
        CLIPK_TRY((launch_gemm2<256, false, false, epi::DlSymTma>(&a, &b, 1, ksD, ksD, mc, N, 1, ep, st)));
      } else if (ce_engine(mc) == 2 && ce_dl_tma()) {
        epi::DlOutTma::Params ep{{w.dL, ldd, (int64_t)mc * ldd, mc, N, 1}, row_lse + m0, row_w + m0,
                                 labels ? labels + m0 : nullptr, label_offset + (labels ? 0 : m0), mc, N, scale, bias, scale_dev, slab_counts, slab_n0, slab_rows};
        CLIPK_TRY((launch_gemm2<256, false, false, epi::DlOutTma>(&a, &b, 1, ksD, ksD, mc, N, 1, ep, st)));
      } else {
        epi::DlOut::Params ep{row_lse + m0, row_w + m0, labels ? labels + m0 : nullptr, label_offset + (labels ? 0 : m0),
                              w.dL, ldd, mc, N, scale, bias, scale_dev, slab_counts, slab_n0, slab_rows};
        CLIPK_TRY((ce_launch<256, false, false, epi::DlOut>(ce_engine(mc), &a, &b, ksD, mc, N, ep, st)));
      }
    }
    // dX[m0:m0+mc] (+)= scale * dL Y          (A = dL K-major over n, B = Y MN-major)
    const DxSplit sp = dx_split(mc, N, D);
    if (sym.phase == 1) {
      // the caller wants dY first (its reduce-scatter overlaps the dX GEMM of phase 2)
    } else if (dX != nullptr && sp.nsplit > 1 && w.slabs != nullptr && D % 4 == 0) {
      // split-K: the splits are the batches of ONE launch (slab s = batch s); batch s starts at column s * ksplit of
      // the single long reduction dim, the operand maps keep the true extent N, so the last split's tail is
      // zero-filled by TMA; then a fixed-order slab sum
      const int nfull = (N + sp.ksplit - 1) / sp.ksplit;
      {
        OperandDesc a, b;
        a.ptr = w.dL; a.rows = mc; a.k = N; a.ld = ldd; a.k_batch_offset = sp.ksplit;
        b.ptr = Y; b.mn_major = true; b.rows = D; b.k = N; b.ld = D;
        const int ks[1] = {(sp.ksplit + 63) / 64};
        epi::Store<false>::Params ep{w.slabs, D, (int64_t)mc * D, mc, D, scale, 0, scale_dev};
        CLIPK_TRY((launch_gemm2<256, false, true, epi::Store<false>>(&a, &b, 1, ks, ks, mc, D, nfull, ep, st)));
      }
      const int64_t n4 = (int64_t)mc * D / 4;
      if (grads_bf16)
        ce_slab_reduce_bf16_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(w.slabs), nfull, n4, reinterpret_cast<uint2*>(dXv));
      else
        ce_slab_reduce_kernel<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(
            reinterpret_cast<const float4*>(w.slabs), nfull, n4, accX, reinterpret_cast<float4*>(dX + (int64_t)m0 * D));
      clipk::count_launches(1);
      CLIPK_CHECK_CUDA(cudaGetLastError());
    } else if (dX != nullptr) {
      OperandDesc a, b;
      a.ptr = w.dL; a.rows = mc; a.k = N; a.ld = ldd;
      b.ptr = Y; b.mn_major = true; b.rows = D; b.k = N; b.ld = D;
      const int ks[1] = {(N + 63) / 64};
      if (grads_bf16) {
        epi::StoreTma::Params ep{{dXv, D, (int64_t)mc * D, mc, D, 1}, scale, scale_dev};
        CLIPK_TRY((launch_gemm2<256, false, true, epi::StoreTma>(&a, &b, 1, ks, ks, mc, D, 1, ep, st)));
      } else {
        epi::Store<false>::Params ep{dX + (int64_t)m0 * D, D, 0, mc, D, scale, accX, scale_dev};
        CLIPK_TRY((ce_launch<256, false, true, epi::Store<false>>(ce_engine(mc), &a, &b, ks, mc, D, ep, st)));
      }
    }
    // dY (+)= scale * dL^T X[m0:m0+mc]        (A = dL MN-major (rows = n), B = X MN-major)
    if (dY != nullptr && sym.phase != 2) {
      OperandDesc a, b;
      a.ptr = w.dL; a.mn_major = true; a.rows = N; a.k = mc; a.ld = ldd;
      b.ptr = X + (int64_t)m0 * D; b.mn_major = true; b.rows = D; b.k = mc; b.ld = D;
      const int ks[1] = {(mc + 63) / 64};
      if (grads_bf16) {
        epi::StoreTma::Params ep{{dYv, D, (int64_t)N * D, N, D, 1}, scale, scale_dev};
        CLIPK_TRY((launch_gemm2<256, true, true, epi::StoreTma>(&a, &b, 1, ks, ks, N, D, 1, ep, st)));
      } else {
        epi::Store<false>::Params ep{dY, D, 0, N, D, scale, (accY || m0 > 0) ? 1 : 0, scale_dev};
        CLIPK_TRY((ce_launch<256, true, true, epi::Store<false>>(ce_engine(N), &a, &b, ks, N, D, ep, st)));
      }
    }
  }
  return 0;
}

}  // namespace clipk

extern "C" {

int clipk_ce_rows(const float* L, int M, int N, int64_t ld, const int64_t* labels, int64_t label_offset,
                  float* row_lse, float* row_loss, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M >= 0 && N > 0 && ld >= N, "ce_rows: bad shape M=%d N=%d ld=%lld", M, N, (long long)ld);
  if (M == 0) return 0;
  clipk::ce_rows_kernel<<<(M + 7) / 8, 256, 0, static_cast<cudaStream_t>(stream)>>>(L, M, N, ld, labels, label_offset,
                                                                                    row_lse, row_loss);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_ce_cols(const float* L, int M, int N, int64_t ld, float* col_max, float* col_sum, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M >= 0 && N > 0 && ld >= N, "ce_cols: bad shape M=%d N=%d ld=%lld", M, N, (long long)ld);
  clipk::ce_cols_kernel<<<(N + 31) / 32, 1024, 0, static_cast<cudaStream_t>(stream)>>>(L, M, N, ld, col_max, col_sum);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_ce_rowsums(const float* row_lse, const float* row_loss, int M, float* out2, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M > 0, "ce_rowsums: empty input");
  clipk::ce_rowsums_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(row_lse, row_loss, M, out2);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_ce_merge(const float* gathered, int W, int N, float* col_lse, float* loss, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(W > 0 && N > 0, "ce_merge: empty input");
  clipk::ce_merge_kernel<<<1, 1024, 0, static_cast<cudaStream_t>(stream)>>>(gathered, W, N, col_lse, loss);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_ce_scores_grad(const float* L, int M, int N, int64_t ld, const float* row_lse, const float* col_lse,
                         int64_t label_offset, float w_row, float w_col, float* dL, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M >= 0 && N > 0 && ld >= N, "ce_scores_grad: bad shape M=%d N=%d ld=%lld", M, N, (long long)ld);
  if (M == 0) return 0;
  const int64_t n = (int64_t)M * N;
  clipk::ce_scores_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      L, M, N, ld, row_lse, col_lse, label_offset, w_row, w_col, dL);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_ce_rows_grad(const float* L, int M, int N, int64_t ld, const float* row_lse, const int64_t* labels,
                       int64_t label_offset, const float* row_w, float* dL, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M >= 0 && N > 0 && ld >= N, "ce_rows_grad: bad shape M=%d N=%d ld=%lld", M, N, (long long)ld);
  if (M == 0) return 0;
  const int64_t n = (int64_t)M * N;
  clipk::ce_rows_grad_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      L, M, N, ld, row_lse, labels, label_offset, row_w, dL);
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_sgemm_f32(const float* A, int64_t sam, int64_t sak, const float* B, int64_t sbk, int64_t sbn, float* C,
                    int64_t ldc, int M, int N, int K, float alpha, float beta, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(M >= 0 && N >= 0 && K >= 0, "sgemm: bad shape");
  if (M == 0 || N == 0) return 0;
  if ((int64_t)((M + 127) / 128) * ((N + 127) / 128) >= 48) {      // enough 128 x 128 tiles to keep the GPU busy (measured)
    dim3 grid((N + 127) / 128, (M + 127) / 128);
    clipk::sgemm128_strided_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, sam, sak, B, sbk, sbn, C, ldc,
                                                                                        M, N, K, alpha, beta);
  } else {
    dim3 grid((N + 63) / 64, (M + 63) / 64);
    clipk::sgemm_strided_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(A, sam, sak, B, sbk, sbn, C, ldc, M,
                                                                                     N, K, alpha, beta);
  }
  clipk::count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

size_t clipk_ce_feat_workspace_bytes(int M, int N) {
  clipk::CeWorkspace w{};
  return clipk::ce_carve(&w, nullptr, M, N);
}

size_t clipk_ce_feat_bwd_workspace_bytes(int M, int N, int D) {
  clipk::CeWorkspace w{};
  return clipk::ce_carve(&w, nullptr, M, N, D);
}

int clipk_ce_feat_fwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias,
        const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows,
                      const int64_t* labels, int64_t label_offset, float* row_lse, float* row_loss, void* workspace,
                      size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::ce_feat_fwd(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), M, N, D, scale,
                            bias, scale_dev, slab_counts, slab_n0, slab_rows, labels, label_offset, row_lse, row_loss, workspace, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

int clipk_ce_feat_bwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias,
        const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows,
                      const int64_t* labels, int64_t label_offset, const float* row_lse, const float* row_w, float* dX,
                      int accX, float* dY, int accY, void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::ce_feat_bwd(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), M, N, D, scale,
                            bias, scale_dev, slab_counts, slab_n0, slab_rows, labels, label_offset, row_lse, row_w, dX, accX, dY, accY, 0, workspace, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}

int clipk_ce_sym_fwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                     const int* slab_counts, int slab_n0, int slab_rows, int64_t label_offset, int ncol, float* row_lse,
                     float* row_loss, float* col_sum, void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::ce_sym_fwd(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), M, N, D, scale, bias,
                           scale_dev, slab_counts, slab_n0, slab_rows, label_offset, ncol, row_lse, row_loss, col_sum,
                           workspace, ws_bytes, static_cast<cudaStream_t>(stream));
}

int clipk_ce_sym_bwd(const void* X, const void* Y, int M, int N, int D, float scale, float bias, const float* scale_dev,
                     const int* slab_counts, int slab_n0, int slab_rows, int64_t label_offset, int ncol,
                     const float* row_lse, const float* row_w, const float* col_bias, const float* col_w, void* dX,
                     void* dY, int grads_bf16, int phase, void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  CLIPK_REQUIRE(col_bias != nullptr && col_w != nullptr && ncol > 0 && ncol <= N, "ce_sym_bwd: column terms missing");
  CLIPK_REQUIRE(phase >= 0 && phase <= 2 && (phase == 0 || M <= clipk::kCeChunkRows),
                "ce_sym_bwd: the two-phase backward needs a single row chunk (M=%d <= %d)", M, clipk::kCeChunkRows);
  clipk::CeSymBwd sym;
  sym.col_bias = col_bias;
  sym.col_w = col_w;
  sym.ncol = ncol;
  sym.phase = phase;
  return clipk::ce_feat_bwd(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), M, N, D, scale,
                            bias, scale_dev, slab_counts, slab_n0, slab_rows, nullptr, label_offset, row_lse, row_w, dX, 0,
                            dY, 0, grads_bf16, workspace, ws_bytes, static_cast<cudaStream_t>(stream), sym);
}

int clipk_ce_feat_bwd_bf16(const void* X, const void* Y, int M, int N, int D, float scale, float bias,
        const float* scale_dev, const int* slab_counts, int slab_n0, int slab_rows,
                           const int64_t* labels, int64_t label_offset, const float* row_lse, const float* row_w,
                           void* dX, void* dY, void* workspace, size_t ws_bytes, void* stream) {
  CLIPK_TRY(clipk::check_device());
  return clipk::ce_feat_bwd(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), M, N, D, scale,
                            bias, scale_dev, slab_counts, slab_n0, slab_rows, labels, label_offset, row_lse, row_w, dX, 0, dY, 0, 1, workspace, ws_bytes,
                            static_cast<cudaStream_t>(stream));
}
}
