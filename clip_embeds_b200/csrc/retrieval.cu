// Retrieval ranks fused into the logits GEMM: open_clip's get_clip_metrics
// (open_clip/src/open_clip_train/train.py:360-377) ranks the ground truth of every query by a full argsort of the
// [N, N] logits on the CPU.  Here the rank of the diagonal is counted tile by tile in the epilogue of the tcgen05
// GEMM -- rank_i = #{ j != i : logit[i,j] > logit[i,i] } for the rows (image -> text) and, from the same tiles,
// #{ i != j : logit[i,j] > logit[j,j] } for the columns (text -> image) -- and the logits are never written.
#include "common.cuh"
#include "epilogues.cuh"

namespace epi {

struct RankCount {
  struct Params {
    const float* diag;     // [min(M, N)] raw <x_i, y_i> (same scale as the accumulator)
    int* count_row;        // [M]
    int* count_col;        // [N]
    int M, N;
  };
  Params p;
  float dr;
  int cnt;
  __device__ explicit RankCount(const Params& pp) : p(pp), dr(0.f), cnt(0) {}
  __device__ void tile_begin(int, int m, int) {
    cnt = 0;
    dr = (m < p.M && m < p.N) ? __ldg(p.diag + m) : INFINITY;
  }
  __device__ void chunk(int, int m, int n, float* v) {
    if (n >= p.N) return;                                   // warp-uniform
    const int lane = (int)ptx::lane_id();
    const float dc_l = (n + lane < p.N && n + lane < p.M) ? __ldg(p.diag + n + lane) : INFINITY;
    const bool row_ok = m < p.M;
    int mine = 0;
#pragma unroll
    for (int j = 0; j < 32; ++j) {
      const bool ok = row_ok && (n + j < p.N) && (n + j != m);
      cnt += (ok && v[j] > dr) ? 1 : 0;
      const float dc = __shfl_sync(0xffffffffu, dc_l, j);
      const unsigned bal = __ballot_sync(0xffffffffu, ok && v[j] > dc);
      if (lane == j) mine = __popc(bal);
    }
    if (mine != 0) atomicAdd(p.count_col + n + lane, mine);
  }
  __device__ void tile_end(int, int m, int, int, int) {
    if (m < p.M && cnt != 0) atomicAdd(p.count_row + m, cnt);
  }
};

}  // namespace epi

namespace clipk {

// diag[i] = <x_i, y_i> (bf16 inputs, fp32 accumulation); one warp per row
__global__ void rowdot_bf16_kernel(const __nv_bfloat16* __restrict__ X, const __nv_bfloat16* __restrict__ Y, int n, int D,
                                   float* __restrict__ out) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n) return;
  const int lane = threadIdx.x & 31;
  float s = 0.f;
  for (int d = lane; d < D; d += 32)
    s = fmaf(__bfloat162float(X[(int64_t)row * D + d]), __bfloat162float(Y[(int64_t)row * D + d]), s);
  s = ptx::warp_sum(s);
  if (lane == 0) out[row] = s;
}

// out4 = (sum of ranks, #rank < 1, #rank < 5, #rank < 10) as int64
__global__ void rank_stats_kernel(const int* __restrict__ ranks, int n, unsigned long long* __restrict__ out4) {
  unsigned long long s = 0, r1 = 0, r5 = 0, r10 = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int r = ranks[i];
    s += (unsigned long long)r;
    r1 += r < 1;
    r5 += r < 5;
    r10 += r < 10;
  }
  atomicAdd(out4 + 0, s);
  atomicAdd(out4 + 1, r1);
  atomicAdd(out4 + 2, r5);
  atomicAdd(out4 + 3, r10);
}

}  // namespace clipk

extern "C" {

// X bf16 [M, D] (queries of the rows), Y bf16 [N, D]; ground truth of row i is column i, of column j row j.
// diag_ws: fp32 [min(M, N)] scratch.  rank_row [M], rank_col [N]: int32 (0 = ground truth ranked first).
int clipk_retrieval_ranks(const void* X, const void* Y, int M, int N, int D, float* diag_ws, int* rank_row,
                          int* rank_col, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(M > 0 && N > 0 && D > 0 && D % 8 == 0, "retrieval_ranks: bad shape M=%d N=%d D=%d (D %% 8 == 0)", M, N, D);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nd = M < N ? M : N;
  rowdot_bf16_kernel<<<(nd + 7) / 8, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(X), static_cast<const __nv_bfloat16*>(Y), nd, D, diag_ws);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(rank_row, 0, (size_t)M * sizeof(int), st));
  CLIPK_CHECK_CUDA(cudaMemsetAsync(rank_col, 0, (size_t)N * sizeof(int), st));
  OperandDesc a, b;
  a.ptr = X; a.rows = M; a.k = D; a.ld = D;
  b.ptr = Y; b.rows = N; b.k = D; b.ld = D;
  const int ks[1] = {(D + 63) / 64};
  epi::RankCount::Params ep{diag_ws, rank_row, rank_col, M, N};
  if (M > eng::BM) return launch_gemm2<256, false, false, epi::RankCount>(&a, &b, 1, ks, ks, M, N, 1, ep, st);
  return launch_gemm<256, false, false, epi::RankCount>(&a, &b, 1, ks, ks, M, N, 1, ep, st);
}

// stats4 (device, uint64 [4]) = (sum of ranks, #rank<1, #rank<5, #rank<10)
int clipk_rank_stats(const int* ranks, int n, unsigned long long* stats4, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(n > 0, "rank_stats: empty input");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(stats4, 0, 4 * sizeof(unsigned long long), st));
  const int blocks = (n + 255) / 256 < 128 ? (n + 255) / 256 : 128;
  rank_stats_kernel<<<blocks, 256, 0, st>>>(ranks, n, stats4);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
