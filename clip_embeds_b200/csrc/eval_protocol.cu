// Batched eval protocol: the accuracy accounting of the reference's eval scripts on the device,
// fed by the one-launch scorer (`clipk_pacl_paired_fwd` with v_div = K) instead of one tiny forward per item.
//
//   What'sUp / COCO-spatial / GQA-spatial   PACL/eval_pacl.py:26-104 (eval), :106-186 (eval_4)
//       correct_i = score[i,0] > score[i,k] for all k > 0 (strict; the ground truth is always caption 0);
//       eval_dict[(object pair)][relation] = correct (a later item of the same key overwrites an earlier one);
//       individual / pair (left&right, on&under, in-front&behind) / set (all four of a set correct) counts.
//   MMVP-style pairs                         PACL/eval_pacl.py:268-349
//       per pair two images x two texts; pred_t = img1 iff softmax([s(img1,t), s(img2,t)])[0] > 0.5 (fp32 softmax);
//       pair / single counts per category (category = pair index / pairs_per_category).
// Integer bookkeeping: results are bit-exact against the oracle's restatement of the same lines.
#include "common.cuh"

namespace clipk {

__global__ void eval_correct_kernel(const float* __restrict__ scores, int items, int K, int* __restrict__ correct) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= items) return;
  const float s0 = scores[(int64_t)i * K];
  int ok = 1;
  for (int k = 1; k < K; ++k) ok &= (s0 > scores[(int64_t)i * K + k]) ? 1 : 0;
  correct[i] = ok;
}

// winner[set][rel] = largest item index with that key (the reference's dict assignment: the last item wins)
__global__ void eval_winner_kernel(const int* __restrict__ set_id, const int* __restrict__ rel_id, int items, int nsets,
                                   int* __restrict__ winner) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= items) return;
  const int s = set_id[i], r = rel_id[i];
  if (s < 0 || s >= nsets || r < 0 || r >= 6) return;
  atomicMax(winner + s * 6 + r, i);
}

// counts: [0..2] individual (left/right, on/under, in-front/behind), [3..5] pairs, [6] sets, [7] items
__global__ void eval_whatsup_count_kernel(const int* __restrict__ correct, const int* __restrict__ winner, int nsets,
                                          int items, int* __restrict__ counts) {
  __shared__ int sh[8];
  if (threadIdx.x < 8) sh[threadIdx.x] = 0;
  __syncthreads();
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < nsets; s += gridDim.x * blockDim.x) {
    int c[6];
#pragma unroll
    for (int r = 0; r < 6; ++r) {
      const int w = winner[s * 6 + r];
      c[r] = w >= 0 ? correct[w] : 0;
    }
    int tot = 0;
#pragma unroll
    for (int g = 0; g < 3; ++g) {
      const int ind = c[2 * g] + c[2 * g + 1];
      tot += ind;
      if (ind) atomicAdd(&sh[g], ind);
      if (c[2 * g] && c[2 * g + 1]) atomicAdd(&sh[3 + g], 1);
    }
    if (tot == 4) atomicAdd(&sh[6], 1);          // eval_pacl.py:82: sum(correct_dict.values()) == 4
  }
  __syncthreads();
  if (threadIdx.x < 7 && sh[threadIdx.x]) atomicAdd(counts + threadIdx.x, sh[threadIdx.x]);
  if (blockIdx.x == 0 && threadIdx.x == 7) counts[7] = items;
}

// s1, s2 [pairs][2]: diagonal scores of image 1 / image 2 against (text 1, text 2); gt [pairs][2]: 1 = img1.
// pred [pairs][2]; counts [ncat][2] = (pairs with both right, single predictions right)
__global__ void eval_mmvp_kernel(const float* __restrict__ s1, const float* __restrict__ s2, const int* __restrict__ gt,
                                 int pairs, int pairs_per_cat, int ncat, int* __restrict__ pred,
                                 int* __restrict__ counts) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pairs) return;
  int ok = 0;
  for (int t = 0; t < 2; ++t) {
    const float a = s1[2 * i + t], b = s2[2 * i + t];
    // softmax([a, b])[0] > 0.5 in fp32 (eval_pacl.py:310-316): exp(x - max) / sum
    const float m = fmaxf(a, b);
    const float ea = expf(a - m), eb = expf(b - m);
    const int p = (ea / (ea + eb) > 0.5f) ? 1 : 0;
    pred[2 * i + t] = p;
    ok += (p == (gt[2 * i + t] != 0 ? 1 : 0)) ? 1 : 0;
  }
  int cat = pairs_per_cat > 0 ? i / pairs_per_cat : 0;
  if (cat >= ncat) cat = ncat - 1;
  if (ok == 2) atomicAdd(counts + 2 * cat, 1);
  if (ok) atomicAdd(counts + 2 * cat + 1, ok);
}

}  // namespace clipk

extern "C" {

int clipk_eval_correct(const float* scores, int items, int K, int* correct, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(items >= 0 && K >= 1, "eval_correct: bad shape items=%d K=%d", items, K);
  if (items == 0) return 0;
  eval_correct_kernel<<<(items + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(scores, items, K, correct);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// winner_ws: int32 [nsets * 6] scratch; counts: int32 [8] (overwritten)
int clipk_eval_whatsup(const int* correct, const int* set_id, const int* rel_id, int items, int nsets, int* winner_ws,
                       int* counts, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(items > 0 && nsets > 0, "eval_whatsup: empty input (items=%d nsets=%d)", items, nsets);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(winner_ws, 0xFF, (size_t)nsets * 6 * sizeof(int), st));   // -1
  CLIPK_CHECK_CUDA(cudaMemsetAsync(counts, 0, 8 * sizeof(int), st));
  eval_winner_kernel<<<(items + 255) / 256, 256, 0, st>>>(set_id, rel_id, items, nsets, winner_ws);
  const int blocks = (nsets + 255) / 256 < 64 ? (nsets + 255) / 256 : 64;
  eval_whatsup_count_kernel<<<blocks, 256, 0, st>>>(correct, winner_ws, nsets, items, counts);
  count_launches(2);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int clipk_eval_mmvp(const float* s_img1, const float* s_img2, const int* gt, int pairs, int pairs_per_cat, int ncat,
                    int* pred, int* counts, void* stream) {
  using namespace clipk;
  CLIPK_TRY(check_device());
  CLIPK_REQUIRE(pairs > 0 && ncat > 0, "eval_mmvp: empty input (pairs=%d ncat=%d)", pairs, ncat);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  CLIPK_CHECK_CUDA(cudaMemsetAsync(counts, 0, (size_t)ncat * 2 * sizeof(int), st));
  eval_mmvp_kernel<<<(pairs + 255) / 256, 256, 0, st>>>(s_img1, s_img2, gt, pairs, pairs_per_cat, ncat, pred, counts);
  count_launches(1);
  CLIPK_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
