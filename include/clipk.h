/* clipk — C ABI of the B200-native (sm_100a) patch-aligned contrastive scoring + loss kernels.
 *
 * The reference (lst627/CLIP-Embeds) has no FFI / plugin registry: its boundary for this path is Python
 * (`nn.Module.forward` + autograd).  This header is the boundary a maintainer binds instead of the ATen op
 * sequences cited on each entry point (paths relative to the reference root;
 * PACL = Patch-Aligned-Contrastive-Learning).  Python binding: clip_embeds_b200/_lib.py (ctypes), see
 * INTEGRATION.md.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the name ends in _host; tensors are row-major and contiguous
 *     unless a leading dimension is given; sizes are element counts;
 *   - the library never allocates, frees or synchronises: the caller owns inputs, outputs, saved statistics
 *     and workspaces (`*_workspace_bytes` queries); all work is enqueued on `stream` (a cudaStream_t);
 *   - return 0 on success, a negative CLIPK_ERR_* otherwise; clipk_last_error() gives a thread-local message;
 *   - requires an sm_100 device: any other device returns CLIPK_ERR_ARCH (there is no fallback path);
 *   - re-entrant, no global mutable state besides per-device attribute caches.
 */
#ifndef CLIPK_H_
#define CLIPK_H_
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CLIPK_OK 0
#define CLIPK_ERR_INVALID (-1)
#define CLIPK_ERR_CUDA (-2)
#define CLIPK_ERR_ARCH (-3)

#define CLIPK_BF16 0
#define CLIPK_F32 1

#define CLIPK_ACT_SIGMOID10 0 /* sigmoid(10*cos): reference parity, pacl.py:133 */
#define CLIPK_ACT_ONES 1      /* activations overwritten with ones: the checked-in forward(), pacl.py:141-142 */

const char* clipk_last_error(void);
int clipk_version(void);
/* 0 when the current CUDA device can run the kernels (compute capability 10.x), else CLIPK_ERR_ARCH. */
int clipk_check_device(void);

/* ---------------------------------------------------------------------------------------------------------
 * Engine building block (also the unit-test entry of the tcgen05 pipeline):
 *   C[b] (+)= alpha * A[b] * B[b]^T,   A: M x K, B: N x K, bf16 operands, fp32 accumulation in TMEM.
 *   a_mn / b_mn = 0: operand stored [rows][K] (K contiguous, leading dim ld = row stride)
 *               = 1: operand stored [K][rows] (rows contiguous, ld = stride between k)
 *   out_dtype CLIPK_F32 | CLIPK_BF16; accumulate != 0 adds into C (fp32 only).
 * Replaces: torch `@` / `einsum` / `bmm` on the path (cuBLAS), e.g. pacl.py:129, :460, :472, :499-500.
 */
int clipk_gemm_bf16(const void* A, int a_mn, int64_t lda, int64_t strideA, const void* B, int b_mn, int64_t ldb,
                    int64_t strideB, void* C, int64_t ldc, int64_t strideC, int out_dtype, int M, int N, int K,
                    int batches, float alpha, int accumulate, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* CLIPK_H_ */
