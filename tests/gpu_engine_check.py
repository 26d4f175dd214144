"""Standalone diagnostic for the tcgen05 engine (run on the GPU box; prints per-case errors).

    python tests/gpu_engine_check.py [case ...]

Each case runs C = A @ B^T through clipk_gemm_bf16 and compares with torch.matmul (fp32, same bf16 inputs).
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200 import _lib  # noqa: E402

CASES = {
    # name: (M, N, K, batches, a_mn, b_mn, out_bf16, bcastA, accumulate)
    "kk_small": (128, 256, 64, 1, 0, 0, 0, 0, 0),
    "kk": (256, 512, 256, 1, 0, 0, 0, 0, 0),
    "kk_big": (2048, 2048, 768, 1, 0, 0, 0, 0, 0),
    "kk_tail": (200, 300, 200, 1, 0, 0, 0, 0, 0),
    "kk_batched": (256, 576, 768, 5, 0, 0, 1, 1, 0),
    "kk_n64": (128, 64, 128, 1, 0, 0, 0, 0, 0),
    "kk_n128": (256, 128, 128, 2, 0, 0, 0, 0, 1),
    "kk_n192": (256, 192, 192, 2, 0, 0, 0, 0, 0),
    "k_mn": (256, 512, 256, 1, 0, 1, 0, 0, 0),
    "mn_k": (256, 512, 256, 1, 1, 0, 0, 0, 0),
    "mn_mn": (256, 512, 256, 1, 1, 1, 0, 0, 0),
    "k_mn_batched_tail": (200, 768, 576, 3, 0, 1, 1, 0, 0),
    "mn_mn_batched_tail": (200, 768, 1000, 3, 1, 1, 0, 0, 0),
}


def run(name):
    M, N, K, nb, a_mn, b_mn, out_bf16, bcastA, acc = CASES[name]
    g = torch.Generator(device="cpu").manual_seed(sum(map(ord, name)))
    nA = 1 if bcastA else nb
    A = torch.randn(nA, M, K, generator=g).to(torch.bfloat16).cuda()
    B = torch.randn(nb, N, K, generator=g).to(torch.bfloat16).cuda()
    ref = torch.matmul(A.float().expand(nb, M, K), B.float().transpose(1, 2))
    Ad = A.transpose(1, 2).contiguous() if a_mn else A       # [nb][K][M] when MN-major
    Bd = B.transpose(1, 2).contiguous() if b_mn else B
    lda = M if a_mn else K
    ldb = N if b_mn else K
    if out_bf16:
        C = torch.zeros(nb, M, N, dtype=torch.bfloat16, device="cuda")
    else:
        C = torch.full((nb, M, N), 1.0 if acc else 0.0, dtype=torch.float32, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    _lib.call("clipk_gemm_bf16", Ad.data_ptr(), a_mn, lda, 0 if bcastA else M * K, Bd.data_ptr(), b_mn, ldb, N * K,
              C.data_ptr(), N, M * N, 0 if out_bf16 else 1, M, N, K, nb, 1.0, acc, st)
    torch.cuda.synchronize()
    if acc:
        ref = ref + 1.0
    err = (C.float() - ref).abs().max().item()
    scale = ref.abs().max().item()
    tol = (2e-2 if out_bf16 else 2e-3) * scale
    bad = (C.float() - ref).abs() > tol
    print(f"{name}: max_abs_err={err:.4e} ref_max={scale:.3e} bad={int(bad.sum())}/{bad.numel()} "
          f"{'OK' if err <= tol else 'FAIL'}", flush=True)
    if err > tol:
        idx = bad.nonzero()[:5].tolist()
        print("   first bad idx:", idx, "got", [C[tuple(i)].item() for i in idx], "ref", [ref[tuple(i)].item() for i in idx])
    return err <= tol


if __name__ == "__main__":
    names = sys.argv[1:] or list(CASES)
    ok = all([run(n) for n in names])
    sys.exit(0 if ok else 1)
