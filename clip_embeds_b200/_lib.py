"""ctypes binding of libclipk.so (the C ABI declared in include/clipk.h).

There is NO fallback: if the shared library is missing or a call fails, an exception is raised.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libclipk.so")

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t

# name -> (restype, argtypes); must list EVERY symbol include/clipk.h declares (tests/test_abi.py checks it).
SIGNATURES = {
    "clipk_last_error": (ctypes.c_char_p, []),
    "clipk_version": (c_int, []),
    "clipk_launch_count": (ctypes.c_ulonglong, []),
    "clipk_check_device": (c_int, []),
    "clipk_trace_dump": (c_int, []),
    "clipk_gemm_bf16": (c_int, [c_vp, c_int, c_i64, c_i64, c_vp, c_int, c_i64, c_i64, c_vp, c_i64, c_i64, c_int,
                                c_int, c_int, c_int, c_int, c_f32, c_int, c_vp]),
    "clipk_pacl_paired_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "clipk_pacl_paired_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp]),
    "clipk_pacl_allpairs_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int, c_int, c_int, c_int]),
    "clipk_pacl_allpairs_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_sz, c_int, c_int, c_int, c_vp]),
    "clipk_pacl_allpairs_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp,
                                        c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_int, c_int, c_vp]),
    "clipk_ce_rows": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "clipk_ce_cols": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_vp]),
    "clipk_ce_rowsums": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp]),
    "clipk_ce_merge": (c_int, [c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "clipk_ce_scores_grad": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_f32, c_f32, c_vp, c_vp]),
    "clipk_ce_rows_grad": (c_int, [c_vp, c_int, c_int, c_i64, c_vp, c_vp, c_i64, c_vp, c_vp, c_vp]),
    "clipk_sgemm_f32": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_f32, c_f32,
                                c_vp]),
    "clipk_gemm_f32_split_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "clipk_gemm_f32_split": (c_int, [c_vp, c_i64, c_i64, c_vp, c_i64, c_i64, c_vp, c_i64, c_int, c_int, c_int, c_f32,
                                     c_int, c_vp, c_sz, c_vp]),
    "clipk_ce_feat_workspace_bytes": (c_sz, [c_int, c_int]),
    "clipk_ce_feat_bwd_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "clipk_ce_feat_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_sz,
                                  c_vp]),
    "clipk_ce_feat_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_int,
                                  c_vp, c_int, c_vp, c_sz, c_vp]),
    "clipk_ce_sym_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_vp,
                                 c_vp, c_vp, c_vp, c_sz, c_vp]),
    "clipk_ce_sym_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_int, c_int, c_i64, c_int, c_vp,
                                 c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_sz, c_vp]),
    "clipk_ce_feat_bwd_bf16": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_f32, c_f32, c_vp, c_vp, c_int, c_int, c_vp, c_i64, c_vp, c_vp, c_vp, c_vp,
                                       c_vp, c_sz, c_vp]),
    "clipk_sparc_workspace_bytes": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "clipk_sparc_align_fwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_sz, c_vp]),
    "clipk_sparc_align_bwd": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                      c_vp, c_vp, c_int, c_vp, c_vp, c_sz, c_vp]),
    "clipk_sparc_local_workspace_bytes": (c_sz, [c_int, c_int, c_int]),
    "clipk_sparc_local_fwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_sz, c_vp]),
    "clipk_sparc_local_bwd": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f32, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz,
                                      c_vp]),
    "clipk_mean_dim1": (c_int, [c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp]),
    "clipk_normalize_rows_fwd": (c_int, [c_vp, c_i64, c_int, c_vp, c_vp, c_vp]),
    "clipk_normalize_rows_bwd": (c_int, [c_vp, c_vp, c_vp, c_i64, c_int, c_vp, c_vp]),
    "clipk_ln_fwd": (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_f32, c_vp, c_vp, c_vp, c_f32, ctypes.c_uint64, c_vp,
                             c_vp]),
    "clipk_ln_bwd_workspace_bytes": (c_sz, [c_i64, c_int]),
    "clipk_ln_bwd": (c_int, [c_vp, c_int, c_i64, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_f32, c_vp, c_vp, c_sz,
                             c_vp]),
    "clipk_patch_proj_fwd": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp]),
    "clipk_patch_proj_bwd_workspace_bytes": (c_sz, [c_i64, c_int, c_int]),
    "clipk_patch_proj_bwd": (c_int, [c_vp, c_vp, c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp,
                                     c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "clipk_linear_fwd": (c_int, [c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "clipk_linear_bwd_workspace_bytes": (c_sz, [c_i64, c_int, c_int]),
    "clipk_linear_bwd": (c_int, [c_vp, c_vp, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_vp, c_vp, c_sz, c_vp]),
    "clipk_eval_correct": (c_int, [c_vp, c_int, c_int, c_vp, c_vp]),
    "clipk_eval_whatsup": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp, c_vp]),
    "clipk_eval_mmvp": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp]),
    "clipk_rope": (c_int, [c_vp, c_int, c_i64, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_vp]),
    "clipk_retrieval_ranks": (c_int, [c_vp, c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_vp]),
    "clipk_rank_stats": (c_int, [c_vp, c_int, c_vp, c_vp]),
}

_lib = None


class ClipkError(RuntimeError):
    pass


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClipkError(
                f"{LIB_PATH} not found: build it with `python -m clip_embeds_b200.build` "
                "(there is no CPU / eager fallback for this path)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc):
    if rc != 0:
        msg = lib().clipk_last_error()
        raise ClipkError(f"clipk error {rc}: {msg.decode() if msg else '?'}")


_NVTX = os.environ.get("CLIPK_NVTX", "0") != "0"      # NVTX range around every C-ABI call (timelines; off by default)


def call(name, *args):
    if _NVTX:
        import torch
        torch.cuda.nvtx.range_push(name)
        try:
            check(getattr(lib(), name)(*args))
        finally:
            torch.cuda.nvtx.range_pop()
        return
    check(getattr(lib(), name)(*args))
