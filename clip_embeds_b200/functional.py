"""Autograd wrappers over the clipk C ABI (include/clipk.h).

The host side is PyTorch only for device memory, streams and autograd bookkeeping; all arithmetic on the hot
path runs in the sm_100a kernels of libclipk.so.  There is no CPU / eager fallback: CPU tensors raise.

Reference call sites (relative to the reference root, PACL = Patch-Aligned-Contrastive-Learning):
  patch_alignment / forward     PACL/model/pacl.py:120-145 (copies :249-275, :341-365)
  ClipLoss                      PACL/model/pacl.py:489-514
  eval scoring                  PACL/eval_pacl.py:50-57, :303-309; PACL/eval_llm2pacl.py:62-67
  open_clip ClipLoss            open_clip/src/open_clip/loss.py:89-193
"""
import os

import torch

from . import _lib

ACT = {"sigmoid": 0, "sigmoid10": 0, "ones": 1, "softmax": 2, "softmax10": 2}
_DT = {torch.bfloat16: 0, torch.float32: 1, torch.float16: 2}


def _stream():
    # raw handle of the current stream of the current device: torch.cuda.current_stream() builds a Python Stream object on
    # every call (~17 us measured; with ~17 C-ABI calls per SPARC step that was a third of the host time of a step)
    return torch._C._cuda_getCurrentRawStream(torch._C._cuda_getDevice())


def _p(t):
    return 0 if t is None else t.data_ptr()


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.ClipkError("clip_embeds_b200 runs on B200 (CUDA) tensors only; there is no CPU fallback")


def _f32(*shape, device):
    return torch.empty(*shape, dtype=torch.float32, device=device)


# --------------------------------------------------------------------------------------------- paired PACL
def _paired_forward(V, T, act, want_act=False, want_cos=False):
    _need_cuda(V, T)
    if V.dtype not in _DT:
        raise _lib.ClipkError(f"unsupported dtype {V.dtype} (bf16, fp16 or fp32)")
    T = T.to(V.dtype)
    V = V.contiguous()
    T = T.contiguous()
    Bv, P, D = V.shape
    B = T.shape[0]
    if Bv == B:
        v_div = 1
    elif B % Bv == 0:
        v_div = B // Bv          # each image scored against B / Bv consecutive captions
    else:
        raise ValueError(f"images {Bv} and texts {B}: need equal batch or one image per K captions")
    dev = V.device
    img = _f32(B, D, device=dev)
    txt = _f32(B, D, device=dev)
    stats = _f32(B, 2, device=dev)
    a = _f32(B, P, device=dev) if want_act else None
    cos = _f32(B, device=dev) if want_cos else None
    _lib.call("clipk_pacl_paired_fwd", V.data_ptr(), T.data_ptr(), _DT[V.dtype], B, v_div, P, D, act, _p(a),
              img.data_ptr(), txt.data_ptr(), _p(cos), stats.data_ptr(), _stream())
    return img, txt, stats, a, cos, V, T


class _PaclPaired(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V, T, act):
        img, txt, stats, _, _, Vc, Tc = _paired_forward(V, T, act)
        if Vc.shape[0] != Tc.shape[0]:
            ctx.mark_non_differentiable(img, txt)
        ctx.save_for_backward(Vc, Tc, img, stats)
        ctx.act = act
        ctx.t_dtype = T.dtype
        return img, txt

    @staticmethod
    def backward(ctx, d_img, d_txt):
        V, T, img, stats = ctx.saved_tensors
        B, P, D = V.shape
        d_img = d_img.float().contiguous()
        d_txt = d_txt.float().contiguous() if d_txt is not None else None
        dV = torch.empty_like(V)
        dT = torch.empty_like(T)
        _lib.call("clipk_pacl_paired_bwd", V.data_ptr(), T.data_ptr(), _DT[V.dtype], B, P, D, ctx.act, img.data_ptr(),
                  stats.data_ptr(), d_img.data_ptr(), _p(d_txt), dV.data_ptr(), dT.data_ptr(), _stream())
        return dV, dT.to(ctx.t_dtype), None


def patch_alignment(visual_patch_proj, text_cls_proj, activation="sigmoid"):
    """sigmoid(10 * cos(patch, text)) -> [B, P] fp32 (pacl.py:120-133).  Not differentiable on its own; use
    `pacl_pool` for the training path (activation + pooling + normalisation fused, one pass over V).
    activation='softmax' returns softmax_p(10 * cos) (the kernel's un-normalised exp weights, normalised here)."""
    a = _paired_forward(visual_patch_proj, text_cls_proj, ACT[activation], want_act=True)[3]
    if ACT[activation] == ACT["softmax"]:
        a = a / a.sum(dim=-1, keepdim=True)
    return a


def pacl_pool(visual_proj, text_proj, activation="sigmoid"):
    """PACL model forward after the projection heads (pacl.py:140-145 and variants): returns
    (normalised text-conditioned pooled image feature [B,D], normalised text feature [B,D]) in fp32.
    activation='ones' reproduces the checked-in "Eval only" forward (pacl.py:141-142)."""
    return _PaclPaired.apply(visual_proj, text_proj, ACT[activation])


def pacl_eval_scores(visual_proj, text_proj, c=100.0, activation="sigmoid"):
    """Eval protocol (eval_pacl.py:50-57): `visual_proj` [items,P,D], `text_proj` [items,K,D] ->
    diagonal scores [items,K] (= c * cos) and top-1 index [items] (fp32 arithmetic, no tensor cores)."""
    items, K, D = text_proj.shape
    cos = _paired_forward(visual_proj, text_proj.reshape(items * K, D), ACT[activation], want_cos=True)[4]
    scores = (c * cos).reshape(items, K)
    return scores, scores.argmax(dim=-1)


# --------------------------------------------------------------------------------------------- all-pairs PACL
def default_schedule(Bt, P, D, backward=True):
    """(group, lanes) handed to the C ABI.  group 0 = the persistent dependency-driven kernel with an automatically
    sized image group (DESIGN.md "mega kernel"); group < 0 = the same with -group images per group and `lanes` groups
    in lock-step; group > 0 = the staged path (one engine launch per GEMM, `group` images per launch on `lanes`
    internal streams)."""
    env = os.environ.get("CLIPK_AP_SCHEDULE")            # e.g. "128:2" (staged) or "-16:3" (persistent kernel)
    if env:
        parts = [int(x) for x in env.split(":")]
        return (parts[0], parts[1] if len(parts) > 1 else 1)
    return _SCHEDULE["bwd" if backward else "fwd"]


_SCHEDULE = {"fwd": (128, 2), "bwd": (128, 2)}      # staged path; the persistent kernel (group <= 0) is opt-in


def _resolve(group, Bi, Bt, P, D, backward):
    if group is None:
        g, lanes = default_schedule(Bt, P, D, backward)
    elif isinstance(group, (tuple, list)):
        g, lanes = group
    else:
        g, lanes = int(group), 1
    if g <= 0:
        return max(g, -Bi), max(0, min(lanes, 8))
    return max(1, min(g, Bi)), max(1, min(lanes, 4))


# The forward may keep the un-normalised pooled vectors u_ik (bf16 [Bi,Bt,D]) for the backward pass, which then needs
# 5 GEMM units instead of 6 (no recompute of activations + pooling).  Budget in bytes per call; above it (or with
# CLIPK_AP_SAVE_POOLED=0) the backward recomputes (flash-style, O(Bi*Bt) statistics only).
_POOLED_BUDGET = int(float(os.environ.get("CLIPK_AP_POOLED_BUDGET_GB", "48")) * (1 << 30))


def _save_pooled(Bi, Bt, D, g, needs_grad):
    if not needs_grad or g <= 0 or os.environ.get("CLIPK_AP_SAVE_POOLED", "1") == "0":
        return False
    return Bi * Bt * D * 2 <= _POOLED_BUDGET


class _PaclAllPairs(torch.autograd.Function):
    @staticmethod
    def forward(ctx, V, T, c, act, group, v_sqnorm=None):
        _need_cuda(V, T)
        Vb = V.to(torch.bfloat16).contiguous()
        Tb = T.to(torch.bfloat16).contiguous()
        Bi, P, D = Vb.shape
        Bt = Tb.shape[0]
        dev = Vb.device
        g, lanes = _resolve(group, Bi, Bt, P, D, False)
        needs_grad = any(ctx.needs_input_grad[:2])
        gb, _ = _resolve(group, Bi, Bt, P, D, True)
        save = _save_pooled(Bi, Bt, D, min(g, gb), needs_grad)
        if v_sqnorm is not None:      # squared row norms from the producer of V (the head's output GEMM): no pass over V
            rnV = v_sqnorm.detach().float().reshape(Bi, P).clamp_min(1e-24).rsqrt().contiguous()
        else:
            rnV = _f32(Bi, P, device=dev)
        rnT = _f32(Bt, device=dev)
        num, usq, scores = _f32(Bi, Bt, device=dev), _f32(Bi, Bt, device=dev), _f32(Bi, Bt, device=dev)
        pooled = torch.empty(Bi, Bt, D, dtype=torch.bfloat16, device=dev) if save else None
        nbytes = _lib.lib().clipk_pacl_allpairs_workspace_bytes(Bi, Bt, P, D, g, lanes, 2 if save else 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_pacl_allpairs_fwd", Vb.data_ptr(), Tb.data_ptr(), Bi, Bt, P, D, act, c, rnV.data_ptr(),
                  rnT.data_ptr(), num.data_ptr(), usq.data_ptr(), scores.data_ptr(), _p(pooled), ws.data_ptr(), nbytes,
                  g, lanes, 1 if v_sqnorm is not None else 0, _stream())
        if save:
            ctx.save_for_backward(Vb, Tb, rnV, rnT, num, usq, pooled)
        else:
            ctx.save_for_backward(Vb, Tb, rnV, rnT, num, usq)
        ctx.cfg = (c, act, group, V.dtype, T.dtype, save)
        return scores

    @staticmethod
    def backward(ctx, dscores):
        c, act, group, v_dtype, t_dtype, save = ctx.cfg
        if save:
            Vb, Tb, rnV, rnT, num, usq, pooled = ctx.saved_tensors
        else:
            Vb, Tb, rnV, rnT, num, usq = ctx.saved_tensors
            pooled = None
        Bi, P, D = Vb.shape
        Bt = Tb.shape[0]
        dev = Vb.device
        g, lanes = _resolve(group, Bi, Bt, P, D, True)
        dscores = dscores.float().contiguous()
        dV = torch.empty_like(Vb)
        dT = _f32(Bt, D, device=dev)
        nbytes = _lib.lib().clipk_pacl_allpairs_workspace_bytes(Bi, Bt, P, D, g, lanes, 3 if save else 1)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_pacl_allpairs_bwd", Vb.data_ptr(), Tb.data_ptr(), Bi, Bt, P, D, act, c, rnV.data_ptr(),
                  rnT.data_ptr(), num.data_ptr(), usq.data_ptr(), dscores.data_ptr(), _p(pooled), dV.data_ptr(),
                  dT.data_ptr(), ws.data_ptr(), nbytes, g, lanes, _stream())
        return dV.to(v_dtype), dT.to(t_dtype), None, None, None, None


def pacl_scores(visual_proj, text_proj, c=1.0, activation="sigmoid", group=None, v_sqnorm=None):
    """All-pairs text-conditioned scores [Bi,Bt] = c * cos(pool(V_i | t_k), t_k) (bf16 tensor cores, fp32
    accumulation).  Inputs are cast to bf16; gradients come back in the input dtypes.
    `group`: None (default schedule), images-per-group, or (images-per-group, lanes).
    `v_sqnorm` [Bi,P] (optional): squared L2 norms of the patch rows as emitted by `heads.VisualProjection(...,
    return_sqnorm=True)`; the scorer then skips its own norm pass over V."""
    return _PaclAllPairs.apply(visual_proj, text_proj, float(c), ACT[activation], group, v_sqnorm)


# --------------------------------------------------------------------------------------------- CE on a score matrix
def _k_ce_rows(L, offset):
    """row_lse [M], row_loss [M] of a fp32 matrix (label of row i = i + offset)."""
    M, N = L.shape
    row_lse, row_loss = _f32(M, device=L.device), _f32(M, device=L.device)
    _lib.call("clipk_ce_rows", L.data_ptr(), M, N, N, 0, offset, row_lse.data_ptr(), row_loss.data_ptr(), _stream())
    return row_lse, row_loss


def _k_ce_cols(L):
    """per-column (max, sum exp(L - max)) over the local rows."""
    M, N = L.shape
    col_max, col_sum = _f32(N, device=L.device), _f32(N, device=L.device)
    _lib.call("clipk_ce_cols", L.data_ptr(), M, N, N, col_max.data_ptr(), col_sum.data_ptr(), _stream())
    return col_max, col_sum


def _k_ce_payload(L, row_lse, row_loss):
    """[2N + 2] = (per-column max [N], per-column sum exp(L - max) [N], sum row_loss, sum label logits)."""
    M, N = L.shape
    payload = _f32(2 * N + 2, device=L.device)
    _lib.call("clipk_ce_cols", L.data_ptr(), M, N, N, payload.data_ptr(), payload.data_ptr() + 4 * N, _stream())
    _lib.call("clipk_ce_rowsums", row_lse.data_ptr(), row_loss.data_ptr(), M, payload.data_ptr() + 8 * N, _stream())
    return payload


def _k_ce_merge(gathered, W, N):
    """(col_lse [N], global loss [1]) from the W gathered payloads."""
    col_lse, loss = _f32(N, device=gathered.device), _f32(1, device=gathered.device)
    _lib.call("clipk_ce_merge", gathered.data_ptr(), W, N, col_lse.data_ptr(), loss.data_ptr(), _stream())
    return col_lse, loss


def _k_ce_scores_grad(L, row_lse, col_lse, offset, w_row, w_col):
    M, N = L.shape
    dL = torch.empty_like(L)
    _lib.call("clipk_ce_scores_grad", L.data_ptr(), M, N, N, row_lse.data_ptr(), col_lse.data_ptr(), offset, w_row,
              w_col, dL.data_ptr(), _stream())
    return dL


class _ScoreInfoNCE(torch.autograd.Function):
    """loss = 1/2 [ CE(L, arange) + CE(L^T, arange) ] for a score matrix L (pacl.py:509-512 applied to scores).

    Image-sharded form: this rank holds rows [offset, offset + M) of the global [N, N] matrix; column statistics
    are merged across ranks, and the returned loss is the GLOBAL loss (identical on every rank)."""

    @staticmethod
    def forward(ctx, L, offset, group):
        L = L.float().contiguous()
        M, N = L.shape
        dev = L.device
        row_lse, row_loss = _k_ce_rows(L, offset)
        # payload of this rank: (col_max [N], col_sum [N], sum row_loss, sum label logits) -- ONE all-gather carries the
        # column statistics and both loss partials
        payload = _k_ce_payload(L, row_lse, row_loss)
        W = 1
        gathered = payload
        if group is not None:
            import torch.distributed as dist
            W = dist.get_world_size(group)
            if W > 1:
                gathered = _f32(W * (2 * N + 2), device=dev)
                if L.is_cuda:
                    dist.all_gather_into_tensor(gathered, payload, group=group)
                else:
                    dist.all_gather(list(gathered.chunk(W)), payload, group=group)
        col_lse, loss = _k_ce_merge(gathered, W, N)
        dL = _k_ce_scores_grad(L, row_lse, col_lse, offset, 0.5 / N, 0.5 / N)
        ctx.save_for_backward(dL)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, g):
        (dL,) = ctx.saved_tensors
        return dL * g, None, None


def score_infonce(scores, offset=0, group=None):
    return _ScoreInfoNCE.apply(scores, int(offset), group)


# --------------------------------------------------------------------------------------------- fp32 GEMM
_SPLIT_MIN_WORK = 1 << 21          # M*N*K below this: the fp32 SIMT kernel (launch-bound either way)


def _gemm_f32(A, sam, sak, B, sbk, sbn, C, ldc, M, N, K, alpha, accumulate=False):
    """C[m,n] (+)= alpha * sum_k A[m*sam + k*sak] * B[k*sbk + n*sbn] on fp32 tensors.  Large problems run on the
    tensor cores with fp32 accuracy (bf16 x 3 split, six products folded into one tcgen05 GEMM); small ones on the
    fp32 SIMT kernel.  Both are kernels of libclipk (no library / eager path)."""
    st = _stream()
    if M * N * K >= _SPLIT_MIN_WORK:
        nbytes = _lib.lib().clipk_gemm_f32_split_workspace_bytes(M, N, K)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=C.device)
        _lib.call("clipk_gemm_f32_split", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc, M, N, K,
                  float(alpha), 1 if accumulate else 0, ws.data_ptr(), nbytes, st)
    else:
        _lib.call("clipk_sgemm_f32", A.data_ptr(), sam, sak, B.data_ptr(), sbk, sbn, C.data_ptr(), ldc, M, N, K,
                  float(alpha), 1.0 if accumulate else 0.0, st)


class _MatmulNT(torch.autograd.Function):
    """L = alpha * X @ Y.T for fp32 features (pacl.py:499: `logit_scale * image_features @ text_features.T`)."""

    @staticmethod
    def forward(ctx, X, Y, alpha):
        _need_cuda(X, Y)
        Xc, Yc = X.float().contiguous(), Y.float().contiguous()
        M, D = Xc.shape
        N = Yc.shape[0]
        L = _f32(M, N, device=Xc.device)
        _gemm_f32(Xc, D, 1, Yc, 1, D, L, N, M, N, D, alpha)
        ctx.save_for_backward(Xc, Yc)
        ctx.cfg = (float(alpha), X.dtype, Y.dtype)
        return L

    @staticmethod
    def backward(ctx, dL):
        Xc, Yc = ctx.saved_tensors
        alpha, xdt, ydt = ctx.cfg
        M, D = Xc.shape
        N = Yc.shape[0]
        dL = dL.float().contiguous()
        dX, dY = _f32(M, D, device=Xc.device), _f32(N, D, device=Xc.device)
        _gemm_f32(dL, N, 1, Yc, D, 1, dX, D, M, D, N, alpha)          # dX = alpha * dL Y
        _gemm_f32(dL, 1, N, Xc, D, 1, dY, D, N, D, M, alpha)          # dY = alpha * dL^T X
        return dX.to(xdt), dY.to(ydt), None


class _MatmulNTBf16(torch.autograd.Function):
    """L (fp32) = alpha * X @ Y.T for bf16 features on the tcgen05 engine; backward: dL is rounded to bf16 and feeds
    two engine GEMMs (fp32 gradients)."""

    @staticmethod
    def forward(ctx, X, Y, alpha):
        _need_cuda(X, Y)
        Xc, Yc = X.contiguous(), Y.contiguous()
        M, D = Xc.shape
        N = Yc.shape[0]
        L = _f32(M, N, device=Xc.device)
        _lib.call("clipk_gemm_bf16", Xc.data_ptr(), 0, D, 0, Yc.data_ptr(), 0, D, 0, L.data_ptr(), N, 0, 1, M, N, D, 1,
                  float(alpha), 0, _stream())
        ctx.save_for_backward(Xc, Yc)
        ctx.alpha = float(alpha)
        return L

    @staticmethod
    def backward(ctx, dL):
        Xc, Yc = ctx.saved_tensors
        M, D = Xc.shape
        N = Yc.shape[0]
        g = dL.to(torch.bfloat16).contiguous()
        dX, dY = _f32(M, D, device=Xc.device), _f32(N, D, device=Xc.device)
        st = _stream()
        # dX = alpha * dL Y      (A = dL [M,N] K-major, B = Y as a [K = N][D] MN-major operand)
        _lib.call("clipk_gemm_bf16", g.data_ptr(), 0, N, 0, Yc.data_ptr(), 1, D, 0, dX.data_ptr(), D, 0, 1, M, D, N, 1,
                  ctx.alpha, 0, st)
        # dY = alpha * dL^T X    (A = dL read MN-major: rows = n, k = m; B = X MN-major)
        _lib.call("clipk_gemm_bf16", g.data_ptr(), 1, N, 0, Xc.data_ptr(), 1, D, 0, dY.data_ptr(), D, 0, 1, N, D, M, 1,
                  ctx.alpha, 0, st)
        return dX.to(Xc.dtype), dY.to(Yc.dtype), None


_SYMMETRIC_DENSE_MAX = 8192        # [B,B] fp32 logits up to 256 MB; larger batches stream (logits never written)


def symmetric_infonce(X, Y, scale):
    """1/2 [CE(scale X Y^T, I) + CE(scale Y X^T, I)] (pacl.py:498-514).  fp32 features: ONE logits GEMM (the text-side
    logits are its transpose), row + column CE on the fp32 matrix, two gradient GEMMs; bf16 features: the
    tensor-core feature CE (logits never written)."""
    if X.dtype == torch.bfloat16 and Y.dtype == torch.bfloat16:
        if X.shape[0] == Y.shape[0] and X.shape[0] <= _SYMMETRIC_DENSE_MAX and X.shape[1] % 8 == 0:
            # small symmetric batch (PACL ClipLoss): one logits GEMM instead of two forward + two recompute GEMMs,
            # row and column CE on the fp32 matrix, two gradient GEMMs -- 3 tensor-core launches instead of 8
            return score_infonce(_MatmulNTBf16.apply(X, Y, float(scale)), 0, None)
        return (feat_row_ce(X, Y, scale) + feat_row_ce(Y, X, scale)) / 2
    if X.shape[0] != Y.shape[0]:
        raise ValueError(f"symmetric InfoNCE needs as many rows in X as in Y ({X.shape[0]} vs {Y.shape[0]})")
    return score_infonce(_MatmulNT.apply(X, Y, float(scale)), 0, None)


# --------------------------------------------------------------------------------------------- CE from features
class _FeatRowCE(torch.autograd.Function):
    """sum over valid rows of CE(scale * X Y^T + bias, labels) and the number of valid rows.

    bf16 inputs run on the tcgen05 engine (logits never materialised); fp32 inputs run the fp32 path (one logits GEMM at
    fp32 accuracy + row CE).  `scale` may be a python number or a device scalar tensor (open_clip hands over
    `logit_scale.exp()` as a CUDA tensor, open_clip_train/train.py:107): a tensor is read by the kernels from device
    memory -- no host copy, hence no stream stall -- and receives its gradient.  A constant `bias` added to every logit of
    a row has zero gradient under cross-entropy (softmax sums to one), which is what is returned for a tensor bias."""

    @staticmethod
    def forward(ctx, X, Y, scale, bias, labels, label_offset, slab):
        _need_cuda(X, Y)
        M, D = X.shape
        N = Y.shape[0]
        dev = X.device
        st = _stream()
        row_lse, row_loss = _f32(M, device=dev), _f32(M, device=dev)
        # hard-negative slabs of Y (see clipk.h): (int32 fill counts [W] on the device, first slab column, rows per slab)
        if slab is not None:
            slab_counts, slab_n0, slab_rows = slab
            slab_counts = slab_counts.to(device=dev, dtype=torch.int32).contiguous()
            if slab_n0 % 32 or slab_rows % 32 or slab_n0 + slab_counts.numel() * slab_rows != N:
                raise ValueError("slabs: slab_n0 and slab_rows must be multiples of 32 and cover Y's tail exactly")
        else:
            slab_counts, slab_n0, slab_rows = None, 0, 0
        lab_ptr = 0
        if labels is not None:
            labels = labels.to(device=dev, dtype=torch.int64).contiguous()
            lab_ptr = labels.data_ptr()
            valid = (labels >= 0)
        else:
            valid = torch.ones(M, dtype=torch.bool, device=dev)
        ctx.scale_needs_grad = torch.is_tensor(scale) and scale.requires_grad
        ctx.bias_is_tensor = torch.is_tensor(bias)
        scale_dev = None
        if torch.is_tensor(scale) and scale.is_cuda:
            scale_dev = scale.detach().to(torch.float32).reshape(1).contiguous()      # stays on the device
            sc = 1.0
        else:
            sc = float(scale)
        if torch.is_tensor(bias) and bias.is_cuda and bias.numel() == 1 and bias.requires_grad:
            bi = float(bias.detach())      # (rare: a learnable bias; its value is needed on the host by the fp32 GEMM call)
        else:
            bi = float(bias) if bias is not None else 0.0
        if X.dtype == torch.bfloat16 and Y.dtype == torch.bfloat16:
            Xc, Yc = X.contiguous(), Y.contiguous()
            nbytes = _lib.lib().clipk_ce_feat_bwd_workspace_bytes(M, N, D)     # shared by forward and backward
            ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            _lib.call("clipk_ce_feat_fwd", Xc.data_ptr(), Yc.data_ptr(), M, N, D, sc, bi, _p(scale_dev), _p(slab_counts),
                      slab_n0, slab_rows, lab_ptr, label_offset, row_lse.data_ptr(), row_loss.data_ptr(), ws.data_ptr(),
                      nbytes, st)
            ctx.ws = ws
            logits = None
        else:
            Xc, Yc = X.float().contiguous(), Y.float().contiguous()
            logits = _f32(M, N, device=dev)
            if scale_dev is None:
                if bi != 0.0:
                    logits.fill_(bi)
                _gemm_f32(Xc, D, 1, Yc, 1, D, logits, N, M, N, D, sc, accumulate=bi != 0.0)
            else:
                _gemm_f32(Xc, D, 1, Yc, 1, D, logits, N, M, N, D, 1.0)
                logits.mul_(scale_dev)
                if bi != 0.0:
                    logits.add_(bi)
            if slab_counts is not None:        # mask the unfilled slab rows of Y out of the softmax (device-side mask)
                col = torch.arange(N - slab_n0, device=dev)
                dead = (col % slab_rows) >= slab_counts.to(torch.int64)[col // slab_rows]
                logits[:, slab_n0:].masked_fill_(dead[None, :], float("-inf"))
            _lib.call("clipk_ce_rows", logits.data_ptr(), M, N, N, lab_ptr, label_offset, row_lse.data_ptr(),
                      row_loss.data_ptr(), st)
            ctx.ws = None
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(Xc, Yc, row_lse, labels if labels is not None else empty, valid,
                              logits if logits is not None else empty, scale_dev if scale_dev is not None else empty,
                              slab_counts if slab_counts is not None else empty)
        ctx.cfg = (sc, bi, labels is not None, label_offset, X.dtype, Y.dtype, scale_dev is not None,
                   (slab_n0, slab_rows) if slab_counts is not None else None,
                   scale.shape if torch.is_tensor(scale) else None)
        loss_sum = (row_loss * valid).sum()
        ctx.mark_non_differentiable(valid)
        return loss_sum, valid

    @staticmethod
    def backward(ctx, g_sum, _g_valid):
        Xc, Yc, row_lse, labels, valid, logits, scale_dev, slab_counts = ctx.saved_tensors
        sc, bi, has_labels, label_offset, xdt, ydt, dev_scale, slab_geo, scale_shape = ctx.cfg
        sl_ptr = slab_counts.data_ptr() if slab_geo is not None else 0
        slab_n0, slab_rows = slab_geo if slab_geo is not None else (0, 0)
        M, D = Xc.shape
        N = Yc.shape[0]
        dev = Xc.device
        st = _stream()
        lab_ptr = labels.data_ptr() if has_labels else 0
        sd_ptr = scale_dev.data_ptr() if dev_scale else 0
        row_w = (valid.float() * g_sum.float()).contiguous()
        dbias = torch.zeros((), device=dev) if ctx.bias_is_tensor else None
        if (Xc.dtype == torch.bfloat16 and xdt == torch.bfloat16 and ydt == torch.bfloat16 and 128 < M <= 4096 and N > 128
                and not ctx.scale_needs_grad):
            # bf16 gradients straight from the GEMM epilogues (no fp32 round trip + cast pass)
            dXb = torch.empty(M, D, dtype=torch.bfloat16, device=dev)
            dYb = torch.empty(N, D, dtype=torch.bfloat16, device=dev)
            ws = ctx.ws
            _lib.call("clipk_ce_feat_bwd_bf16", Xc.data_ptr(), Yc.data_ptr(), M, N, D, sc, bi, sd_ptr, sl_ptr, slab_n0,
                      slab_rows, lab_ptr, label_offset, row_lse.data_ptr(), row_w.data_ptr(), dXb.data_ptr(),
                      dYb.data_ptr(), ws.data_ptr(), ws.numel(), st)
            return dXb, dYb, None, dbias, None, None, None
        dX, dY = _f32(M, D, device=dev), _f32(N, D, device=dev)
        if Xc.dtype == torch.bfloat16:
            ws = ctx.ws
            _lib.call("clipk_ce_feat_bwd", Xc.data_ptr(), Yc.data_ptr(), M, N, D, sc, bi, sd_ptr, sl_ptr, slab_n0,
                      slab_rows, lab_ptr, label_offset, row_lse.data_ptr(), row_w.data_ptr(), dX.data_ptr(), 0,
                      dY.data_ptr(), 0, ws.data_ptr(), ws.numel(), st)
        else:
            dL = torch.empty_like(logits)
            _lib.call("clipk_ce_rows_grad", logits.data_ptr(), M, N, N, row_lse.data_ptr(), lab_ptr, label_offset,
                      row_w.data_ptr(), dL.data_ptr(), st)
            # dX = scale * dL Y ; dY = scale * dL^T X
            _gemm_f32(dL, N, 1, Yc, D, 1, dX, D, M, D, N, sc)
            _gemm_f32(dL, 1, N, Xc, D, 1, dY, D, N, D, M, sc)
            if dev_scale:
                dX.mul_(scale_dev)
                dY.mul_(scale_dev)
        dscale = None
        if ctx.scale_needs_grad:      # d/dscale sum(dlogits * x.y) = <dX, X> / scale
            dscale = ((dX * Xc.float()).sum() / (scale_dev.reshape(()) if dev_scale else sc)).reshape(scale_shape)
        return dX.to(xdt), dY.to(ydt), dscale, dbias, None, None, None


class _SymFeatCE(torch.autograd.Function):
    """This rank's share of the symmetric InfoNCE from ONE logits GEMM (open_clip local-loss semantics, loss.py:137-170):

        loss = 1/2 [ mean_i CE(scale X_i Y^T, off + i)  +  mean_i CE(scale Y_{off+i} Xall^T, off + i) ]

    X [b,D] are this rank's images, Y [Nall,D] every caption ([N0 = W b originals | hard-negative slabs]).  The second
    term is a COLUMN cross-entropy of the same logits over the image rows of all ranks: every rank adds up
    exp(logit - ref) per column over its own rows, one all-reduce makes the sums global.  The backward writes
    d(sum of all ranks' losses) / dX (complete: every column term of these rows is known locally) and this rank's partial
    of dY (summed over the ranks by the gather's reduce-scatter) -- the image features are never gathered.
    Assumes the same upstream gradient and the same b on every rank (true for a loss that is back-propagated as is)."""

    @staticmethod
    def forward(ctx, X, Y, scale, bias, off, n0, slab, group):
        _need_cuda(X, Y)
        M, D = X.shape
        N = Y.shape[0]
        dev = X.device
        st = _stream()
        Xc, Yc = X.contiguous(), Y.contiguous()
        if slab is not None:
            slab_counts, slab_n0, slab_rows = slab
            slab_counts = slab_counts.to(device=dev, dtype=torch.int32).contiguous()
        else:
            slab_counts, slab_n0, slab_rows = None, 0, 0
        scale_dev = None
        if torch.is_tensor(scale) and scale.is_cuda:
            scale_dev = scale.detach().to(torch.float32).reshape(1).contiguous()
            sc = 1.0
        else:
            sc = float(scale)
        bi = float(bias) if bias is not None else 0.0
        row_lse, row_loss, col_sum = _f32(M, device=dev), _f32(M, device=dev), _f32(n0, device=dev)
        nbytes = _lib.lib().clipk_ce_feat_bwd_workspace_bytes(M, N, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_ce_sym_fwd", Xc.data_ptr(), Yc.data_ptr(), M, N, D, sc, bi, _p(scale_dev), _p(slab_counts), slab_n0,
                  slab_rows, off, n0, row_lse.data_ptr(), row_loss.data_ptr(), col_sum.data_ptr(), ws.data_ptr(), nbytes, st)
        if group is not None:
            import torch.distributed as dist
            if dist.get_world_size(group) > 1:
                dist.all_reduce(col_sum, op=dist.ReduceOp.SUM, group=group)
        s_t = scale_dev.reshape(()) if scale_dev is not None else torch.tensor(sc, device=dev)
        col_lse = 0.5 * s_t + bi + torch.log(col_sum)                       # [n0] over the image rows of ALL ranks
        pos = row_lse - row_loss                                            # logit of pair (i, off + i)
        loss = 0.5 * (row_loss.mean() + (col_lse[off:off + M] - pos).mean())
        ctx.save_for_backward(Xc, Yc, row_lse, col_lse, scale_dev if scale_dev is not None else torch.empty(0, device=dev),
                              slab_counts if slab_counts is not None else torch.empty(0, device=dev))
        ctx.cfg = (sc, bi, off, n0, (slab_n0, slab_rows) if slab_counts is not None else None, scale_dev is not None,
                   torch.is_tensor(scale) and scale.requires_grad, scale.shape if torch.is_tensor(scale) else None, X.dtype,
                   Y.dtype, torch.is_tensor(bias))
        ctx.ws = ws
        return loss

    @staticmethod
    def backward(ctx, g):
        Xc, Yc, row_lse, col_lse, scale_dev, slab_counts = ctx.saved_tensors
        sc, bi, off, n0, slab_geo, dev_scale, scale_grad, scale_shape, xdt, ydt, bias_tensor = ctx.cfg
        M, D = Xc.shape
        N = Yc.shape[0]
        dev = Xc.device
        w = (g.float() * (0.5 / M)).reshape(1)
        row_w = w.expand(M).contiguous()
        col_w = w.expand(n0).contiguous()                                   # every rank weighs its own b captions alike
        col_bias = (torch.log(col_w) - col_lse).contiguous()
        sl_ptr = slab_counts.data_ptr() if slab_geo is not None else 0
        slab_n0, slab_rows = slab_geo if slab_geo is not None else (0, 0)
        bf16_out = M <= 4096 and not scale_grad and M > 128 and N > 128
        dX = torch.empty(M, D, dtype=torch.bfloat16 if bf16_out else torch.float32, device=dev)
        dY = torch.empty(N, D, dtype=torch.bfloat16 if bf16_out else torch.float32, device=dev)
        ws = ctx.ws
        _lib.call("clipk_ce_sym_bwd", Xc.data_ptr(), Yc.data_ptr(), M, N, D, sc, bi, scale_dev.data_ptr() if dev_scale else 0,
                  sl_ptr, slab_n0, slab_rows, off, n0, row_lse.data_ptr(), row_w.data_ptr(), col_bias.data_ptr(),
                  col_w.data_ptr(), dX.data_ptr(), dY.data_ptr(), 1 if bf16_out else 0, 0, ws.data_ptr(), ws.numel(), _stream())
        dscale = None
        if scale_grad:                 # d/dscale of sum(dlogits * x.y) = <dX, X> / scale
            dscale = ((dX.float() * Xc.float()).sum() / (scale_dev.reshape(()) if dev_scale else sc)).reshape(scale_shape)
        dbias = torch.zeros((), device=dev) if bias_tensor else None
        return dX.to(xdt), dY.to(ydt), dscale, dbias, None, None, None, None


class _SymFeatCEDist(torch.autograd.Function):
    """_SymFeatCE with the caption exchange inside (open_clip local loss, `gather_with_grad`, optional `usehardtext`):
    forward all-gathers this rank's captions (fixed-capacity slabs + device-side fill counts when there are hard
    negatives), backward reduce-scatters the caption gradient ITSELF -- started right after the dY GEMM so that it runs
    over NVLink while the dX GEMM is still on the tensor cores (two-phase clipk_ce_sym_bwd)."""

    @staticmethod
    def forward(ctx, X, txt, scale, bias, rank, world, b, usehardtext, group):
        import torch.distributed as dist
        _need_cuda(X, txt)
        dev = X.device
        D = X.shape[1]
        Xc, tc = X.contiguous(), txt.contiguous()
        h = tc.shape[0] - b
        rows = 2 * b if usehardtext else b                  # rows every rank contributes to the gather
        if tc.shape[0] < rows:
            tc_pad = torch.cat([tc, tc.new_zeros((rows - tc.shape[0], D))], dim=0)
        else:
            tc_pad = tc
        gathered = torch.empty(world * rows, D, dtype=tc.dtype, device=dev)
        dist.all_gather_into_tensor(gathered, tc_pad, group=group)
        n0 = world * b
        if usehardtext:
            counts = torch.empty(world, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(counts, torch.tensor([h], dtype=torch.int32, device=dev), group=group)
            g3 = gathered.reshape(world, rows, D)
            Y = torch.cat([g3[:, :b].reshape(n0, D), g3[:, b:].reshape(n0, D)], dim=0)       # [originals | slabs]
            slab_counts, slab_n0, slab_rows = counts, n0, b
        else:
            Y = gathered
            slab_counts, slab_n0, slab_rows = None, 0, 0
        M, N = Xc.shape[0], Y.shape[0]
        off = b * rank
        scale_dev = None
        if torch.is_tensor(scale) and scale.is_cuda:
            scale_dev = scale.detach().to(torch.float32).reshape(1).contiguous()
            sc = 1.0
        else:
            sc = float(scale)
        bi = float(bias) if bias is not None else 0.0
        st = _stream()
        row_lse, row_loss, col_sum = _f32(M, device=dev), _f32(M, device=dev), _f32(n0, device=dev)
        nbytes = _lib.lib().clipk_ce_feat_bwd_workspace_bytes(M, N, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_ce_sym_fwd", Xc.data_ptr(), Y.data_ptr(), M, N, D, sc, bi, _p(scale_dev), _p(slab_counts), slab_n0,
                  slab_rows, off, n0, row_lse.data_ptr(), row_loss.data_ptr(), col_sum.data_ptr(), ws.data_ptr(), nbytes, st)
        dist.all_reduce(col_sum, op=dist.ReduceOp.SUM, group=group)
        s_t = scale_dev.reshape(()) if scale_dev is not None else torch.tensor(sc, device=dev)
        col_lse = 0.5 * s_t + bi + torch.log(col_sum)
        pos = row_lse - row_loss
        loss = 0.5 * (row_loss.mean() + (col_lse[off:off + M] - pos).mean())
        empty = torch.empty(0, device=dev)
        ctx.save_for_backward(Xc, Y, row_lse, col_lse, scale_dev if scale_dev is not None else empty,
                              slab_counts if slab_counts is not None else empty)
        ctx.cfg = (sc, bi, off, n0, (slab_n0, slab_rows) if slab_counts is not None else None, scale_dev is not None,
                   torch.is_tensor(scale) and scale.requires_grad, scale.shape if torch.is_tensor(scale) else None, X.dtype,
                   txt.dtype, torch.is_tensor(bias), world, b, rows, tc.shape[0], bool(usehardtext))
        ctx.ws = ws
        ctx.group = group
        return loss

    @staticmethod
    def backward(ctx, g):
        import torch.distributed as dist
        Xc, Y, row_lse, col_lse, scale_dev, slab_counts = ctx.saved_tensors
        (sc, bi, off, n0, slab_geo, dev_scale, scale_grad, scale_shape, xdt, tdt, bias_tensor, world, b, rows, ntxt,
         hard) = ctx.cfg
        M, D = Xc.shape
        N = Y.shape[0]
        dev = Xc.device
        w = (g.float() * (0.5 / M)).reshape(1)
        row_w = w.expand(M).contiguous()
        col_w = w.expand(n0).contiguous()
        col_bias = (torch.log(col_w) - col_lse).contiguous()
        sl_ptr = slab_counts.data_ptr() if slab_geo is not None else 0
        slab_n0, slab_rows = slab_geo if slab_geo is not None else (0, 0)
        two_phase = M <= 4096 and M > 128 and N > 128
        bf16_out = two_phase and not scale_grad
        gdt = torch.bfloat16 if bf16_out else torch.float32
        dX = torch.empty(M, D, dtype=gdt, device=dev)
        dY = torch.empty(N, D, dtype=gdt, device=dev)
        ws = ctx.ws

        def run(phase):
            _lib.call("clipk_ce_sym_bwd", Xc.data_ptr(), Y.data_ptr(), M, N, D, sc, bi, scale_dev.data_ptr() if dev_scale else 0,
                      sl_ptr, slab_n0, slab_rows, off, n0, row_lse.data_ptr(), row_w.data_ptr(), col_bias.data_ptr(),
                      col_w.data_ptr(), dX.data_ptr(), dY.data_ptr(), 1 if bf16_out else 0, phase, ws.data_ptr(), ws.numel(),
                      _stream())

        run(1 if two_phase else 0)
        # caption gradient back to its owners: [originals | slabs] -> [W][rows][D], reduce-scatter(SUM)
        if hard:
            send = torch.cat([dY[:n0].reshape(world, b, D), dY[n0:].reshape(world, b, D)], dim=1).contiguous()
        else:
            send = dY.reshape(world, rows, D)
        d_txt_pad = torch.empty(rows, D, dtype=gdt, device=dev)
        work = dist.reduce_scatter_tensor(d_txt_pad, send.reshape(world * rows, D), op=dist.ReduceOp.SUM, group=ctx.group,
                                          async_op=True)
        if two_phase:
            run(2)                     # the dX GEMM runs while the reduce-scatter is on the wire
        work.wait()
        dscale = None
        if scale_grad:
            dscale = ((dX.float() * Xc.float()).sum() / (scale_dev.reshape(()) if dev_scale else sc)).reshape(scale_shape)
        dbias = torch.zeros((), device=dev) if bias_tensor else None
        return dX.to(xdt), d_txt_pad[:ntxt].to(tdt), dscale, dbias, None, None, None, None, None


def sym_feat_ce_dist(X, text_local, scale, bias, rank, world, b, usehardtext, group):
    """See _SymFeatCEDist: this rank's local-loss share with the caption gather / gradient reduce-scatter inside."""
    return _SymFeatCEDist.apply(X, text_local, scale, bias, int(rank), int(world), int(b), bool(usehardtext), group)


def sym_feat_ce(X, Y_all, scale, bias=0.0, offset=0, n_orig=None, slab=None, group=None):
    """See _SymFeatCE.  bf16 features with more than 128 rows / columns (CTA-pair engine); `n_orig` = number of original
    captions at the head of Y_all (default: all of them)."""
    n0 = Y_all.shape[0] if n_orig is None else int(n_orig)
    return _SymFeatCE.apply(X, Y_all, scale, bias, int(offset), n0, slab, group)


def sym_feat_ce_ok(X, Y):
    return (X.is_cuda and X.dtype == torch.bfloat16 and Y.dtype == torch.bfloat16 and X.shape[0] > 128 and Y.shape[0] > 128
            and X.shape[1] % 8 == 0 and os.environ.get("CLIPK_CE_SYM", "1") != "0")


def feat_row_ce(X, Y, scale, bias=0.0, labels=None, label_offset=0, slab=None):
    """Mean over non-ignored rows of cross_entropy(scale * X @ Y.T + bias, labels)  (F.cross_entropy semantics,
    ignore_index = any negative label).  labels=None means label_i = i + label_offset.
    slab = (counts int32 [W], first slab column, rows per slab): Y's tail holds W fixed-capacity slabs of hard negatives,
    the unfilled rows are masked out of the softmax on the device."""
    loss_sum, valid = _FeatRowCE.apply(X, Y, scale, bias, labels, int(label_offset), slab)
    return loss_sum / valid.sum().clamp_min(1)


# --------------------------------------------------------------------------------------------- SPARC
class _SparcAlign(torch.autograd.Function):
    """sparc.forward alignment (pacl.py:453-478): (V raw [B,P,D], L raw [B,T,D]) -> (n(L), n(G)) fp32."""

    @staticmethod
    def forward(ctx, V, L, sigma):
        _need_cuda(V, L)
        Vb = V.to(torch.bfloat16).contiguous()
        Lb = L.to(torch.bfloat16).contiguous()
        B, P, D = Vb.shape
        T = Lb.shape[1]
        dev = Vb.device
        l_hat, g_hat = _f32(B, T, D, device=dev), _f32(B, T, D, device=dev)
        lnorm, gnorm = _f32(B, T, device=dev), _f32(B, T, device=dev)
        nbytes = _lib.lib().clipk_sparc_workspace_bytes(B, T, P, D, 0)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        # `pooled` = mean over patches of the raw V (SparcLoss's global image feature, pacl.py:561): produced by the same
        # pass, so that its gradient -- a [B,D] vector broadcast over all patches -- is added inside the dV kernel (g_add)
        # instead of being materialised as a second dense [B,P,D] gradient and summed by autograd
        pooled = _f32(B, D, device=dev)
        _lib.call("clipk_sparc_align_fwd", Vb.data_ptr(), Lb.data_ptr(), B, T, P, D, float(sigma), l_hat.data_ptr(),
                  g_hat.data_ptr(), lnorm.data_ptr(), gnorm.data_ptr(), pooled.data_ptr(), ws.data_ptr(), nbytes, _stream())
        ctx.save_for_backward(Vb, Lb, l_hat, g_hat, lnorm, gnorm)
        ctx.cfg = (float(sigma), V.dtype, L.dtype)
        return l_hat, g_hat, pooled

    @staticmethod
    def backward(ctx, d_l_hat, d_g_hat, d_pooled):
        Vb, Lb, l_hat, g_hat, lnorm, gnorm = ctx.saved_tensors
        sigma, v_dtype, l_dtype = ctx.cfg
        B, P, D = Vb.shape
        T = Lb.shape[1]
        dev = Vb.device
        d_l_hat = (torch.zeros_like(l_hat) if d_l_hat is None else d_l_hat.float()).contiguous()
        d_g_hat = (torch.zeros_like(g_hat) if d_g_hat is None else d_g_hat.float()).contiguous()
        out_bf16 = v_dtype == torch.bfloat16
        dV = torch.empty(B, P, D, dtype=torch.bfloat16 if out_bf16 else torch.float32, device=dev)
        dL = _f32(B, T, D, device=dev)
        nbytes = _lib.lib().clipk_sparc_workspace_bytes(B, T, P, D, 1)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        g_add = None if d_pooled is None else (d_pooled.float() / P).contiguous()
        _lib.call("clipk_sparc_align_bwd", Vb.data_ptr(), Lb.data_ptr(), B, T, P, D, sigma, l_hat.data_ptr(),
                  g_hat.data_ptr(), lnorm.data_ptr(), gnorm.data_ptr(), d_g_hat.data_ptr(), d_l_hat.data_ptr(),
                  0 if g_add is None else g_add.data_ptr(), dV.data_ptr(), 1 if out_bf16 else 0, dL.data_ptr(),
                  ws.data_ptr(), nbytes, _stream())
        return dV.to(v_dtype), dL.to(l_dtype), None


def sparc_align(v_patch_embed, l_token_embed, sigma):
    """Returns (l_token_embed_normalised, l_grouped_v_patch_embed_normalised), both fp32 [B,T,D].

    Also leaves the patch-mean of `v_patch_embed` on that tensor object (`_clipk_pooled`): `SparcLoss` picks it up
    when it is handed the same tensor (the reference's `sparc.forward` returns `v_patch_embed` itself, pacl.py:478),
    so the gradient of the global term's mean-pool is fused into the alignment backward."""
    l_hat, g_hat, pooled = _SparcAlign.apply(v_patch_embed, l_token_embed, sigma)
    try:
        v_patch_embed._clipk_pooled = (pooled, v_patch_embed._version)
    except Exception:      # exotic tensor subclasses without a __dict__: SparcLoss falls back to mean_dim1
        pass
    return l_hat, g_hat


def pooled_patch_mean(v_patch_embed):
    """mean over patches [B,D] fp32: the value `sparc_align` left on this tensor, else a fresh `mean_dim1`."""
    tagged = getattr(v_patch_embed, "_clipk_pooled", None)
    if tagged is not None:
        del v_patch_embed._clipk_pooled                  # one use: its autograd graph is consumed by one backward
        pooled, version = tagged
        if version == v_patch_embed._version and pooled.device == v_patch_embed.device:
            return pooled
    return mean_dim1(v_patch_embed)


class _MeanDim1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        _need_cuda(X)
        if X.dtype not in _DT:
            X = X.float()
        Xc = X.contiguous()
        B, R, D = Xc.shape
        out = _f32(B, D, device=Xc.device)
        _lib.call("clipk_mean_dim1", Xc.data_ptr(), _DT[Xc.dtype], B, R, D, out.data_ptr(), _stream())
        ctx.cfg = (R, X.dtype)
        return out

    @staticmethod
    def backward(ctx, g):
        R, dt = ctx.cfg
        return (g / R).to(dt).unsqueeze(1).expand(-1, R, -1)


def mean_dim1(X):
    return _MeanDim1.apply(X)


class _NormalizeRows(torch.autograd.Function):
    @staticmethod
    def forward(ctx, X):
        _need_cuda(X)
        Xc = X.float().contiguous()
        D = Xc.shape[-1]
        rows = Xc.numel() // D
        out = torch.empty_like(Xc)
        norm = _f32(rows, device=Xc.device)
        _lib.call("clipk_normalize_rows_fwd", Xc.data_ptr(), rows, D, out.data_ptr(), norm.data_ptr(), _stream())
        ctx.save_for_backward(out, norm)
        ctx.dt = X.dtype
        return out

    @staticmethod
    def backward(ctx, g):
        out, norm = ctx.saved_tensors
        D = out.shape[-1]
        g = g.float().contiguous()
        dx = torch.empty_like(out)
        _lib.call("clipk_normalize_rows_bwd", out.data_ptr(), g.data_ptr(), norm.data_ptr(), out.numel() // D, D,
                  dx.data_ptr(), _stream())
        return dx.to(ctx.dt)


def normalize_rows(X):
    """F.normalize(X, dim=-1) in fp32."""
    return _NormalizeRows.apply(X)


class _SparcLocal(torch.autograd.Function):
    """1/2 [ masked_pairwise(g_hat, l_hat) + masked_pairwise(l_hat, g_hat) ]  (pacl.py:522-556, :575-582).
    `mask_sum` may be a GLOBAL mask count (image-sharded training): the returned value is then this rank's
    contribution sum_b(...) / (2 * mask_sum)."""

    @staticmethod
    def forward(ctx, g_hat, l_hat, mask, scale, mask_sum):
        _need_cuda(g_hat, l_hat, mask)
        a = g_hat.float().contiguous()
        b = l_hat.float().contiguous()
        m = mask.float().contiguous()
        B, T, D = a.shape
        dev = a.device
        loss_sum = _f32(B, device=dev)
        nbytes = _lib.lib().clipk_sparc_local_workspace_bytes(B, T, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_sparc_local_fwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), B, T, D, float(scale),
                  loss_sum.data_ptr(), ws.data_ptr(), nbytes, _stream())
        msum = m.sum() if mask_sum is None else mask_sum.float()
        ctx.save_for_backward(a, b, m, msum)
        ctx.cfg = (float(scale), g_hat.dtype, l_hat.dtype)
        return loss_sum.sum() * 0.5 / msum

    @staticmethod
    def backward(ctx, g):
        a, b, m, msum = ctx.saved_tensors
        scale, adt, bdt = ctx.cfg
        B, T, D = a.shape
        dev = a.device
        wgt = (g.float() * 0.5 / msum).reshape(1).contiguous()
        d_a, d_b = torch.empty_like(a), torch.empty_like(b)
        scratch = _f32(B, device=dev)
        nbytes = _lib.lib().clipk_sparc_local_workspace_bytes(B, T, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_sparc_local_bwd", a.data_ptr(), b.data_ptr(), m.data_ptr(), B, T, D, scale, wgt.data_ptr(),
                  d_a.data_ptr(), d_b.data_ptr(), scratch.data_ptr(), ws.data_ptr(), nbytes, _stream())
        return d_a.to(adt), d_b.to(bdt), None, None, None


def sparc_local_loss(g_hat, l_hat, mask, scale, mask_sum=None):
    return _SparcLocal.apply(g_hat, l_hat, mask, scale, mask_sum)
