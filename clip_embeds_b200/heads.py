"""Projection heads of the PACL models: the producer of the patch tensor the scorer reads.

  Patch_Projection(in_dim, out_dim)                                   PACL/model/pacl.py:35-48
  visual_projection = Sequential(LayerNorm, Dropout(0.1), Patch_Projection)      pacl.py:70-74 (ViT-L/14-336: 1024 -> 768)
  text_projection   = Sequential(LayerNorm, Dropout(0.1), Linear)                pacl.py:75-79

The modules keep the reference's constructor arguments, sub-module names and state-dict keys (including the
`linear_projection` / `text_projection` alias of ONE Linear, pacl.py:39), so a reference checkpoint loads with
`load_state_dict`.  Parameters stay fp32 masters; every call casts the (small) weight matrices to bf16 and runs
LayerNorm, the GEMMs (bias / GELU / GELU' fused into the tcgen05 epilogues) and all gradients in libclipk.  The output
is bf16 (what the scorer consumes); there is no eager fallback.  Dropout (training mode, pacl.py:72,77) is FUSED into
the LayerNorm kernel: a counter-based mask drawn in the kernel from a seed of torch's CPU generator, kept as one bit
per element and replayed by the backward (the reference's own Philox stream cannot be reproduced bit-wise by any fused
mask; the tests replay OUR mask through the CPU restatement).  In eval mode it is the identity.
"""
import torch
import torch.nn as nn

from . import _lib
from .functional import _DT, _need_cuda, _stream, _f32


def _rows(x):
    D = x.shape[-1]
    return x.reshape(-1, D), x.shape[:-1]


class _LayerNormBf16(torch.autograd.Function):
    """xn = Dropout_p(LayerNorm(x)) as bf16 rows; x bf16 | fp32 [..., D].  p = 0: plain LayerNorm.  The dropout mask is
    drawn inside the kernel (counter-based on `seed`), kept as one bit per element and replayed by the backward."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, drop_p=0.0, seed=0):
        _need_cuda(x)
        if x.dtype not in _DT or x.dtype == torch.float16:
            x = x.float()
        x2, lead = _rows(x)
        x2 = x2.contiguous()
        R, D = x2.shape
        g, b = weight.float().contiguous(), bias.float().contiguous()
        xn = torch.empty(R, D, dtype=torch.bfloat16, device=x.device)
        mean, rstd = _f32(R, device=x.device), _f32(R, device=x.device)
        keep = torch.empty(R * D // 8, dtype=torch.uint8, device=x.device) if drop_p > 0 else None
        _lib.call("clipk_ln_fwd", x2.data_ptr(), _DT[x2.dtype], R, D, g.data_ptr(), b.data_ptr(), float(eps),
                  xn.data_ptr(), mean.data_ptr(), rstd.data_ptr(), float(drop_p), int(seed) & (2 ** 64 - 1),
                  0 if keep is None else keep.data_ptr(), _stream())
        ctx.save_for_backward(x2, g, mean, rstd, keep if keep is not None else torch.empty(0, device=x.device))
        ctx.lead = lead
        ctx.wdt = (weight.dtype, bias.dtype)
        ctx.drop_p = float(drop_p)
        ctx.mark_non_differentiable(*([keep] if keep is not None else []))
        if keep is not None:
            return xn.reshape(*lead, D), keep
        return xn.reshape(*lead, D)

    @staticmethod
    def backward(ctx, dxn, *_unused):
        x2, g, mean, rstd, keep = ctx.saved_tensors
        R, D = x2.shape
        dxn2 = dxn.reshape(R, D).to(torch.bfloat16).contiguous()
        dgamma, dbeta = _f32(D, device=x2.device), _f32(D, device=x2.device)
        dx = torch.empty_like(x2) if ctx.needs_input_grad[0] else None
        nbytes = _lib.lib().clipk_ln_bwd_workspace_bytes(R, D)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x2.device)
        _lib.call("clipk_ln_bwd", x2.data_ptr(), _DT[x2.dtype], R, D, g.data_ptr(), mean.data_ptr(), rstd.data_ptr(),
                  dxn2.data_ptr(), dgamma.data_ptr(), dbeta.data_ptr(), 0 if dx is None else dx.data_ptr(),
                  ctx.drop_p, keep.data_ptr() if ctx.drop_p > 0 else 0, ws.data_ptr(), nbytes, _stream())
        return (None if dx is None else dx.reshape(*ctx.lead, D), dgamma.to(ctx.wdt[0]), dbeta.to(ctx.wdt[1]), None,
                None, None)


def layer_norm_bf16(x, weight, bias, eps=1e-5, drop_p=0.0, seed=0, return_mask=False):
    """bf16 LayerNorm rows with the reference's Dropout fused behind it (training mode: `drop_p` > 0).  `return_mask`
    additionally returns the keep bits (uint8 [rows * D / 8], bit j of byte v = element 8 v + j) for mask-replay tests."""
    if drop_p > 0:
        xn, keep = _LayerNormBf16.apply(x, weight, bias, eps, float(drop_p), int(seed))
        return (xn, keep) if return_mask else xn
    xn = _LayerNormBf16.apply(x, weight, bias, eps)
    return (xn, None) if return_mask else xn


def _dropout_seed():
    """A fresh 63-bit seed from torch's CPU generator (follows torch.manual_seed; no device synchronisation)."""
    return int(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).item())


def unpack_keep_bits(keep, shape):
    """keep bits of `layer_norm_bf16(..., return_mask=True)` -> bool tensor of `shape` (test / debugging helper)."""
    bits = (keep.reshape(-1, 1) >> torch.arange(8, device=keep.device, dtype=torch.uint8)) & 1
    return bits.reshape(shape).bool()


class _PatchProj(torch.autograd.Function):
    """y = W1 xn + b1 + W3 gelu(W2 xn + b2) + b3 on bf16 rows (pacl.py:47-48)."""

    @staticmethod
    def forward(ctx, xn, W1, b1, W2, b2, W3, b3, want_sqnorm=False):
        _need_cuda(xn)
        x2, lead = _rows(xn)
        x2 = x2.to(torch.bfloat16).contiguous()
        R, Din = x2.shape
        Dout = W1.shape[0]
        dev = x2.device
        w1, w2, w3 = (w.detach().to(torch.bfloat16).contiguous() for w in (W1, W2, W3))
        b13 = (b1.detach().float() + b3.detach().float()).contiguous()
        b2f = b2.detach().float().contiguous()
        Gp = torch.empty(R, Dout, dtype=torch.bfloat16, device=dev)      # gelu'(z), saved for the backward
        H = torch.empty_like(Gp)
        Y = torch.empty_like(Gp)
        ysq = _f32(R, device=dev) if want_sqnorm else None
        _lib.call("clipk_patch_proj_fwd", x2.data_ptr(), R, Din, Dout, w1.data_ptr(), w2.data_ptr(), w3.data_ptr(),
                  b13.data_ptr(), b2f.data_ptr(), Gp.data_ptr(), H.data_ptr(), Y.data_ptr(),
                  0 if ysq is None else ysq.data_ptr(), _stream())
        ctx.save_for_backward(x2, Gp, H, w1, w2, w3)
        ctx.lead = lead
        ctx.dts = tuple(t.dtype for t in (xn, W1, b1, W2, b2, W3, b3))
        if want_sqnorm:
            ctx.mark_non_differentiable(ysq)
            return Y.reshape(*lead, Dout), ysq.reshape(*lead)
        return Y.reshape(*lead, Dout)

    @staticmethod
    def backward(ctx, dY, *_unused):
        x2, Gp, H, w1, w2, w3 = ctx.saved_tensors
        R, Din = x2.shape
        Dout = w1.shape[0]
        dev = x2.device
        dy = dY.reshape(R, Dout).to(torch.bfloat16).contiguous()
        dxn = torch.empty(R, Din, dtype=torch.bfloat16, device=dev) if ctx.needs_input_grad[0] else None
        dW1, dW2 = _f32(Dout, Din, device=dev), _f32(Dout, Din, device=dev)
        dW3 = _f32(Dout, Dout, device=dev)
        db13, db2 = _f32(Dout, device=dev), _f32(Dout, device=dev)
        nbytes = _lib.lib().clipk_patch_proj_bwd_workspace_bytes(R, Din, Dout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_patch_proj_bwd", x2.data_ptr(), Gp.data_ptr(), H.data_ptr(), dy.data_ptr(), R, Din, Dout,
                  w1.data_ptr(), w2.data_ptr(), w3.data_ptr(), 0 if dxn is None else dxn.data_ptr(), dW1.data_ptr(),
                  dW2.data_ptr(), dW3.data_ptr(), db13.data_ptr(), db2.data_ptr(), ws.data_ptr(), nbytes, _stream())
        dt = ctx.dts
        return (None if dxn is None else dxn.reshape(*ctx.lead, Din).to(dt[0]), dW1.to(dt[1]), db13.to(dt[2]),
                dW2.to(dt[3]), db2.to(dt[4]), dW3.to(dt[5]), db13.to(dt[6]), None)


class _Linear(torch.autograd.Function):
    """y = x W^T + b on bf16 rows (text_projection's Linear, pacl.py:78)."""

    @staticmethod
    def forward(ctx, x, W, b):
        _need_cuda(x)
        x2, lead = _rows(x)
        x2 = x2.to(torch.bfloat16).contiguous()
        R, Din = x2.shape
        Dout = W.shape[0]
        w = W.detach().to(torch.bfloat16).contiguous()
        bf = b.detach().float().contiguous()
        y = torch.empty(R, Dout, dtype=torch.bfloat16, device=x2.device)
        _lib.call("clipk_linear_fwd", x2.data_ptr(), R, Din, Dout, w.data_ptr(), bf.data_ptr(), y.data_ptr(), _stream())
        ctx.save_for_backward(x2, w)
        ctx.lead = lead
        ctx.dts = (x.dtype, W.dtype, b.dtype)
        return y.reshape(*lead, Dout)

    @staticmethod
    def backward(ctx, dY):
        x2, w = ctx.saved_tensors
        R, Din = x2.shape
        Dout = w.shape[0]
        dev = x2.device
        dy = dY.reshape(R, Dout).to(torch.bfloat16).contiguous()
        dx = torch.empty(R, Din, dtype=torch.bfloat16, device=dev) if ctx.needs_input_grad[0] else None
        dW, db = _f32(Dout, Din, device=dev), _f32(Dout, device=dev)
        nbytes = _lib.lib().clipk_linear_bwd_workspace_bytes(R, Din, Dout)
        ws = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        _lib.call("clipk_linear_bwd", x2.data_ptr(), dy.data_ptr(), R, Din, Dout, w.data_ptr(),
                  0 if dx is None else dx.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(), nbytes, _stream())
        dt = ctx.dts
        return (None if dx is None else dx.reshape(*ctx.lead, Din).to(dt[0]), dW.to(dt[1]), db.to(dt[2]))


class Patch_Projection(nn.Module):
    """Drop-in for the reference's `Patch_Projection` (pacl.py:35-48): same constructor, sub-modules and state dict.
    `forward(x)` expects the (LayerNorm-ed) token rows and returns bf16."""

    def __init__(self, in_dim=768, out_dim=512):
        super().__init__()
        self.linear_projection = self.text_projection = nn.Sequential(nn.Linear(in_dim, out_dim))
        self.non_linear_projection = nn.Sequential(nn.Linear(in_dim, out_dim), nn.GELU(), nn.Linear(out_dim, out_dim))

    def forward(self, x, return_sqnorm=False):
        """return_sqnorm: also return the squared L2 norm of every output row ([...] fp32), accumulated in the epilogue of
        the output GEMM -- hand it to `functional.pacl_scores(..., v_sqnorm=)` / `PaclAllPairsLoss` and the scorer skips
        its own pass over the patch tensor."""
        lin, nl0, nl2 = self.linear_projection[0], self.non_linear_projection[0], self.non_linear_projection[2]
        return _PatchProj.apply(x, lin.weight, lin.bias, nl0.weight, nl0.bias, nl2.weight, nl2.bias, return_sqnorm)


class VisualProjection(nn.Sequential):
    """`visual_projection` of the PACL models (pacl.py:70-74): Sequential(LayerNorm, Dropout, Patch_Projection) with
    the reference's state-dict keys ("0.weight", "2.linear_projection.0.weight", ...)."""

    def __init__(self, in_dim=1024, out_dim=768, p=0.1):
        super().__init__(nn.LayerNorm(in_dim), nn.Dropout(p), Patch_Projection(in_dim, out_dim))

    def forward(self, x, return_sqnorm=False):
        ln, drop, proj = self[0], self[1], self[2]
        p = drop.p if self.training else 0.0
        xn = layer_norm_bf16(x, ln.weight, ln.bias, ln.eps, p, _dropout_seed() if p > 0 else 0)     # dropout fused in
        return proj(xn, return_sqnorm)


class TextProjection(nn.Sequential):
    """`text_projection` of the PACL models (pacl.py:75-79): Sequential(LayerNorm, Dropout, Linear)."""

    def __init__(self, dim=768, out_dim=None, p=0.1):
        super().__init__(nn.LayerNorm(dim), nn.Dropout(p), nn.Linear(dim, out_dim or dim))

    def forward(self, x):
        ln, drop, lin = self[0], self[1], self[2]
        p = drop.p if self.training else 0.0
        xn = layer_norm_bf16(x, ln.weight, ln.bias, ln.eps, p, _dropout_seed() if p > 0 else 0)
        return _Linear.apply(xn, lin.weight, lin.bias)


# ------------------------------------------------------------------------------------------------ RoPE (§8f rank 3)
_ROPE_TABLES = {}


def _rope_tables(S, D, device):
    """sin / cos [S, D/2] fp32, computed on the CPU with the reference's own expressions (pacl.py:160-171) so that the
    angles are bit-identical, cached per (S, D, device)."""
    key = (S, D, str(device))
    if key not in _ROPE_TABLES:
        inv_freq = 1.0 / (10000 ** (torch.arange(0, D, 2).float() / D))
        angles = torch.arange(S, dtype=torch.float).unsqueeze(1) * inv_freq
        _ROPE_TABLES[key] = (torch.sin(angles).to(device).contiguous(), torch.cos(angles).to(device).contiguous())
    return _ROPE_TABLES[key]


class _Rope(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, out_dtype):
        _need_cuda(x)
        if x.dtype not in _DT:
            x = x.float()
        xc = x.contiguous()
        B, S, D = xc.shape
        sn, cs = _rope_tables(S, D, xc.device)
        y = torch.empty(B, S, D, dtype=out_dtype, device=xc.device)
        _lib.call("clipk_rope", xc.data_ptr(), _DT[xc.dtype], B * S, S, D, sn.data_ptr(), cs.data_ptr(), y.data_ptr(),
                  _DT[out_dtype], 0, _stream())
        ctx.cfg = (x.dtype, S, D)
        return y

    @staticmethod
    def backward(ctx, g):
        xdt, S, D = ctx.cfg
        if g.dtype not in _DT:
            g = g.float()
        gc = g.contiguous()
        sn, cs = _rope_tables(S, D, gc.device)
        dx = torch.empty(gc.shape, dtype=xdt if xdt in _DT else torch.float32, device=gc.device)
        _lib.call("clipk_rope", gc.data_ptr(), _DT[gc.dtype], gc.shape[0] * S, S, D, sn.data_ptr(), cs.data_ptr(),
                  dx.data_ptr(), _DT[dx.dtype], 1, _stream())
        return dx.to(xdt), None


def apply_rope(embeddings, out_dtype=None):
    """Drop-in for `apply_rope` (pacl.py:147-181): embeddings [B, S, D] -> rotated [B, S, D] (same dtype unless
    `out_dtype` is given; bf16 output feeds `VisualProjection` directly, as `open_clip_pacl_rope.forward` does)."""
    assert embeddings.shape[-1] % 2 == 0, "Embedding dimension must be even for RoPE."
    od = out_dtype or (embeddings.dtype if embeddings.dtype in _DT else torch.float32)
    return _Rope.apply(embeddings, od)
