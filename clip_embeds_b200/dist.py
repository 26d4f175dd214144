"""torch.distributed plumbing for the sharded losses (one process per GPU; NCCL on GPUs, gloo in the CPU tests).

Reference: open_clip/src/open_clip/loss.py:21-87 (gather_features, gather_features_diffsize).  Differences by
design: the variable-size text gather is ONE fixed-capacity all-gather (capacity 2b rows: at
most one hard caption per sample, open_clip_train/data.py:110-116) plus the row counts, instead of size gather ->
host max() -> padded gather; the gradient of a gather is a reduce-scatter (what torch.distributed.nn.all_gather's
backward amounts to).
"""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


def all_gather_nograd(x, group=None):
    """Concatenation over ranks of equally-shaped tensors (no gradient)."""
    W = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x, group=group)
    else:
        dist.all_gather(list(out.chunk(W, dim=0)), x, group=group)
    return out


class _AllGatherGrad(torch.autograd.Function):
    """Concatenating all-gather; backward = reduce-scatter(SUM) of the gradient of the gathered tensor."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_nograd(x, group)

    @staticmethod
    def backward(ctx, g):
        W = dist.get_world_size(ctx.group)
        g = g.contiguous()
        n = g.shape[0] // W
        if g.is_cuda:
            out = torch.empty((n,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
            return out, None
        g = g.clone()                      # gloo has no reduce_scatter: all-reduce, then keep the local slice
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        r = dist.get_rank(ctx.group)
        return g[r * n:(r + 1) * n].clone(), None


def all_gather_with_grad(x, group=None):
    return _AllGatherGrad.apply(x, group)


def all_reduce_sum_(x, group=None):
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


class _AllReduceSumGrad(torch.autograd.Function):
    """y = sum over ranks of x (the global loss, for logging parity); backward passes the upstream gradient to the
    local x unchanged, which is the gradient of the GLOBAL loss w.r.t. this rank's contribution."""

    @staticmethod
    def forward(ctx, x, group):
        y = x.detach().clone()
        dist.all_reduce(y, op=dist.ReduceOp.SUM, group=group)
        return y

    @staticmethod
    def backward(ctx, g):
        return g, None


def all_reduce_sum_with_grad(x, group=None):
    return _AllReduceSumGrad.apply(x, group)


class _ScaleGrad(torch.autograd.Function):
    """Identity in the forward pass; multiplies the gradient by `factor` in the backward pass."""

    @staticmethod
    def forward(ctx, x, factor):
        ctx.factor = factor
        return x.view_as(x)

    @staticmethod
    def backward(ctx, g):
        return g * ctx.factor, None


def scale_grad(x, factor):
    return _ScaleGrad.apply(x, float(factor))


def _splice_local(all_x, local_x, rank_):
    n = local_x.shape[0]
    return torch.cat([all_x[: rank_ * n], local_x, all_x[(rank_ + 1) * n:]], dim=0)


def gather_features(image_features, text_features, b, usehardtext, gather_with_grad, local_loss, rank_, world,
                    group=None, keep_padding=False):
    """(all_image [N,D], all_text [N + sum_r H_r, D]); text rows ordered [orig_0 .. orig_{W-1}, hard_0 .. hard_{W-1}]
    exactly as loss.py:147-153 re-orders them.

    keep_padding=True (usehardtext only) returns (all_image, all_text [2N, D], counts int32 [W]) instead: the hard
    negatives of rank r occupy rows [N + r b, N + r b + counts[r]) and the rest of each b-row slab is zero padding that
    the CE kernels mask out -- nothing about the ragged sizes ever reaches the host."""
    gather = all_gather_with_grad if gather_with_grad else all_gather_nograd
    all_img = gather(image_features, group) if image_features is not None else None   # None: the caller needs no images
    if not usehardtext:
        all_txt = gather(text_features, group)
        if not gather_with_grad and not local_loss:
            all_img = _splice_local(all_img, image_features, rank_)      # keep the local slice differentiable
            all_txt = _splice_local(all_txt, text_features, rank_)       # (loss.py:57-60)
        return all_img, all_txt
    D = text_features.shape[1]
    h = text_features.shape[0] - b
    if not 0 <= h <= b:
        raise ValueError("expected b originals followed by at most b hard negatives "
                         "(one hard caption per sample, open_clip_train/data.py:110-116)")
    # fixed-capacity slab: [b originals | h hard negatives | zero padding up to 2b rows]
    padded = text_features
    if h < b:
        padded = torch.cat([text_features, text_features.new_zeros((b - h, D))], dim=0)
    counts = all_gather_nograd(torch.tensor([h], dtype=torch.int32, device=text_features.device), group)
    slabs = gather(padded, group).reshape(world, 2 * b, D)
    if keep_padding:
        # no host sync, no ragged cat: [orig_0 .. orig_{W-1} | slab_0 .. slab_{W-1}] with every hard-negative slab at its
        # full capacity of b rows; the consumer masks rows >= counts[r] of slab r on the device (feat_row_ce `slab`)
        all_txt = torch.cat([slabs[:, :b].reshape(world * b, D), slabs[:, b:].reshape(world * b, D)], dim=0)
        return all_img, all_txt, counts
    hs = counts.tolist()                     # legacy compact layout: one host read of the ragged tail's sizes
    orig = slabs[:, :b].reshape(world * b, D)
    hard = [slabs[r, b:b + hs[r]] for r in range(world)]
    return all_img, torch.cat([orig] + hard, dim=0)
