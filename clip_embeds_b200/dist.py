"""torch.distributed plumbing for the sharded losses (one process per GPU, NCCL on GPUs, gloo in CPU tests).

Reference: open_clip/src/open_clip/loss.py:21-87 (gather_features, gather_features_diffsize).  Differences by
design: the variable-size text gather uses ONE fixed-capacity all-gather (capacity 2b rows:
at most one hard caption per sample, open_clip_train/data.py:110-116) plus the row counts, with no host-side
`max()` sync between collectives; the gradient of a gather is a reduce-scatter (what
torch.distributed.nn.all_gather's backward does).
"""
import torch
import torch.distributed as dist


def world_size(group=None):
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


def rank(group=None):
    return dist.get_rank(group) if dist.is_available() and dist.is_initialized() else 0


class _AllGatherGrad(torch.autograd.Function):
    """Concatenating all-gather of equally-shaped tensors; backward = reduce-scatter(SUM) of the gradient."""

    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        W = dist.get_world_size(group)
        x = x.contiguous()
        out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
        dist.all_gather_into_tensor(out, x, group=group)
        return out

    @staticmethod
    def backward(ctx, g):
        W = dist.get_world_size(ctx.group)
        g = g.contiguous()
        out = torch.empty((g.shape[0] // W,) + tuple(g.shape[1:]), dtype=g.dtype, device=g.device)
        if g.is_cuda:
            dist.reduce_scatter_tensor(out, g, op=dist.ReduceOp.SUM, group=ctx.group)
        else:   # gloo has no reduce_scatter: all-reduce then slice (CPU tests only)
            g = g.clone()
            dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
            r = dist.get_rank(ctx.group)
            out.copy_(g[r * out.shape[0]:(r + 1) * out.shape[0]])
        return out, None


def all_gather_with_grad(x, group=None):
    return _AllGatherGrad.apply(x, group)


def all_gather_nograd(x, group=None):
    W = dist.get_world_size(group)
    x = x.contiguous()
    out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x, group=group)
    else:
        parts = list(out.chunk(W, dim=0))
        dist.all_gather(parts, x, group=group)
    return out


def gather_features(image_features, text_features, b, usehardtext, gather_with_grad, local_loss, rank_, world, group=None):
    """Returns (all_image [N,D], all_text [N (+ sum H_r), D]) with the text rows ordered
    [orig_0 .. orig_{W-1}, hard_0 .. hard_{W-1}] (loss.py:147-153)."""
    gather = all_gather_with_grad if gather_with_grad else all_gather_nograd
    if image_features.is_cuda is False and gather_with_grad:
        gather = all_gather_with_grad
    all_img = _gather_cpu_safe(image_features, gather, group)
    if not usehardtext:
        all_txt = _gather_cpu_safe(text_features, gather, group)
        if not gather_with_grad and not local_loss:
            # keep the local slice differentiable (loss.py:57-60)
            all_img = _splice_local(all_img, image_features, rank_)
            all_txt = _splice_local(all_txt, text_features, rank_)
        return all_img, all_txt
    # fixed-capacity gather: [b originals | up to b hard negatives, zero padded] + the true hard count
    D = text_features.shape[1]
    h = text_features.shape[0] - b
    assert 0 <= h <= b, "at most one hard negative per sample (open_clip_train/data.py:110-116)"
    pad = torch.zeros((2 * b, D), dtype=text_features.dtype, device=text_features.device)
    padded = torch.cat([text_features, pad[: b - h]], dim=0) if h < b else text_features
    counts = torch.tensor([h], dtype=torch.int64, device=text_features.device)
    all_counts = all_gather_nograd(counts, group)                      # [W]
    all_padded = _gather_cpu_safe(padded, gather, group).reshape(world, 2 * b, D)
    orig = all_padded[:, :b].reshape(world * b, D)
    hs = all_counts.tolist()                                           # one sync for the ragged concat
    hard = [all_padded[r, b:b + hs[r]] for r in range(world)]
    return all_img, torch.cat([orig] + hard, dim=0)


def _gather_cpu_safe(x, gather, group):
    if x.is_cuda or gather is all_gather_nograd:
        return gather(x, group)
    return _AllGatherGradCPU.apply(x, group)


class _AllGatherGradCPU(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, group):
        ctx.group = group
        return all_gather_nograd(x, group)

    @staticmethod
    def backward(ctx, g):
        W = dist.get_world_size(ctx.group)
        g = g.contiguous().clone()
        dist.all_reduce(g, op=dist.ReduceOp.SUM, group=ctx.group)
        r = dist.get_rank(ctx.group)
        n = g.shape[0] // W
        return g[r * n:(r + 1) * n], None


def _splice_local(all_x, local_x, rank_):
    n = local_x.shape[0]
    return torch.cat([all_x[: rank_ * n], local_x, all_x[(rank_ + 1) * n:]], dim=0)


def all_reduce_sum_(x, group=None):
    dist.all_reduce(x, op=dist.ReduceOp.SUM, group=group)
    return x


class _AllReduceSumGrad(torch.autograd.Function):
    """y = sum over ranks of x (for logging the global loss); backward is the identity on the local x, which is
    the correct gradient of the GLOBAL loss w.r.t. this rank's contribution."""

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
