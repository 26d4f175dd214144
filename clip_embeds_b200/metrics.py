"""Retrieval metrics and the remaining InfoNCE users.

  get_clip_metrics(image_features, text_features, logit_scale)      open_clip/src/open_clip_train/train.py:360-377
      reference: [N,N] logits moved to the CPU, a full argsort per direction, position of the ground truth
      here: the rank of the diagonal is counted in the epilogue of the tcgen05 logits GEMM (both directions from the
      same tiles, logits never written); only the handful of statistics crosses to the host.
  SimpleContrastiveLoss(temperature)(x, y, target=None, reduction)   VLM2Vec/src/loss.py:7-19
      = the feature cross-entropy kernels with VLM2Vec's strided label map.
"""
import torch

from . import _lib
from . import functional as Fk


def retrieval_ranks(x, y):
    """rank_row[i] = #{j != i : <x_i, y_j> > <x_i, y_i>}, rank_col[j] = #{i != j : <x_i, y_j> > <x_j, y_j>} (int32, on
    the device).  Features are used in bf16 (tensor cores, fp32 accumulation)."""
    Fk._need_cuda(x, y)
    xb, yb = x.detach().to(torch.bfloat16).contiguous(), y.detach().to(torch.bfloat16).contiguous()
    M, D = xb.shape
    N = yb.shape[0]
    dev = xb.device
    diag = torch.empty(min(M, N), dtype=torch.float32, device=dev)
    rr = torch.empty(M, dtype=torch.int32, device=dev)
    rc = torch.empty(N, dtype=torch.int32, device=dev)
    _lib.call("clipk_retrieval_ranks", xb.data_ptr(), yb.data_ptr(), M, N, D, diag.data_ptr(), rr.data_ptr(), rc.data_ptr(),
              Fk._stream())
    return rr, rc


def _rank_metrics(ranks, name, out):
    n = ranks.numel()
    stats = torch.empty(4, dtype=torch.int64, device=ranks.device)
    _lib.call("clipk_rank_stats", ranks.data_ptr(), n, stats.data_ptr(), Fk._stream())
    srt = torch.sort(ranks).values
    mid = srt[[(n - 1) // 2, n // 2]].tolist()            # np.median: mean of the two middle values for even n
    s, r1, r5, r10 = stats.tolist()
    out[f"{name}_mean_rank"] = s / n + 1
    out[f"{name}_median_rank"] = float((mid[0] + mid[1]) / 2 // 1 + 1)
    out[f"{name}_R@1"], out[f"{name}_R@5"], out[f"{name}_R@10"] = r1 / n, r5 / n, r10 / n


def get_clip_metrics(image_features, text_features, logit_scale):
    """Drop-in for open_clip's get_clip_metrics (train.py:360-377).  `logit_scale` must be positive (it does not
    change the ranking).  Ties are ranked in favour of the ground truth (the reference's argsort leaves their order
    unspecified)."""
    if float(logit_scale) <= 0:
        raise ValueError("get_clip_metrics: logit_scale must be positive")
    rr, rc = retrieval_ranks(image_features, text_features)
    metrics = {}
    _rank_metrics(rr, "image_to_text", metrics)
    _rank_metrics(rc, "text_to_image", metrics)
    return metrics


class SimpleContrastiveLoss:
    """VLM2Vec/src/loss.py:7-19: cross_entropy(x @ y.T / temperature, target), target = arange(0, n * t, t) with
    t = len(y) // len(x) candidates per query."""

    def __init__(self, temperature: float = 0.02):
        self.temperature = temperature

    def __call__(self, x, y, target=None, reduction="mean"):
        if target is None:
            tpq = y.size(0) // x.size(0)
            target = torch.arange(0, x.size(0) * tpq, tpq, device=x.device, dtype=torch.long)
        loss_sum, valid = Fk._FeatRowCE.apply(x, y, 1.0 / self.temperature, 0.0, target, 0, None)
        if reduction == "sum":
            return loss_sum
        if reduction != "mean":
            raise ValueError(f"reduction {reduction!r} is not supported (mean | sum)")
        return loss_sum / valid.sum().clamp_min(1)
