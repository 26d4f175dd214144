"""GPU parity: retrieval metrics fused into the logits GEMM and the VLM2Vec contrastive loss against
the oracle's restatement of open_clip's get_clip_metrics (train.py:360-377) and VLM2Vec's SimpleContrastiveLoss
(src/loss.py:7-19)."""
import pytest
import torch

from conftest import rel_l2
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("N,D,noise", [(1000, 64, 0.7), (2000, 64, 3.0), (5000, 512, 4.0), (300, 768, 0.7)])
def test_clip_metrics_match_oracle(N, D, noise):
    """Features are bf16-representable, so the oracle's fp32 logits and the tensor-core logits differ only in the fp32
    accumulation order; ranks are integers: every position must agree except for near-ties (<= 0.2 % of the rows may
    move, by one place), and the metrics within the corresponding bounds."""
    from clip_embeds_b200.metrics import get_clip_metrics, retrieval_ranks
    g = torch.Generator().manual_seed(N)
    base = torch.randn(N, D, generator=g)
    img = O.l2n(base + noise * torch.randn(N, D, generator=g)).to(torch.bfloat16)
    txt = O.l2n(base + noise * torch.randn(N, D, generator=g)).to(torch.bfloat16)
    want, preds = O.clip_metrics(img.float(), txt.float(), 100.0)
    rr, rc = retrieval_ranks(img.cuda(), txt.cuda())
    for name, r in (("image_to_text", rr), ("text_to_image", rc)):
        diff = (r.cpu().long() - torch.from_numpy(preds[name]).long()).abs()
        print(f"N={N} noise={noise} {name}: {int((diff != 0).sum())} of {N} ranks differ (max {int(diff.max())}); mean rank {preds[name].mean():.1f}")
        assert int((diff != 0).sum()) <= max(1, N // 500) and int(diff.max()) <= 1
    got = get_clip_metrics(img.cuda(), txt.cuda(), 100.0)
    for k, v in want.items():
        tol = 0.01 if "mean_rank" in k else (1.0 if "median" in k else 2.0 / N)
        assert abs(got[k] - float(v)) <= tol, (k, got[k], v)


def test_vlm2vec_simple_contrastive_loss():
    from clip_embeds_b200.metrics import SimpleContrastiveLoss
    n, tpq, D = 256, 3, 256
    xb = O.l2n(O.rn(91, n, D)).to(torch.bfloat16)
    yb = O.l2n(O.rn(92, n * tpq, D)).to(torch.bfloat16)
    xo, yo = xb.float().requires_grad_(), yb.float().requires_grad_()
    lo = O.simple_contrastive_loss(xo, yo, 0.02)
    lo.backward()
    x, y = xb.cuda().requires_grad_(), yb.cuda().requires_grad_()
    loss = SimpleContrastiveLoss(0.02)(x, y)
    loss.backward()
    assert abs(loss.item() - lo.item()) < 2e-3 * max(1.0, abs(lo.item()))
    assert rel_l2(x.grad.float().cpu(), xo.grad) < 2e-2 and rel_l2(y.grad.float().cpu(), yo.grad) < 2e-2
