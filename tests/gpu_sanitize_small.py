"""Small-shape tour of every tcgen05 kernel family for compute-sanitizer (diagnostic; one tool per gpurun call):

    compute-sanitizer --tool memcheck  python tests/gpu_sanitize_small.py
    compute-sanitizer --tool racecheck python tests/gpu_sanitize_small.py

Covers: fused all-pairs forward (both variants), dual dS GEMM / dT / dV backward, staged forward (shape the fused kernel
does not take), recompute backward, persistent kernel, feature CE (with slabs + device scale), SPARC, projection heads,
paired path.  Shapes are tiny so that the tools finish in minutes; results are checked for finiteness only (parity is
what tests/ is for)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clip_embeds_b200.functional as Fk  # noqa: E402
from clip_embeds_b200 import losses  # noqa: E402
from clip_embeds_b200.models import SparcHead  # noqa: E402


def ok(name, *ts):
    torch.cuda.synchronize()
    good = all(bool(torch.isfinite(t.float()).all()) for t in ts)
    print(("ok   " if good else "BAD  ") + name, flush=True)
    return good


def allpairs(Bi, Bt, P, D, group=None, act="sigmoid"):
    V = torch.randn(Bi, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
    T = torch.randn(Bt, D, device="cuda").to(torch.bfloat16).requires_grad_()
    s = Fk.pacl_scores(V, T, 10.0, act, group)
    s.backward(torch.randn_like(s) / Bt)
    return s, V.grad, T.grad


def main():
    torch.manual_seed(0)
    good = True
    good &= ok("all-pairs fused fwd (save pooled) + pooled bwd, P=196 D=512", *allpairs(5, 300, 196, 512))
    good &= ok("all-pairs fused fwd P=576 D=768 softmax", *allpairs(3, 130, 576, 768, act="softmax"))
    with torch.no_grad():
        good &= ok("all-pairs fused fwd (no save)", Fk.pacl_scores(torch.randn(4, 100, 128, device="cuda"),
                                                                  torch.randn(70, 128, device="cuda"), 10.0))
    good &= ok("all-pairs staged fwd (D=64) + pooled bwd", *allpairs(6, 40, 50, 64, group=(2, 2)))
    os.environ["CLIPK_AP_SAVE_POOLED"] = "0"
    good &= ok("all-pairs recompute bwd", *allpairs(5, 130, 196, 512))
    os.environ["CLIPK_AP_SAVE_POOLED"] = "1"
    good &= ok("all-pairs persistent kernel", *allpairs(7, 130, 196, 512, group=(-3, 2)))
    # feature CE with hard-negative slabs and a device scale
    X = torch.nn.functional.normalize(torch.randn(160, 256, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_()
    Y = torch.nn.functional.normalize(torch.randn(192 + 3 * 64, 256, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_()
    cnt = torch.tensor([17, 0, 64], dtype=torch.int32, device="cuda")
    loss = Fk.feat_row_ce(X, Y, torch.tensor(20.0, device="cuda"), 0.0, None, 3, (cnt, 192, 64))
    loss.backward()
    good &= ok("feature CE (slabs, device scale)", loss, X.grad, Y.grad)
    img = torch.nn.functional.normalize(torch.randn(300, 256, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_()
    txt = torch.nn.functional.normalize(torch.randn(340, 256, device="cuda"), dim=-1).to(torch.bfloat16).requires_grad_()
    l2 = losses.OpenClipLoss(usehardtext=True)(img, txt, 30.0)
    l2.backward()
    good &= ok("OpenClipLoss bf16", l2, img.grad, txt.grad)
    # SPARC
    V = torch.randn(4, 196, 256, device="cuda").to(torch.bfloat16).requires_grad_()
    L = torch.randn(4, 77, 256, device="cuda").to(torch.bfloat16).requires_grad_()
    mask = (torch.arange(77, device="cuda")[None] <= torch.tensor([5, 20, 76, 40], device="cuda")[:, None]).float()
    v2, lh, gh, m2 = SparcHead(1.0 / 196)(V, L, mask)
    l3 = losses.SparcLoss(0.1)(v2, lh, gh, m2)
    l3.backward()
    good &= ok("SPARC align + SparcLoss", l3, V.grad, L.grad)
    # paired path + ClipLoss
    Vp = torch.randn(8, 196, 512, device="cuda").requires_grad_()
    Tp = torch.randn(8, 512, device="cuda").requires_grad_()
    i2, t2 = Fk.pacl_pool(Vp, Tp, "sigmoid")
    l4 = losses.ClipLoss(0.1)(i2, t2)
    l4.backward()
    good &= ok("paired + ClipLoss", l4, Vp.grad, Tp.grad)
    # projection heads
    from clip_embeds_b200 import heads
    vp = heads.VisualProjection(256, 128).cuda().eval()
    x = torch.randn(3, 50, 256, device="cuda", requires_grad=True)
    y = vp(x)
    y.float().sum().backward()
    good &= ok("visual projection head", y, x.grad)
    print("ALL OK" if good else "FAILURES", flush=True)
    sys.exit(0 if good else 1)


if __name__ == "__main__":
    main()
