"""GPU parity: cross-entropy paths (fp32 SIMT, bf16 tcgen05 engine) vs the oracle / reference goldens."""
import pytest
import torch

from conftest import rel_l2
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def test_openclip_hardneg_fp32_golden(goldens):
    from clip_embeds_b200.losses import OpenClipLoss
    G = goldens["G3"]
    img = O.l2n(O.rn(5, 8, 16)).cuda().requires_grad_()
    txt = O.l2n(O.rn(6, 11, 16)).cuda().requires_grad_()
    loss = OpenClipLoss(usehardtext=True)(img, txt, torch.tensor(100.0))
    loss.backward()
    assert abs(loss.item() - G["loss"].item()) < 1e-4 * G["loss"].item()
    assert rel_l2(img.grad.cpu(), G["dimg"]) < 1e-4
    assert rel_l2(txt.grad.cpu(), G["dtxt"]) < 1e-4
    img = O.l2n(O.rn(5, 8, 16)).cuda().requires_grad_()
    txt = O.l2n(O.rn(6, 8, 16)).cuda().requires_grad_()
    out = OpenClipLoss()(img, txt, torch.tensor(20.0), logit_bias=torch.tensor(-3.0), output_dict=True)
    out["contrastive_loss"].backward()
    assert abs(out["contrastive_loss"].item() - G["loss_plain"].item()) < 1e-5 * max(1, G["loss_plain"].item())
    assert rel_l2(img.grad.cpu(), G["dimg_plain"]) < 1e-4


@pytest.mark.parametrize("shape", [(8, 11, 16), (300, 700, 256), (1024, 1300, 768)])
def test_openclip_hardneg_bf16_engine(shape):
    """bf16 tensor-core CE vs fp32 oracle on the same bf16-rounded features.  Tolerance: loss 2e-3 rel,
    gradients 2e-2 rel-L2 (dL is stored in bf16 before the two gradient GEMMs)."""
    from clip_embeds_b200.losses import OpenClipLoss
    B, Nt, D = shape
    ib = O.l2n(O.rn(51, B, D)).to(torch.bfloat16)
    tb = O.l2n(O.rn(52, Nt, D)).to(torch.bfloat16)
    io = ib.float().requires_grad_()
    to = tb.float().requires_grad_()
    lo = O.openclip_loss_single(io, to, 30.0, usehardtext=True)
    lo.backward()
    img = ib.cuda().requires_grad_()
    txt = tb.cuda().requires_grad_()
    scale = torch.tensor(30.0, device="cuda", requires_grad=True)
    loss = OpenClipLoss(usehardtext=True)(img, txt, scale)
    loss.backward()
    print(f"ce bf16 {shape}: loss {loss.item():.6f} vs {lo.item():.6f} rel_dimg {rel_l2(img.grad.float().cpu(), io.grad):.3e} "
          f"rel_dtxt {rel_l2(txt.grad.float().cpu(), to.grad):.3e}")
    assert abs(loss.item() - lo.item()) < 2e-3 * max(1.0, abs(lo.item()))
    assert rel_l2(img.grad.float().cpu(), io.grad) < 2e-2
    assert rel_l2(txt.grad.float().cpu(), to.grad) < 2e-2
    # d loss / d logit_scale
    so = torch.tensor(30.0, requires_grad=True)
    O.openclip_loss_single(ib.float(), tb.float(), so, usehardtext=True).backward()
    assert abs(scale.grad.item() - so.grad.item()) < 2e-2 * max(1e-3, abs(so.grad.item())) + 1e-5


@pytest.mark.parametrize("B", [512, 2048])
def test_pacl_cliploss_bf16(B):
    """bf16 ClipLoss (small symmetric batch: one tensor-core logits GEMM, fp32 row + column CE, two gradient GEMMs)."""
    from clip_embeds_b200.losses import ClipLoss
    D = 768
    ib = O.l2n(O.rn(61, B, D)).to(torch.bfloat16)
    tb = O.l2n(O.rn(62, B, D)).to(torch.bfloat16)
    io = ib.float().requires_grad_()
    to = tb.float().requires_grad_()
    lo = O.pacl_clip_loss(io, to, 0.1)
    lo.backward()
    img = ib.cuda().requires_grad_()
    txt = tb.cuda().requires_grad_()
    loss = ClipLoss(0.1)(img, txt)
    loss.backward()
    assert abs(loss.item() - lo.item()) < 2e-3 * abs(lo.item())
    assert rel_l2(img.grad.float().cpu(), io.grad) < 2e-2
    assert rel_l2(txt.grad.float().cpu(), to.grad) < 2e-2


def test_negclip_full_size_rank_share():
    """BASELINE.json configs[3] at full size, one rank's share at W = 8: b = 4096 local images against N + sum H =
    32768 + 8192 gathered texts, and the b + H_r local texts against the 32768 gathered images (loss.py:156-164 with
    `usehardtext`).  Too large for the CPU oracle in seconds: the tensor-core path (logits never written) is compared
    with the same formula evaluated by fp32 torch ops on the same bf16-rounded features (logits materialised),
    plus two size-independent properties: rows with ignore_index contribute no gradient, and the image-side loss is
    invariant to a permutation of the gathered texts that keeps the positives in place."""
    import torch.nn.functional as F
    import clip_embeds_b200.functional as Fk
    b, N, H, D, Hr = 4096, 32768, 8192, 768, 1024
    g = torch.Generator().manual_seed(5)
    f = lambda n: torch.nn.functional.normalize(torch.randn(n, D, generator=g), dim=-1).to(torch.bfloat16).cuda()
    img_loc, txt_all, txt_loc, img_all = f(b), f(N + H), f(b + Hr), f(N)
    lab_t = torch.full((b + Hr,), -100, dtype=torch.int64, device="cuda")
    lab_t[:b] = torch.arange(b, device="cuda") + 3 * b          # rank 3's label offset
    ours = [t.detach().requires_grad_() for t in (img_loc, txt_all, txt_loc, img_all)]
    li = Fk.feat_row_ce(ours[0], ours[1], 100.0, 0.0, None, 3 * b)
    lt = Fk.feat_row_ce(ours[2], ours[3], 100.0, 0.0, lab_t, 0)
    ((li + lt) / 2).backward()
    ref = [t.detach().float().requires_grad_() for t in (img_loc, txt_all, txt_loc, img_all)]
    Li = 100.0 * ref[0] @ ref[1].T
    Lt = 100.0 * ref[2] @ ref[3].T
    lri = F.cross_entropy(Li, torch.arange(b, device="cuda") + 3 * b)
    lrt = F.cross_entropy(Lt, lab_t, ignore_index=-100)
    ((lri + lrt) / 2).backward()
    assert abs(li.item() - lri.item()) < 2e-3 * max(1.0, lri.item())
    assert abs(lt.item() - lrt.item()) < 2e-3 * max(1.0, lrt.item())
    for o, r in zip(ours, ref):
        assert rel_l2(o.grad.float(), r.grad) < 2e-2
    assert ours[2].grad[b:].abs().max().item() == 0.0            # hard-negative text rows are ignore_index rows
    # permuting the negatives among themselves does not change the image-side loss
    perm = torch.arange(N + H, device="cuda")
    perm[4 * b:] = perm[4 * b:].flip(0)
    li_p = Fk.feat_row_ce(img_loc, txt_all[perm].contiguous(), 100.0, 0.0, None, 3 * b)
    assert abs(li_p.item() - li.item()) < 1e-4 * max(1.0, li.item())


@pytest.mark.parametrize("shape", [(512, 512, 768), (1024, 768, 1024), (300, 200, 100), (129, 257, 65)])
def test_gemm_f32_split_matches_fp64(shape):
    """fp32 GEMM on the bf16 tensor cores (3-term split, six products in one tcgen05 GEMM) against an fp64 product,
    in the three operand layouts the fp32 feature path uses (X Y^T, dL Y, dL^T X).  Tolerance: max abs error
    <= 2e-6 * sqrt(K) * max|a| max|b| -- the error level of an fp32-accumulating GEMM."""
    import clip_embeds_b200.functional as Fk
    M, N, K = shape
    g = torch.Generator().manual_seed(11)
    for layout in ("nt", "nn", "tn"):
        A = torch.randn(M, K, generator=g).cuda()
        B = torch.randn(N, K, generator=g).cuda()
        ref = (A.double() @ B.double().T) * 0.5
        C = torch.full((M, N), float("nan"), device="cuda")
        if layout == "nt":        # A [M,K] K-major, B [N,K] K-major
            Fk._gemm_f32(A, K, 1, B, 1, K, C, N, M, N, K, 0.5)
        elif layout == "nn":      # B stored [K,N]
            Bt = B.T.contiguous()
            Fk._gemm_f32(A, K, 1, Bt, N, 1, C, N, M, N, K, 0.5)
        else:                     # A stored [K,M], B stored [K,N]
            At, Bt = A.T.contiguous(), B.T.contiguous()
            Fk._gemm_f32(At, 1, M, Bt, N, 1, C, N, M, N, K, 0.5)
        err = (C.double() - ref).abs().max().item()
        tol = 2e-6 * (K ** 0.5) * A.abs().max().item() * B.abs().max().item()
        print(f"gemm_f32_split {shape} {layout}: max abs err {err:.3e} (tol {tol:.3e})")
        assert err <= tol
        C2 = torch.ones(M, N, device="cuda")
        if layout == "nt":
            Fk._gemm_f32(A, K, 1, B, 1, K, C2, N, M, N, K, 0.5, accumulate=True)
            assert (C2.double() - 1.0 - ref).abs().max().item() <= tol + 1e-6


@pytest.mark.parametrize("B", [64, 512, 1024])
def test_pacl_cliploss_fp32(B):
    """fp32 ClipLoss (reference training dtype): one logits GEMM on the tensor cores at fp32 accuracy + row/column CE,
    against the oracle.  Tolerance: loss 1e-5 relative, gradients 2e-5 relative L2."""
    from clip_embeds_b200.losses import ClipLoss
    D = 768
    io = O.l2n(O.rn(61, B, D)).requires_grad_()
    to = O.l2n(O.rn(62, B, D)).requires_grad_()
    lo = O.pacl_clip_loss(io, to, 0.1)
    lo.backward()
    img = io.detach().cuda().requires_grad_()
    txt = to.detach().cuda().requires_grad_()
    loss = ClipLoss(0.1)(img, txt)
    loss.backward()
    print(f"cliploss fp32 B={B}: {loss.item():.7f} vs {lo.item():.7f}; rel dimg {rel_l2(img.grad.cpu(), io.grad):.2e} "
          f"dtxt {rel_l2(txt.grad.cpu(), to.grad):.2e}")
    assert abs(loss.item() - lo.item()) < 1e-5 * abs(lo.item())
    assert rel_l2(img.grad.cpu(), io.grad) < 2e-5
    assert rel_l2(txt.grad.cpu(), to.grad) < 2e-5


@pytest.mark.parametrize("dtype", [torch.bfloat16, torch.float32])
def test_feat_row_ce_hard_negative_slabs(dtype):
    """Fixed-capacity hard-negative slabs masked on the device (`slab`; reference: the host-side size exchange of
    gather_features_diffsize, loss.py:78-86): CE over Y = [N originals | W slabs of b rows, counts[r] of them filled] equals
    CE over the compact Y, and the padding rows receive exactly zero gradient -- with the logit scale as a device tensor."""
    import clip_embeds_b200.functional as Fk
    W, b, D, M = 3, 64, 256, 160
    counts = [17, 0, 64]
    N0 = W * b
    g = torch.Generator().manual_seed(123)
    X = O.l2n(torch.randn(M, D, generator=g)).to(dtype)
    orig = O.l2n(torch.randn(N0, D, generator=g)).to(dtype)
    hard = [O.l2n(torch.randn(c, D, generator=g)).to(dtype) for c in counts]
    compact = torch.cat([orig] + hard, 0)
    padded = torch.cat([orig] + [torch.cat([h, torch.zeros(b - h.shape[0], D, dtype=dtype)], 0) for h in hard], 0)
    scale = torch.tensor(25.0, device="cuda")
    res = []
    for Y, slab in ((compact, None), (padded, (torch.tensor(counts, dtype=torch.int32, device="cuda"), N0, b))):
        Xg = X.cuda().requires_grad_()
        Yg = Y.cuda().requires_grad_()
        loss = Fk.feat_row_ce(Xg, Yg, scale, 0.0, None, 5, slab)
        loss.backward()
        res.append((loss.item(), Xg.grad.float().cpu(), Yg.grad.float().cpu()))
    (l0, dx0, dy0), (l1, dx1, dy1) = res
    assert abs(l0 - l1) < 1e-4 * max(1.0, abs(l0))
    assert rel_l2(dx1, dx0) < 2e-3
    # gradient rows of the padded layout: filled rows match the compact layout, padding rows are exactly zero
    off = N0
    assert rel_l2(dy1[:N0], dy0[:N0]) < 2e-3
    for r, c in enumerate(counts):
        rows = dy1[N0 + r * b:N0 + (r + 1) * b]
        if c:
            assert rel_l2(rows[:c], dy0[off:off + c]) < 2e-3
        assert float(rows[c:].abs().max()) == 0.0 if c < b else True
        off += c
    # oracle cross-check of the compact problem
    Xo, Yo = X.float().requires_grad_(), compact.float().requires_grad_()
    lo = torch.nn.functional.cross_entropy(25.0 * Xo @ Yo.T, torch.arange(M) + 5)
    assert abs(l0 - lo.item()) < 2e-3 * max(1.0, abs(lo.item()))
