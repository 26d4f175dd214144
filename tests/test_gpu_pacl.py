"""GPU parity: PACL paired + all-pairs + eval scorer vs the oracle / committed reference goldens.

Tolerances (stated per north_star):
  fp32 paired path ............ loss 1e-5 rel, gradients 1e-4 rel-L2 (fp32 CUDA-core arithmetic)
  bf16 tensor-core paths ...... scores 2e-2 abs at logit scale 10, loss 5e-3 abs, gradients 3e-2 rel-L2, measured against
                                the fp32 oracle evaluated on the SAME bf16-rounded inputs
  eval top-1 .................. bit-exact indices (fp32 scorer) on the reference golden G4
"""
import pytest
import torch

from conftest import rel_l2
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _cuda():
    import clip_embeds_b200.functional as Fk
    from clip_embeds_b200 import losses
    return Fk, losses


@pytest.mark.parametrize("activation,key", [("sigmoid", "G1"), ("ones", "G1b")])
def test_paired_fp32_golden(goldens, activation, key):
    Fk, losses = _cuda()
    G = goldens[key]
    V = O.rn(1, 4, 196, 512).cuda().requires_grad_()
    T = O.rn(2, 4, 512).cuda().requires_grad_()
    a = Fk.patch_alignment(V.detach(), T.detach())
    assert torch.allclose(a.cpu(), G["a"], atol=2e-6)
    img, txt = Fk.pacl_pool(V, T, activation)
    assert torch.allclose(img.detach().cpu(), G["img"], atol=2e-6)
    assert torch.allclose(txt.detach().cpu(), G["txt"], atol=2e-6)
    loss = losses.ClipLoss(0.1)(img, txt)
    loss.backward()
    assert abs(loss.item() - G["loss"].item()) <= 1e-5 * max(1.0, abs(G["loss"].item()))
    assert rel_l2(V.grad.cpu(), G["dV"]) < 1e-4
    assert rel_l2(T.grad.cpu(), G["dT"]) < 1e-4


def test_paired_bf16_vs_oracle():
    Fk, losses = _cuda()
    B, P, D = 48, 576, 768
    Vb = O.rn(21, B, P, D).to(torch.bfloat16)
    Tb = O.rn(22, B, D).to(torch.bfloat16)
    # oracle on the same bf16-rounded inputs, fp32 arithmetic
    Vo = Vb.float().requires_grad_()
    To = Tb.float().requires_grad_()
    io, to = O.pacl_forward(Vo, To, "sigmoid")
    gi = O.rn(23, B, D)
    gt = O.rn(24, B, D)
    ((io * gi).sum() + (to * gt).sum()).backward()
    V = Vb.cuda().requires_grad_()
    T = Tb.cuda().requires_grad_()
    img, txt = Fk.pacl_pool(V, T, "sigmoid")
    ((img * gi.cuda()).sum() + (txt * gt.cuda()).sum()).backward()
    assert torch.allclose(img.detach().cpu(), io.detach(), atol=1e-5)
    assert rel_l2(V.grad.float().cpu(), Vo.grad) < 1e-2      # bf16 rounding of the stored gradient
    assert rel_l2(T.grad.float().cpu(), To.grad) < 1e-2


def test_eval_top1_golden(goldens):
    Fk, _ = _cuda()
    V = O.rn(7, 16, 576, 768).cuda()
    T = O.rn(8, 16, 4, 768).cuda()
    scores, top1 = Fk.pacl_eval_scores(V, T, 100.0)
    assert torch.equal(top1.cpu(), goldens["G4"]["top1"])          # bit-exact indices
    assert torch.allclose(scores.cpu(), goldens["G4"]["scores"], atol=2e-4)


def _allpairs_case(Bi, Bt, P, D, seed, group=None, activation="sigmoid", c=10.0):
    Fk, _ = _cuda()
    Vb = O.rn(seed, Bi, P, D).to(torch.bfloat16)
    Tb = O.rn(seed + 1, Bt, D).to(torch.bfloat16)
    gs = O.rn(seed + 2, Bi, Bt) / (Bi * Bt) ** 0.5
    Vo = Vb.float().requires_grad_()
    To = Tb.float().requires_grad_()
    so = O.pacl_allpairs_scores(Vo, To, c, activation)
    (so * gs).sum().backward()
    V = Vb.cuda().requires_grad_()
    T = Tb.cuda().requires_grad_()
    s = Fk.pacl_scores(V, T, c, activation, group)
    (s * gs.cuda()).sum().backward()
    err_s = (s.detach().cpu() - so.detach()).abs().max().item()
    rv = rel_l2(V.grad.float().cpu(), Vo.grad)
    rt = rel_l2(T.grad.float().cpu(), To.grad)
    print(f"allpairs Bi={Bi} Bt={Bt} P={P} D={D} g={group} act={activation}: |ds|max={err_s:.3e} relV={rv:.3e} relT={rt:.3e}")
    return err_s, rv, rt


@pytest.mark.parametrize("shape", [(6, 6, 50, 64), (8, 200, 196, 512), (5, 130, 576, 768), (3, 70, 625, 768)])
def test_allpairs_vs_oracle(shape):
    err_s, rv, rt = _allpairs_case(*shape, seed=31)
    assert err_s < 2e-2 and rv < 3e-2 and rt < 3e-2


def test_allpairs_groups_and_ones():
    e1 = _allpairs_case(7, 40, 100, 128, seed=41, group=2)
    e2 = _allpairs_case(7, 40, 100, 128, seed=41, group=7)
    assert max(e1) < 3e-2 and max(e2) < 3e-2
    e4 = _allpairs_case(7, 40, 100, 128, seed=41, group=(2, 3))      # 3 internal lanes
    assert max(e4) < 3e-2
    e3 = _allpairs_case(4, 40, 100, 128, seed=43, activation="ones")
    assert max(e3) < 3e-2


@pytest.mark.parametrize("sched", [(-4, 2), (-3, 3), (-16, 1), 0])
def test_allpairs_persistent_kernel(sched):
    """The dependency-driven persistent kernel (group <= 0: -group images per group, `lanes` groups in lock-step; 0 =
    automatic): ragged last group, more groups than scratch slots, N tails."""
    e1 = _allpairs_case(11, 130, 196, 512, seed=51, group=sched)
    assert e1[0] < 2e-2 and e1[1] < 3e-2 and e1[2] < 3e-2
    e2 = _allpairs_case(7, 40, 100, 128, seed=41, group=sched, activation="ones")
    assert max(e2) < 3e-2


def test_allpairs_dv_orientations(monkeypatch):
    """dV GEMM with rows = patches (CLIPK_AP_K6T=0) and transposed, rows = features (default where it wastes less)."""
    monkeypatch.setenv("CLIPK_AP_K6T", "0")
    e0 = _allpairs_case(5, 130, 576, 768, seed=31)
    monkeypatch.setenv("CLIPK_AP_K6T", "1")
    e1 = _allpairs_case(5, 130, 576, 768, seed=31)
    assert e0[1] < 3e-2 and e1[1] < 3e-2 and e0[2] < 3e-2 and e1[2] < 3e-2


def test_softmax_activation():
    """Third activation of SURVEY Appendix A.1 / north_star (2): softmax over patches of 10 * cos.  The kernels pool
    with the un-normalised exp weights (the L2-normalisation cancels the denominator); checked against the oracle's
    restatement of the definition (no reference branch exists for it)."""
    Fk, _ = _cuda()
    V = O.rn(61, 6, 100, 128)
    T = O.rn(62, 6, 128)
    Vo, To = V.clone().requires_grad_(), T.clone().requires_grad_()
    io, to = O.pacl_forward(Vo, To, "softmax")
    (io * to).sum().backward()
    Vg, Tg = V.cuda().requires_grad_(), T.cuda().requires_grad_()
    ig, tg = Fk.pacl_pool(Vg, Tg, "softmax")
    (ig * tg).sum().backward()
    assert torch.allclose(ig.cpu(), io, atol=2e-6) and torch.allclose(tg.cpu(), to, atol=2e-6)
    assert rel_l2(Vg.grad.cpu(), Vo.grad) < 1e-4 and rel_l2(Tg.grad.cpu(), To.grad) < 1e-4
    a = Fk.patch_alignment(V.cuda(), T.cuda(), "softmax")
    ao = torch.softmax(10.0 * torch.einsum("bd,bpd->bp", O.l2n(T), O.l2n(V)), dim=-1)
    assert torch.allclose(a.cpu(), ao, atol=1e-6)
    e = _allpairs_case(5, 40, 100, 128, seed=63, activation="softmax")
    assert e[0] < 3e-2 and e[1] < 4e-2 and e[2] < 4e-2
    e = _allpairs_case(5, 130, 576, 768, seed=64, activation="softmax", group=(-3, 2))
    assert e[0] < 3e-2 and e[1] < 4e-2 and e[2] < 4e-2


def test_allpairs_loss_golden(goldens):
    """Reference golden G6 (fp32 reference on fp32 inputs) vs the bf16 tensor-core path: bf16 input rounding is part of
    the error here, hence the looser loss tolerance."""
    Fk, losses = _cuda()
    G = goldens["G6"]
    V = O.rn(11, 6, 50, 64).cuda().requires_grad_()
    T = O.rn(12, 6, 64).cuda().requires_grad_()
    loss = losses.PaclAllPairsLoss(0.1)(V, T)
    loss.backward()
    assert abs(loss.item() - G["loss"].item()) < 3e-2
    assert rel_l2(V.grad.cpu(), G["dV"]) < 6e-2
    assert rel_l2(T.grad.cpu(), G["dT"]) < 6e-2
    # same bf16-rounded inputs through the oracle: tight check of the loss wiring
    Vo = V.detach().cpu().to(torch.bfloat16).float().requires_grad_()
    To = T.detach().cpu().to(torch.bfloat16).float().requires_grad_()
    lo = O.pacl_allpairs_loss(Vo, To, 0.1)
    lo.backward()
    assert abs(loss.item() - lo.item()) < 5e-3
    assert rel_l2(V.grad.cpu(), Vo.grad) < 3e-2
    assert rel_l2(T.grad.cpu(), To.grad) < 3e-2


def test_full_size_properties():
    """BASELINE.json configs[1] at full size (B 1024, P 576, D 768, bf16): too large for the per-image oracle loop,
    so the CUDA path is checked through size-independent properties:
      (a) the diagonal of the all-pairs score matrix equals c * cosine of the PAIRED path (an independent fp32 SIMT
          kernel) on the same inputs;
      (b) the staged path (128-image groups) and the persistent dependency-driven kernel agree on scores and gradients;
      (c) permuting the images permutes the score rows and the dV rows, and leaves dT unchanged;
      (d) the backward is linear in the upstream gradient."""
    Fk, _ = _cuda()
    B, P, D, c = 1024, 576, 768, 10.0
    g = torch.Generator().manual_seed(77)
    V = torch.randn(B, P, D, generator=g).to(torch.bfloat16).cuda()
    T = torch.randn(B, D, generator=g).to(torch.bfloat16).cuda()
    g1 = (torch.randn(B, B, generator=g) / B).cuda()
    g2 = (torch.randn(B, B, generator=g) / B).cuda()

    def run(Vx, Tx, up, group):
        Vr, Tr = Vx.detach().requires_grad_(), Tx.detach().requires_grad_()
        s = Fk.pacl_scores(Vr, Tr, c, "sigmoid", group)
        s.backward(up)
        return s.detach(), Vr.grad.float(), Tr.grad.float()

    s_st, dV_st, dT_st = run(V, T, g1, (128, 2))
    # (a) diagonal vs the paired fp32 kernel
    cos = Fk._paired_forward(V, T, 0, want_cos=True)[4]
    assert (s_st.diagonal() - c * cos).abs().max().item() < 2e-2
    # (b) staged vs persistent kernel
    s_mg, dV_mg, dT_mg = run(V, T, g1, (-16, 3))
    assert (s_st - s_mg).abs().max().item() < 2e-2
    assert rel_l2(dV_mg, dV_st) < 2e-2 and rel_l2(dT_mg, dT_st) < 2e-2
    del s_mg, dV_mg, dT_mg
    # (c) image permutation equivariance
    perm = torch.randperm(B, generator=g).cuda()
    s_p, dV_p, dT_p = run(V[perm].contiguous(), T, g1[perm].contiguous(), (128, 2))
    assert (s_p - s_st[perm]).abs().max().item() < 1e-3          # same tiles, same arithmetic: only atomics reorder
    assert rel_l2(dV_p, dV_st[perm]) < 2e-3 and rel_l2(dT_p, dT_st) < 2e-3
    del s_p, dV_p, dT_p
    # (d) linearity in the upstream gradient
    _, dV_2, dT_2 = run(V, T, g2, (128, 2))
    _, dV_12, dT_12 = run(V, T, g1 + g2, (128, 2))
    assert rel_l2(dV_12, dV_st + dV_2) < 2e-2 and rel_l2(dT_12, dT_st + dT_2) < 2e-2


def test_allpairs_recompute_path(monkeypatch):
    """Backward WITHOUT the saved pooled vectors (CLIPK_AP_SAVE_POOLED=0: flash-style recompute, 8 GEMM units) and the
    default backward from the saved bf16 pooled vectors (7 units) agree with the oracle and with each other."""
    monkeypatch.setenv("CLIPK_AP_SAVE_POOLED", "0")
    e0 = _allpairs_case(5, 130, 576, 768, seed=31)
    monkeypatch.setenv("CLIPK_AP_SAVE_POOLED", "1")
    e1 = _allpairs_case(5, 130, 576, 768, seed=31)
    assert e0[0] < 2e-2 and e0[1] < 3e-2 and e0[2] < 3e-2
    assert e1[0] < 2e-2 and e1[1] < 3e-2 and e1[2] < 3e-2
    for act in ("ones", "softmax"):
        e = _allpairs_case(6, 70, 196, 512, seed=35, activation=act)
        assert e[0] < 3e-2 and e[1] < 4e-2 and e[2] < 4e-2


def test_c2_full_size_vs_oracle():
    """BASELINE.json configs[1] AT ITS STATED SIZE (B 1024, P 576, D 768, bf16) against the oracle: the full
    1024 x 1024 forward + backward runs once on the GPU; the score rows and the dV slabs of 6 sampled images are compared
    with the oracle's per-image reference loop over ALL 1024 texts on the same bf16-rounded inputs (the oracle needs
    ~0.3 s per image), and so is the InfoNCE row loss of those images."""
    Fk, _ = _cuda()
    B, P, D, c = 1024, 576, 768, 10.0
    g = torch.Generator().manual_seed(91)
    Vb = torch.randn(B, P, D, generator=g).to(torch.bfloat16)
    Tb = torch.randn(B, D, generator=g).to(torch.bfloat16)
    up = torch.randn(B, B, generator=g) / B
    V = Vb.cuda().requires_grad_()
    T = Tb.cuda().requires_grad_()
    s = Fk.pacl_scores(V, T, c, "sigmoid")
    s.backward(up.cuda())
    idx = [0, 127, 128, 511, 640, 1023]            # group edges of the 128-image schedule included
    Vo = Vb[idx].float().requires_grad_()
    To = Tb.float()
    so = O.pacl_allpairs_scores(Vo, To, c)
    (so * up[idx]).sum().backward()
    err = (s.detach().cpu()[idx] - so.detach()).abs().max().item()
    rv = rel_l2(V.grad.float().cpu()[idx], Vo.grad)
    # row-wise InfoNCE term of the sampled images (image -> text direction)
    lab = torch.tensor(idx)
    ce_g = torch.nn.functional.cross_entropy(s.detach().cpu()[idx], lab, reduction="none")
    ce_o = torch.nn.functional.cross_entropy(so.detach(), lab, reduction="none")
    print(f"C2 full size vs oracle: |dscore|max={err:.3e} rel dV={rv:.3e} |dCE|max={(ce_g - ce_o).abs().max().item():.3e}")
    assert err < 2e-2 and rv < 3e-2
    assert (ce_g - ce_o).abs().max().item() < 5e-3


def test_c2_dT_vs_oracle_256():
    """dT sums over every image, so it is checked against the oracle on a full (smaller) problem at the C2 patch / feature
    shape: Bi = Bt = 256, P 576, D 768 (two 128-image groups on two lanes), loss = all-pairs InfoNCE."""
    Fk, losses = _cuda()
    B, P, D = 256, 576, 768
    Vb = O.rn(93, B, P, D).to(torch.bfloat16)
    Tb = O.rn(94, B, D).to(torch.bfloat16)
    V = Vb.cuda().requires_grad_()
    T = Tb.cuda().requires_grad_()
    loss = losses.PaclAllPairsLoss(0.1)(V, T)
    loss.backward()
    Vo = Vb.float().requires_grad_()
    To = Tb.float().requires_grad_()
    lo = O.pacl_allpairs_loss(Vo, To, 0.1)
    lo.backward()
    rv = rel_l2(V.grad.float().cpu(), Vo.grad)
    rt = rel_l2(T.grad.float().cpu(), To.grad)
    print(f"C2-shape 256x256 vs oracle: loss {loss.item():.6f} / {lo.item():.6f}  rel dV={rv:.3e} rel dT={rt:.3e}")
    assert abs(loss.item() - lo.item()) < 5e-3 and rv < 3e-2 and rt < 3e-2


def test_c1_paired_cliploss_full_size():
    """BASELINE.json configs[0] exactly: PACL paired forward + ClipLoss(0.1), V [64,196,512], T [64,512], fp32,
    seeds 1 / 2 -- CUDA path vs the oracle at the stated size."""
    Fk, losses = _cuda()
    V0, T0 = O.rn(1, 64, 196, 512), O.rn(2, 64, 512)
    Vo, To = V0.clone().requires_grad_(), T0.clone().requires_grad_()
    io, to = O.pacl_forward(Vo, To, "sigmoid")
    lo = O.pacl_clip_loss(io, to, 0.1)
    lo.backward()
    V, T = V0.cuda().requires_grad_(), T0.cuda().requires_grad_()
    img, txt = Fk.pacl_pool(V, T, "sigmoid")
    loss = losses.ClipLoss(0.1)(img, txt)
    loss.backward()
    assert torch.allclose(img.detach().cpu(), io.detach(), atol=2e-6)
    assert abs(loss.item() - lo.item()) <= 1e-5 * max(1.0, abs(lo.item()))
    assert rel_l2(V.grad.cpu(), Vo.grad) < 1e-4 and rel_l2(T.grad.cpu(), To.grad) < 1e-4


def test_fp16_inputs_paired_and_eval_top1():
    """fp16 features (the reference evaluates under fp16 autocast, eval_pacl.py:53; NegCLIP trains with --precision amp):
    the paired / eval-scorer kernels take IEEE half directly (fp32 arithmetic inside).  (a) forward + backward against
    the fp32 oracle on the same fp16-rounded inputs; (b) eval top-1 on What'sUp-shaped items: identical indices to the
    fp32 oracle evaluated on the fp16-rounded inputs (bit-exact), and agreement with the all-fp32 pipeline reported
    together with the smallest top-1 margin (items whose margin is below the fp16 input rounding may legitimately flip)."""
    Fk, _ = _cuda()
    B, P, D = 24, 576, 768
    Vh, Th = O.rn(401, B, P, D).half(), O.rn(402, B, D).half()
    Vo, To = Vh.float().requires_grad_(), Th.float().requires_grad_()
    io, to = O.pacl_forward(Vo, To, "sigmoid")
    gi, gt = O.rn(403, B, D), O.rn(404, B, D)
    ((io * gi).sum() + (to * gt).sum()).backward()
    V, T = Vh.cuda().requires_grad_(), Th.cuda().requires_grad_()
    img, txt = Fk.pacl_pool(V, T, "sigmoid")
    ((img * gi.cuda()).sum() + (txt * gt.cuda()).sum()).backward()
    assert V.grad.dtype == torch.float16 and T.grad.dtype == torch.float16
    assert torch.allclose(img.detach().cpu(), io.detach(), atol=1e-5) and torch.allclose(txt.detach().cpu(), to.detach(), atol=1e-5)
    assert rel_l2(V.grad.float().cpu(), Vo.grad) < 2e-3 and rel_l2(T.grad.float().cpu(), To.grad) < 2e-3     # fp16 rounding of the stored gradient
    # (b) eval scorer
    items, K = 200, 4
    Vf, Tf = O.rn(411, items, P, D), O.rn(412, items, K, D)
    top_h, sc_h = O.eval_top1(Vf.half().float(), Tf.half().float(), 100.0)
    top_f, sc_f = O.eval_top1(Vf, Tf, 100.0)
    scores, top1 = Fk.pacl_eval_scores(Vf.half().cuda(), Tf.half().cuda(), 100.0)
    assert torch.equal(top1.cpu(), top_h)                                     # bit-exact indices on identical inputs
    assert torch.allclose(scores.cpu(), sc_h, atol=2e-4)
    srt = sc_f.sort(dim=1, descending=True).values
    margin = (srt[:, 0] - srt[:, 1])
    agree = (top1.cpu() == top_f)
    print(f"fp16 eval: agreement with the all-fp32 pipeline {agree.float().mean().item():.3f}, min margin {margin.min().item():.2e}, "
          f"min margin among agreeing items {margin[agree].min().item():.2e}")
    assert bool(agree[margin > 5e-2].all())          # flips only where the fp32 margin is within the input rounding
