"""CPU: the oracle restatement (oracle/ref_oracle.py) must reproduce the committed outputs of the real reference
(tests/golden/goldens.pt, produced by oracle/make_golden.py from /root/reference)."""
import torch

from conftest import rel_l2
from oracle import ref_oracle as O


def test_g1_paired(goldens):
    for key, act in (("G1", "sigmoid"), ("G1b", "ones")):
        G = goldens[key]
        V = O.rn(1, 4, 196, 512).requires_grad_()
        T = O.rn(2, 4, 512).requires_grad_()
        img, txt = O.pacl_forward(V, T, act)
        loss = O.pacl_clip_loss(img, txt, 0.1)
        loss.backward()
        assert torch.allclose(O.patch_alignment(V, T).detach(), G["a"], atol=1e-6)
        assert torch.allclose(img.detach(), G["img"], atol=1e-6)
        assert abs(loss.item() - G["loss"].item()) < 2e-6
        assert rel_l2(V.grad, G["dV"]) < 1e-5 and rel_l2(T.grad, G["dT"]) < 1e-5


def test_g1_values_match_survey(goldens):
    # SURVEY.md Appendix B pins (survey-time run of the reference)
    assert abs(goldens["G1"]["loss"].item() - 0.5694192051887512) < 1e-6
    assert abs(goldens["G1b"]["loss"].item() - 1.2633897066116333) < 1e-6
    assert abs(goldens["G2"]["loss"].item() - 2.728726863861084) < 2e-6
    assert abs(goldens["G3"]["loss"].item() - 44.59129333496094) < 1e-4
    assert goldens["G4"]["top1"].tolist() == [3, 0, 3, 0, 0, 0, 1, 0, 3, 3, 2, 1, 1, 0, 2, 1]
    assert torch.allclose(goldens["G5"]["local"]["loss"], torch.tensor([6.744038, 9.403412]), atol=1e-5)
    assert torch.allclose(goldens["G5"]["global"]["loss"], torch.tensor([8.073725, 8.073725]), atol=1e-5)


def test_g2_sparc(goldens):
    G = goldens["G2"]
    B, T_, P, D = 4, 77, 196, 512
    V = O.rn(3, B, P, D).requires_grad_()
    L = O.rn(4, B, T_, D).requires_grad_()
    mask = (torch.arange(T_).expand(B, -1) <= G["eot"].unsqueeze(1)).float()
    v, lh, gh, _ = O.sparc_forward(V, L, mask, 1.0 / P)
    loss = O.sparc_loss(v, lh, gh, mask, 0.1)
    loss.backward()
    assert torch.allclose(gh.detach(), G["g_hat"], atol=1e-6)
    assert abs(loss.item() - G["loss"].item()) < 5e-6
    assert rel_l2(V.grad, G["dV"]) < 1e-4 and rel_l2(L.grad, G["dL"]) < 1e-4


def test_g3_openclip(goldens):
    G = goldens["G3"]
    img = O.l2n(O.rn(5, 8, 16)).requires_grad_()
    txt = O.l2n(O.rn(6, 11, 16)).requires_grad_()
    loss = O.openclip_loss_single(img, txt, 100.0, usehardtext=True)
    loss.backward()
    assert abs(loss.item() - G["loss"].item()) < 1e-4
    assert rel_l2(img.grad, G["dimg"]) < 1e-5 and rel_l2(txt.grad, G["dtxt"]) < 1e-5


def test_g4_eval_top1(goldens):
    V = O.rn(7, 16, 576, 768)
    T = O.rn(8, 16, 4, 768)
    top1, scores = O.eval_top1(V, T, 100.0)
    assert torch.equal(top1, goldens["G4"]["top1"])
    assert torch.allclose(scores, goldens["G4"]["scores"], atol=1e-4)


def test_g5_multirank_restatement(goldens):
    g = torch.Generator().manual_seed(0)
    all_img = O.l2n(torch.randn(8, 8, generator=g))
    all_txt = O.l2n(torch.randn(8, 8, generator=g))
    hard0 = O.l2n(torch.randn(1, 8, generator=g))
    hard1 = O.l2n(torch.randn(3, 8, generator=g))
    for key, local in (("local", True), ("global", False)):
        imgs = [all_img[:4].clone().requires_grad_(), all_img[4:].clone().requires_grad_()]
        txts = [torch.cat([all_txt[:4], hard0]).requires_grad_(), torch.cat([all_txt[4:], hard1]).requires_grad_()]
        losses = O.openclip_loss_ranks(imgs, txts, 10.0, local_loss=local, usehardtext=True)
        sum(losses).backward()
        assert torch.allclose(torch.stack([l.detach() for l in losses]), goldens["G5"][key]["loss"], atol=1e-5)
        for r in range(2):
            assert rel_l2(imgs[r].grad, goldens["G5"][key]["dimg"][r]) < 1e-5
            assert rel_l2(txts[r].grad, goldens["G5"][key]["dtxt"][r]) < 1e-5


def test_g6_allpairs(goldens):
    G = goldens["G6"]
    V = O.rn(11, 6, 50, 64).requires_grad_()
    T = O.rn(12, 6, 64).requires_grad_()
    loss = O.pacl_allpairs_loss(V, T, 0.1)
    loss.backward()
    assert abs(loss.item() - G["loss"].item()) < 2e-6
    assert rel_l2(V.grad, G["dV"]) < 1e-5 and rel_l2(T.grad, G["dT"]) < 1e-5


def test_g7_projection_heads():
    """Projection heads: the restatement reproduces the reference's own modules
    (tests/golden/goldens_heads.pt, made by `python oracle/make_golden.py heads`)."""
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)["G7"]
    sd = {k: v.clone().requires_grad_() for k, v in G["vis_sd"].items()}
    sdt = {k: v.clone().requires_grad_() for k, v in G["txt_sd"].items()}
    x = O.rn(21, 3, 50, 128).requires_grad_()
    t = O.rn(22, 5, 64).requires_grad_()
    y, ty = O.visual_projection(x, sd), O.text_projection(t, sdt)
    ((y * O.rn(23, 3, 50, 64)).sum() + (ty * O.rn(24, 5, 64)).sum()).backward()
    assert torch.allclose(y.detach(), G["y"], atol=2e-6) and torch.allclose(ty.detach(), G["ty"], atol=2e-6)
    assert rel_l2(x.grad, G["dx"]) < 1e-5 and rel_l2(t.grad, G["dt"]) < 1e-5
    for k, g in G["vis_grads"].items():
        assert rel_l2(sd[k].grad, g) < 1e-5, k
    for k, g in G["txt_grads"].items():
        assert rel_l2(sdt[k].grad, g) < 1e-5, k


def test_g8_rope():
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)["G8"]
    x = O.rn(25, 2, 50, 64).requires_grad_()
    y = O.apply_rope(x)
    (y * O.rn(26, 2, 50, 64)).sum().backward()
    assert torch.allclose(y.detach(), G["y"], atol=1e-7) and torch.allclose(x.grad, G["dx"], atol=1e-7)


def test_g9_g10_metrics_and_vlm2vec_loss():
    """The reference's own get_clip_metrics (open_clip_train/train.py:360-377, executed unmodified from its source) and
    VLM2Vec's SimpleContrastiveLoss (src/loss.py:7-19): the restatements reproduce their outputs."""
    import os
    G = torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)
    gq = torch.Generator().manual_seed(77)
    base = torch.randn(200, 32, generator=gq)
    gi = O.l2n(base + 2.0 * torch.randn(200, 32, generator=gq))
    gt = O.l2n(base + 2.0 * torch.randn(200, 32, generator=gq))
    m, _ = O.clip_metrics(gi, gt, 100.0)
    assert set(m) == set(G["G9"])
    for k, v in G["G9"].items():
        assert abs(float(m[k]) - v) < 1e-9, k
    x, y = O.l2n(O.rn(28, 6, 16)).requires_grad_(), O.l2n(O.rn(29, 18, 16)).requires_grad_()
    loss = O.simple_contrastive_loss(x, y, 0.02)
    loss.backward()
    assert abs(loss.item() - G["G10"]["loss"].item()) < 1e-5
    assert rel_l2(x.grad, G["G10"]["dx"]) < 1e-5 and rel_l2(y.grad, G["G10"]["dy"]) < 1e-5


def test_g2_sparc_scoring_local_and_global(goldens):
    """sparc.scoring (pacl.py:438-451), both modes, against the reference outputs stored with G2 (the generator's encoder
    stubs return the full V / L, so the oracle is evaluated on the same tensors)."""
    G = goldens["G2"]
    B, T_, P, D = 4, 77, 196, 512
    V, L = O.rn(3, B, P, D), O.rn(4, B, T_, D)
    mask = (torch.arange(T_).expand(B, -1) <= G["eot"].unsqueeze(1)).float()
    assert torch.allclose(O.sparc_scoring(V, L, mask, 1.0 / P, local=True), G["scoring_local"], atol=2e-6)
    assert torch.allclose(O.sparc_scoring(V, L, mask, 1.0 / P, local=False), G["scoring_global"], atol=2e-6)


def test_softmax_activation_independent_fp64_definition():
    """north_star (2) 'softmax over the patches': no reference implementation exists, so the oracle's softmax
    branch is checked against an INDEPENDENT fp64 statement of the definition written from Appendix A.1 alone --
    a_p = softmax_p(10 cos(t, v_p)) with its true denominator, u = sum_p a_p V_p, features = n(u), n(t) -- including the
    gradients (autograd through the fp64 definition)."""
    V = O.rn(61, 5, 37, 24).double().requires_grad_()
    T = O.rn(62, 5, 24).double().requires_grad_()
    vh = V / V.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    th = T / T.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    s = torch.einsum("bd,bpd->bp", th, vh)
    a = torch.exp(10.0 * s) / torch.exp(10.0 * s).sum(dim=-1, keepdim=True)          # plain softmax, written out
    u = (a.unsqueeze(-1) * V).sum(dim=1)
    img = u / u.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    w = O.rn(63, 5, 24).double()
    ((img * w).sum() + (th * w).sum()).backward()
    Vo, To = V.detach().float().requires_grad_(), T.detach().float().requires_grad_()
    io, to = O.pacl_forward(Vo, To, "softmax")
    ((io * w.float()).sum() + (to * w.float()).sum()).backward()
    assert torch.allclose(io.double(), img.detach(), atol=2e-6) and torch.allclose(to.double(), th.detach(), atol=2e-6)
    assert rel_l2(Vo.grad, V.grad) < 1e-5 and rel_l2(To.grad, T.grad) < 1e-5
