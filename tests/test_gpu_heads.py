"""GPU parity: projection heads (PACL/model/pacl.py:35-48, :70-79) through the C ABI vs the oracle
and the reference's own outputs (tests/golden/goldens_heads.pt).

Tolerances (bf16 tensor-core path against the fp32 oracle): outputs 2e-2 relative L2 (bf16 activations and weights,
two chained GEMMs), input / weight / bias gradients 3e-2 relative L2; LayerNorm statistics are fp32."""
import os

import pytest
import torch

from conftest import rel_l2
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _load_g7():
    return torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)["G7"]


def test_heads_state_dict_keys_match_reference():
    from clip_embeds_b200.heads import VisualProjection, TextProjection
    G = _load_g7()
    vis, txt = VisualProjection(128, 64), TextProjection(64)
    assert sorted(vis.state_dict().keys()) == sorted(G["vis_sd"].keys())
    assert sorted(txt.state_dict().keys()) == sorted(G["txt_sd"].keys())
    vis.load_state_dict(G["vis_sd"])
    txt.load_state_dict(G["txt_sd"])
    assert vis[2].linear_projection[0].weight is vis[2].text_projection[0].weight      # one module, two names (pacl.py:39)


def test_heads_golden_small():
    """The reference's own outputs and gradients (eval mode, Din=128, Dout=64, 150 tokens / 5 texts)."""
    from clip_embeds_b200.heads import VisualProjection, TextProjection
    G = _load_g7()
    vis, txt = VisualProjection(128, 64).cuda().eval(), TextProjection(64).cuda().eval()
    vis.load_state_dict(G["vis_sd"])
    txt.load_state_dict(G["txt_sd"])
    x = O.rn(21, 3, 50, 128).cuda().requires_grad_()
    t = O.rn(22, 5, 64).cuda().requires_grad_()
    y, ty = vis(x), txt(t)
    assert y.dtype == torch.bfloat16 and y.shape == (3, 50, 64)
    ((y.float() * O.rn(23, 3, 50, 64).cuda()).sum() + (ty.float() * O.rn(24, 5, 64).cuda()).sum()).backward()
    print(f"heads golden: y {rel_l2(y.float().cpu(), G['y']):.2e} ty {rel_l2(ty.float().cpu(), G['ty']):.2e} "
          f"dx {rel_l2(x.grad.cpu(), G['dx']):.2e} dt {rel_l2(t.grad.cpu(), G['dt']):.2e}")
    assert rel_l2(y.float().cpu(), G["y"]) < 2e-2 and rel_l2(ty.float().cpu(), G["ty"]) < 2e-2
    assert rel_l2(x.grad.cpu(), G["dx"]) < 3e-2 and rel_l2(t.grad.cpu(), G["dt"]) < 3e-2
    for k, p in vis.named_parameters():
        print(f"  vis {k}: {rel_l2(p.grad.cpu(), G['vis_grads'][k]):.2e}")
        assert rel_l2(p.grad.cpu(), G["vis_grads"][k]) < 3e-2, k
    for k, p in txt.named_parameters():
        assert rel_l2(p.grad.cpu(), G["txt_grads"][k]) < 3e-2, k


@pytest.mark.parametrize("shape", [(2, 576, 1024, 768), (3, 196, 768, 512), (1, 257, 1408, 1024)])
def test_visual_projection_vs_oracle(shape):
    """ViT-L/14-336 (1024 -> 768), ViT-B/16 (768 -> 512) and EVA01-g (1408 -> 1024) head shapes, seeded weights, against
    the fp32 oracle evaluated on the same fp32 inputs."""
    from clip_embeds_b200.heads import VisualProjection
    B, P, Din, Dout = shape
    torch.manual_seed(7)
    vis = VisualProjection(Din, Dout).eval()
    with torch.no_grad():
        vis[0].weight.copy_(1.0 + 0.1 * O.rn(31, Din))
        vis[0].bias.copy_(0.1 * O.rn(32, Din))
    sd = {k: v.detach().clone().requires_grad_() for k, v in vis.state_dict().items()}
    x = O.rn(41, B, P, Din)
    gy = O.rn(42, B, P, Dout)
    xo = x.clone().requires_grad_()
    yo = O.visual_projection(xo, sd)
    (yo * gy).sum().backward()
    vis = vis.cuda()
    xg = x.cuda().requires_grad_()
    y = vis(xg)
    (y.float() * gy.cuda()).sum().backward()
    errs = {"y": rel_l2(y.float().cpu(), yo.detach()), "dx": rel_l2(xg.grad.cpu(), xo.grad)}
    for k, p in vis.named_parameters():
        errs[k] = rel_l2(p.grad.cpu(), sd[k].grad)
    print(f"visual projection {shape}: " + " ".join(f"{k}={v:.2e}" for k, v in errs.items()))
    assert errs["y"] < 2e-2
    for k, v in errs.items():
        assert v < 3e-2, (k, v)


def test_heads_feed_scorer_end_to_end():
    """heads -> all-pairs scorer -> InfoNCE, gradients reach every head parameter (the PACL training graph,
    pacl.py:135-145 with the north_star all-pairs scores)."""
    from clip_embeds_b200.heads import VisualProjection, TextProjection
    from clip_embeds_b200.losses import PaclAllPairsLoss
    torch.manual_seed(3)
    B, P = 8, 64
    vis, txt = VisualProjection(256, 128).cuda().eval(), TextProjection(128).cuda().eval()
    patches = O.rn(51, B, P, 256).cuda()
    text_cls = O.rn(52, B, 128).cuda()
    loss = PaclAllPairsLoss(0.1)(vis(patches), txt(text_cls))
    loss.backward()
    assert torch.isfinite(loss)
    for p in list(vis.parameters()) + list(txt.parameters()):
        assert p.grad is not None and torch.isfinite(p.grad).all() and p.grad.abs().max() > 0


def test_rope_matches_reference_golden():
    """apply_rope (pacl.py:147-181): fp32 path against the reference's own output and gradient (golden G8; the angle
    tables are built with the reference's expressions, so only the two products per element can differ: <= 1e-6)."""
    from clip_embeds_b200.heads import apply_rope
    G8 = torch.load(os.path.join(os.path.dirname(__file__), "golden", "goldens_heads.pt"), weights_only=True)["G8"]
    x = O.rn(25, 2, 50, 64).cuda().requires_grad_()
    y = apply_rope(x)
    (y * O.rn(26, 2, 50, 64).cuda()).sum().backward()
    assert (y.detach().cpu() - G8["y"]).abs().max().item() <= 1e-6
    assert (x.grad.cpu() - G8["dx"]).abs().max().item() <= 1e-6
    # bf16 in / bf16 out at the ViT-L/14-336 token shape against the oracle on the same bf16 inputs
    xb = O.rn(27, 3, 576, 1024).to(torch.bfloat16)
    yb = apply_rope(xb.cuda())
    assert yb.dtype == torch.bfloat16
    assert rel_l2(yb.float().cpu(), O.apply_rope(xb.float())) < 4e-3


def test_llm2pacl_c5_heads_to_eval_scorer():
    """BASELINE configs[4] shape (LLM2PACL eval): 576 x 1024 patch tokens through LN + Patch_Projection(1024, 768), K
    precomputed 4096-d LLM text embeddings through LN + Linear(4096, 768), then the one-launch eval scorer.
    (1) given OUR head outputs, the scorer's top-1 is bit-exact against the fp32 oracle scorer on the same tensors;
    (2) the heads themselves match the fp32 oracle heads to the bf16 tolerance; (3) end to end the top-1 agrees with the
    all-fp32 oracle pipeline on at least 90 % of the items (bf16 heads can flip near-ties)."""
    import clip_embeds_b200.functional as Fk
    from clip_embeds_b200.heads import VisualProjection, TextProjection
    items, K, P = 24, 4, 576
    torch.manual_seed(11)
    vis, txt = VisualProjection(1024, 768).eval(), TextProjection(4096, 768).eval()
    sdv = {k: v.detach().clone() for k, v in vis.state_dict().items()}
    sdt = {k: v.detach().clone() for k, v in txt.state_dict().items()}
    patches, llm = O.rn(101, items, P, 1024), O.rn(102, items * K, 4096)
    with torch.no_grad():
        Vo, To = O.visual_projection(patches, sdv), O.text_projection(llm, sdt).reshape(items, K, 768)
        top_o, _ = O.eval_top1(Vo, To, 100.0)
        V = vis.cuda()(patches.cuda())
        T = txt.cuda()(llm.cuda()).reshape(items, K, 768)
        scores, top1 = Fk.pacl_eval_scores(V.float(), T.float())
        top_same_inputs, _ = O.eval_top1(V.float().cpu(), T.float().cpu(), 100.0)
    assert torch.equal(top1.cpu(), top_same_inputs)                                        # (1)
    assert rel_l2(V.float().cpu(), Vo) < 2e-2 and rel_l2(T.float().cpu(), To) < 2e-2       # (2)
    agree = (top1.cpu() == top_o).float().mean().item()
    print(f"C5 heads -> eval scorer: top-1 agreement with the all-fp32 oracle pipeline {agree:.3f}")
    assert agree >= 0.9                                                                    # (3)


def test_training_mode_dropout_fused_mask_replay():
    """Training mode (pacl.py:72: Dropout(0.1) between LayerNorm and Patch_Projection): the dropout mask is drawn inside
    the LayerNorm kernel and kept as one bit per element.  (a) the keep rate is 1 - p within 4 sigma and differs between
    seeds; (b) MASK REPLAY: the same mask applied in the oracle's visual_projection reproduces the CUDA output, the input
    gradient and every parameter gradient within the bf16 tolerances of the eval-mode test."""
    from clip_embeds_b200 import heads
    from oracle import ref_oracle as O
    torch.manual_seed(0)
    B, S, Din, Dout, p = 4, 50, 256, 128, 0.1
    mod = heads.VisualProjection(Din, Dout, p).cuda().train()
    x0 = O.rn(501, B, S, Din)
    x = x0.cuda().requires_grad_()
    ln = mod[0]
    xn, keep = heads.layer_norm_bf16(x, ln.weight, ln.bias, ln.eps, p, 1234, return_mask=True)
    mask = heads.unpack_keep_bits(keep, (B, S, Din))
    rate = mask.float().mean().item()
    n = mask.numel()
    assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / n) ** 0.5 + 1e-4, rate
    _, keep2 = heads.layer_norm_bf16(x, ln.weight, ln.bias, ln.eps, p, 99, return_mask=True)
    assert not torch.equal(keep, keep2)
    _, keep3 = heads.layer_norm_bf16(x, ln.weight, ln.bias, ln.eps, p, 1234, return_mask=True)
    assert torch.equal(keep, keep3)                      # the mask is a pure function of (seed, element index)
    # forward + backward through the fused path with that mask
    y = mod[2](xn)
    gy = O.rn(502, B, S, Dout).cuda()
    (y.float() * gy).sum().backward()
    # oracle with the replayed mask (fp32): LayerNorm -> mask / (1 - p_eff) -> Patch_Projection
    sd = {k: v.detach().cpu().float() for k, v in mod.state_dict().items()}
    xo = x0.clone().requires_grad_()
    params = {k: v.clone().requires_grad_() for k, v in sd.items()}
    thr = int(p * 65536 + 0.5)
    scale = 65536.0 / (65536.0 - thr)
    xno = O.layer_norm(xo, params["0.weight"], params["0.bias"]) * mask.cpu().float() * scale
    yo = O.patch_projection(xno, params, prefix="2.")
    (yo * gy.cpu()).sum().backward()
    assert (y.float().cpu() - yo.detach()).abs().max().item() < 3e-2 * max(1.0, yo.detach().abs().max().item())
    assert rel_l2(x.grad.cpu(), xo.grad) < 3e-2
    for k, prm in mod.named_parameters():
        if prm.grad is not None and k in params and params[k].grad is not None:
            assert rel_l2(prm.grad.cpu().float(), params[k].grad) < 4e-2, k
    # the module itself in train mode: different calls draw different masks, eval mode is deterministic
    y1, y2 = mod(x.detach()), mod(x.detach())
    assert not torch.equal(y1, y2)
    mod.eval()
    assert torch.equal(mod(x.detach()), mod(x.detach()))


def test_head_emits_row_norms_for_the_scorer():
    """§8f-1: the head's output GEMM emits the squared norm of every projected patch row (of the bf16-rounded values);
    the all-pairs scorer fed with them skips its own norm pass and returns the same scores and gradients."""
    import clip_embeds_b200.functional as Fk
    from clip_embeds_b200 import heads
    torch.manual_seed(1)
    B, S, Din, Dout = 6, 196, 256, 128
    vis = heads.VisualProjection(Din, Dout).cuda().eval()
    x = O.rn(601, B, S, Din).cuda()
    T = O.rn(602, 40, Dout).cuda().to(torch.bfloat16).requires_grad_()
    V, sq = vis(x, return_sqnorm=True)
    want = V.detach().float().pow(2).sum(-1)
    assert sq.shape == (B, S) and torch.allclose(sq, want, rtol=1e-5, atol=1e-6)
    Va = V.detach().clone().requires_grad_()
    Vb = V.detach().clone().requires_grad_()
    Tb = T.detach().clone().requires_grad_()
    up = O.rn(603, B, 40).cuda() / 40
    s0 = Fk.pacl_scores(Va, T, 10.0, "sigmoid")
    s0.backward(up)
    s1 = Fk.pacl_scores(Vb, Tb, 10.0, "sigmoid", None, sq)
    s1.backward(up)
    assert (s0 - s1).abs().max().item() < 1e-3
    assert rel_l2(Vb.grad.float(), Va.grad.float()) < 1e-3 and rel_l2(Tb.grad.float(), T.grad.float()) < 1e-3
