"""GPU parity: SPARC alignment + SparcLoss vs the oracle / reference golden G2.

Tolerances: bf16 tensor-core operands (V, L, W, dS, dG are bf16) -> g_hat 5e-3 abs, loss 5e-3 abs,
gradients 4e-2 rel-L2 against the fp32 oracle on the same bf16-rounded inputs."""
import pytest
import torch

from conftest import rel_l2
from oracle import ref_oracle as O

pytestmark = pytest.mark.gpu


def _run(B, T, P, D, seed, eots):
    from clip_embeds_b200.losses import SparcLoss
    from clip_embeds_b200.models import SparcHead
    Vb = O.rn(seed, B, P, D).to(torch.bfloat16)
    Lb = O.rn(seed + 1, B, T, D).to(torch.bfloat16)
    eot = torch.as_tensor(eots)
    mask = (torch.arange(T).expand(B, -1) <= eot.unsqueeze(1)).float()
    Vo = Vb.float().requires_grad_()
    Lo = Lb.float().requires_grad_()
    v, lh, gh, _ = O.sparc_forward(Vo, Lo, mask, 1.0 / P)
    lo = O.sparc_loss(v, lh, gh, mask, 0.1)
    lo.backward()
    V = Vb.cuda().requires_grad_()
    L = Lb.cuda().requires_grad_()
    head = SparcHead(1.0 / P)
    v2, lh2, gh2, m2 = head(V, L, mask.cuda())
    loss = SparcLoss(0.1)(v2, lh2, gh2, m2)
    loss.backward()
    e_g = (gh2.detach().cpu() - gh.detach()).abs().max().item()
    e_l = (lh2.detach().cpu() - lh.detach()).abs().max().item()
    rv = rel_l2(V.grad.float().cpu(), Vo.grad)
    rl = rel_l2(L.grad.float().cpu(), Lo.grad)
    print(f"sparc B={B} T={T} P={P} D={D}: |g_hat|err={e_g:.2e} |l_hat|err={e_l:.2e} loss {loss.item():.5f} vs {lo.item():.5f} "
          f"relV={rv:.3e} relL={rl:.3e}")
    return e_g, e_l, abs(loss.item() - lo.item()), rv, rl


def test_sparc_g2_shape():
    e_g, e_l, dl, rv, rl = _run(4, 77, 196, 512, 3, [5, 10, 76, 20])
    assert e_g < 5e-3 and e_l < 1e-5 and dl < 5e-3 and rv < 4e-2 and rl < 4e-2


def test_sparc_vitl_shape():
    e_g, e_l, dl, rv, rl = _run(6, 77, 576, 768, 71, [5, 30, 76, 20, 9, 50])
    assert e_g < 5e-3 and e_l < 1e-5 and dl < 5e-3 and rv < 4e-2 and rl < 4e-2


@pytest.mark.parametrize("T,P,D", [(33, 625, 256), (16, 50, 128), (40, 700, 128)])
def test_sparc_ragged_shapes(T, P, D):
    """Patch counts that are not multiples of 32 / 64: ViT-B/16 at 400 px (625), a short row (50), a long one (700)."""
    e_g, e_l, dl, rv, rl = _run(3, T, P, D, 90 + T, [T - 1, T // 2, 3])
    assert e_g < 5e-3 and e_l < 1e-5 and dl < 5e-3 and rv < 4e-2 and rl < 4e-2


def test_sparc_golden_fp32_inputs(goldens):
    """Reference golden G2 (fp32 reference, fp32 inputs): input rounding to bf16 is part of the error."""
    from clip_embeds_b200.losses import SparcLoss
    from clip_embeds_b200.models import SparcHead
    G = goldens["G2"]
    B, T, P, D = 4, 77, 196, 512
    V = O.rn(3, B, P, D).cuda().requires_grad_()
    L = O.rn(4, B, T, D).cuda().requires_grad_()
    mask = (torch.arange(T).expand(B, -1) <= G["eot"].unsqueeze(1)).float().cuda()
    v, lh, gh, m = SparcHead(1.0 / P)(V, L, mask)
    loss = SparcLoss(0.1)(v, lh, gh, m)
    loss.backward()
    assert (gh.detach().cpu() - G["g_hat"]).abs().max().item() < 1e-2
    assert abs(loss.item() - G["loss"].item()) < 2e-2
    assert rel_l2(V.grad.cpu(), G["dV"]) < 8e-2
    assert rel_l2(L.grad.cpu(), G["dL"]) < 8e-2


def test_sparc_scoring():
    from clip_embeds_b200.models import SparcHead
    K, T, P, D = 4, 77, 196, 512
    Vb = O.rn(81, 1, P, D).to(torch.bfloat16)
    Lb = O.rn(82, K, T, D).to(torch.bfloat16)
    mask = torch.ones(K, T)
    for local in (False, True):
        ref = O.sparc_scoring(Vb.float(), Lb.float(), mask, 1.0 / P, local=local)
        out = SparcHead(1.0 / P).scoring(Vb.cuda(), Lb.cuda(), mask.cuda(), local=local)
        assert (out.cpu() - ref).abs().max().item() < 5e-3


def test_sparc_step_cuda_graph_replay_is_bit_identical():
    """The whole SPARC forward + backward (alignment, SparcLoss, every libclipk launch incl. programmatic dependent
    launches and host-encoded tensor maps) captured once with `GraphedStep` and replayed: same loss and gradients, bit
    for bit, as the eager step -- the library never synchronises or allocates, so a step is capturable."""
    from clip_embeds_b200.graphs import GraphedStep
    from clip_embeds_b200.losses import SparcLoss
    from clip_embeds_b200.models import SparcHead
    B, T, P, D = 8, 77, 196, 256
    V = O.rn(95, B, P, D).to(torch.bfloat16).cuda().requires_grad_()
    L = O.rn(96, B, T, D).to(torch.bfloat16).cuda().requires_grad_()
    mask = (torch.arange(T).expand(B, -1) <= torch.tensor([5, 76, 20, 9, 50, 30, 11, 70]).unsqueeze(1)).float().cuda()
    head, sl = SparcHead(1.0 / P), SparcLoss(0.1)

    def step():
        if V.grad is not None:
            V.grad.zero_()
            L.grad.zero_()
        v2, lh, gh, m2 = head(V, L, mask)
        loss = sl(v2, lh, gh, m2)
        loss.backward()
        return loss

    ref = step().item()
    gV, gL = V.grad.clone(), L.grad.clone()
    gs = GraphedStep(step)
    for _ in range(2):
        gs.replay()
    torch.cuda.synchronize()
    assert gs.loss.item() == ref
    assert torch.equal(V.grad, gV) and torch.equal(L.grad, gL)


def test_c3_full_size_vs_oracle():
    """BASELINE.json configs[2] AT ITS STATED SIZE: SPARC alignment + SparcLoss, B = 512, T = 77, P = 576, D = 768 (bf16
    inputs), forward + backward against the fp32 oracle on the same bf16-rounded inputs (a few seconds of CPU)."""
    g = torch.Generator().manual_seed(3)
    eots = torch.randint(5, 77, (512,), generator=g).tolist()
    e_g, e_l, dl, rv, rl = _run(512, 77, 576, 768, 301, eots)
    assert e_g < 5e-3 and e_l < 1e-5 and dl < 5e-3 and rv < 4e-2 and rl < 4e-2
