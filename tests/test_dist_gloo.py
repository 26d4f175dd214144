"""CPU, world_size 2 over gloo: host-side logic of the sharded losses (gather order, label offsets, ignore rows,
column-LSE merge, reduce-scatter gradients).  The CUDA kernels are replaced by torch CPU stand-ins HERE ONLY
(monkeypatched inside the test processes) so that the distributed wiring can be checked without a GPU; results are
compared with the reference goldens (G5) and with the single-process oracle."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import ref_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _cpu_feat_row_ce(X, Y, scale, bias=0.0, labels=None, label_offset=0, slab=None):
    assert slab is None
    logits = float(scale) * X.float() @ Y.float().T + float(bias if bias is not None else 0.0)
    if labels is None:
        labels = torch.arange(X.shape[0]) + label_offset
    return F.cross_entropy(logits, labels, ignore_index=-100) if (labels < 0).any() else F.cross_entropy(logits, labels)


def _patch_cpu_kernels():
    import clip_embeds_b200.functional as Fk
    Fk.feat_row_ce = _cpu_feat_row_ce

    def k_rows(L, offset):
        lse = torch.logsumexp(L, dim=1)
        idx = torch.arange(L.shape[0]) + offset
        return lse, lse - L[torch.arange(L.shape[0]), idx]

    def k_cols(L):
        mx = L.max(dim=0).values
        return mx, torch.exp(L - mx).sum(dim=0)

    def k_grad(L, row_lse, col_lse, offset, w_row, w_col):
        hit = torch.zeros_like(L)
        hit[torch.arange(L.shape[0]), torch.arange(L.shape[0]) + offset] = 1.0
        return w_row * (torch.exp(L - row_lse[:, None]) - hit) + w_col * (torch.exp(L - col_lse[None, :]) - hit)

    def k_payload(L, row_lse, row_loss):
        mx, sm = k_cols(L)
        return torch.cat([mx, sm, row_loss.sum()[None], (row_lse - row_loss).sum()[None]])

    def k_merge(gathered, W, N):
        g = gathered.reshape(W, 2 * N + 2)
        gmax = g[:, :N].max(dim=0).values
        col_lse = gmax + torch.log((g[:, N:2 * N] * torch.exp(g[:, :N] - gmax)).sum(dim=0))
        loss = 0.5 / N * (g[:, 2 * N].sum() + col_lse.sum() - g[:, 2 * N + 1].sum())
        return col_lse, loss.reshape(1)

    Fk._k_ce_rows, Fk._k_ce_cols, Fk._k_ce_scores_grad = k_rows, k_cols, k_grad
    Fk._k_ce_payload, Fk._k_ce_merge = k_payload, k_merge
    Fk._need_cuda = lambda *a: None
    Fk.pacl_scores = lambda V, T, c=1.0, activation="sigmoid", group=None, v_sqnorm=None: O.pacl_allpairs_scores(V, T, c, activation)
    # SPARC host logic: CPU stand-ins for the row kernels and the local term
    Fk.normalize_rows = O.l2n
    Fk.mean_dim1 = lambda X: X.float().mean(dim=1)
    Fk.pooled_patch_mean = lambda V: V.float().mean(dim=1)

    def local_loss(g_hat, l_hat, mask, scale, mask_sum=None):
        # 1/2 [ masked_pairwise(g, l) + masked_pairwise(l, g) ] with the (possibly GLOBAL) mask count as denominator
        m = mask.reshape(-1)
        num = 0.0
        for a, b in ((g_hat, l_hat), (l_hat, g_hat)):
            B, T, _ = a.shape
            logits = torch.einsum("bmd,bnd->bmn", a, b) * scale + ((1.0 - mask) * (-1e8)).unsqueeze(1)
            tgt = torch.eye(T).unsqueeze(0).expand(B, -1, -1).reshape(B * T, -1)
            num = num + (F.cross_entropy(logits.reshape(B * T, -1), tgt, reduction="none") * m).sum()
        return num * 0.5 / (m.sum() if mask_sum is None else mask_sum)

    Fk.sparc_local_loss = local_loss


def _worker(rank, world, port, case, q):
    import sys
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    _patch_cpu_kernels()
    from clip_embeds_b200 import losses
    try:
        if case in ("openclip_local", "openclip_global"):
            g = torch.Generator().manual_seed(0)
            all_img = O.l2n(torch.randn(8, 8, generator=g))
            all_txt = O.l2n(torch.randn(8, 8, generator=g))
            hard = [O.l2n(torch.randn(1, 8, generator=g)), O.l2n(torch.randn(3, 8, generator=g))]
            img = all_img[rank * 4:(rank + 1) * 4].clone().requires_grad_()
            txt = torch.cat([all_txt[rank * 4:(rank + 1) * 4], hard[rank]]).requires_grad_()
            fn = losses.OpenClipLoss(local_loss=case == "openclip_local", gather_with_grad=True, rank=rank,
                                     world_size=world, usehardtext=True)
            loss = fn(img, txt, torch.tensor(10.0))
            loss.backward()
            q.put((rank, loss.item(), img.grad.tolist(), txt.grad.tolist()))
        elif case == "allpairs":
            Bi, P, D = 6, 20, 16
            V = O.rn(91, Bi, P, D)
            T = O.rn(92, Bi, D)
            b = Bi // world
            Vl = V[rank * b:(rank + 1) * b].clone().requires_grad_()
            Tl = T[rank * b:(rank + 1) * b].clone().requires_grad_()
            loss = losses.PaclAllPairsLoss(0.1, group=dist.group.WORLD)(Vl, Tl)
            loss.backward()
            q.put((rank, loss.item(), Vl.grad.tolist(), Tl.grad.tolist()))
        elif case == "allpairs_ddp":
            # a small shared head in front of the sharded loss, wrapped in DDP (mean-reduces parameter gradients):
            # grad_reduction="mean" must give the head the gradient of the GLOBAL loss
            from torch.nn.parallel import DistributedDataParallel as DDP
            Bi, P, D = 6, 20, 16
            V = O.rn(91, Bi, P, D)
            T = O.rn(92, Bi, D)
            b = Bi // world
            torch.manual_seed(5)
            head = DDP(torch.nn.Linear(D, D, bias=False))
            loss = losses.PaclAllPairsLoss(0.1, group=dist.group.WORLD, grad_reduction="mean")(
                head(V[rank * b:(rank + 1) * b]), T[rank * b:(rank + 1) * b])
            loss.backward()
            q.put((rank, loss.item(), head.module.weight.grad.tolist(), []))
        elif case == "sparc":
            B, T_, P, D = 6, 9, 12, 16
            V, L = O.rn(93, B, P, D), O.rn(94, B, T_, D)
            mask = (torch.arange(T_).expand(B, -1) <= torch.tensor([2, 8, 5, 0, 7, 3]).unsqueeze(1)).float()
            b = B // world
            sl = slice(rank * b, (rank + 1) * b)
            Vl, Ll = V[sl].clone().requires_grad_(), L[sl].clone().requires_grad_()
            v, lh, gh, m = O.sparc_forward(Vl, Ll, mask[sl], 1.0 / P)      # the alignment itself is per sample
            loss = losses.SparcLoss(0.1, group=dist.group.WORLD)(v, lh, gh, m)
            loss.backward()
            q.put((rank, loss.item(), Vl.grad.tolist(), Ll.grad.tolist()))
    finally:
        dist.barrier()
        dist.destroy_process_group()


def _spawn(case):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, case, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=240) for _ in ps], key=lambda x: x[0])
    [p.join(timeout=60) for p in ps]
    return res


@pytest.mark.parametrize("case,key", [("openclip_local", "local"), ("openclip_global", "global")])
def test_openclip_hardneg_two_ranks(goldens, case, key):
    res = _spawn(case)
    G = goldens["G5"][key]
    for r in range(2):
        assert abs(res[r][1] - G["loss"][r].item()) < 1e-5
        assert torch.allclose(torch.tensor(res[r][2]), G["dimg"][r], atol=1e-5)
        assert torch.allclose(torch.tensor(res[r][3]), G["dtxt"][r], atol=1e-5)


def test_allpairs_sharded_two_ranks():
    res = _spawn("allpairs")
    Bi, P, D = 6, 20, 16
    V = O.rn(91, Bi, P, D).requires_grad_()
    T = O.rn(92, Bi, D).requires_grad_()
    lo = O.pacl_allpairs_loss(V, T, 0.1)
    lo.backward()
    for r in range(2):
        assert abs(res[r][1] - lo.item()) < 1e-5                      # every rank returns the global loss
        assert torch.allclose(torch.tensor(res[r][2]), V.grad[r * 3:(r + 1) * 3], atol=1e-6)
        assert torch.allclose(torch.tensor(res[r][3]), T.grad[r * 3:(r + 1) * 3], atol=1e-6)


def test_sparc_sharded_two_ranks():
    """SparcLoss with `group` (sample-sharded): gathered global term, local term over the GLOBAL mask count, scalar
    all-reduce -- equals the single-process loss on the whole batch (the reference's DataParallel semantics,
    train_sparc.py:92-94), and every rank gets the gradient of the global loss for its samples."""
    res = _spawn("sparc")
    B, T_, P, D = 6, 9, 12, 16
    V, L = O.rn(93, B, P, D).requires_grad_(), O.rn(94, B, T_, D).requires_grad_()
    mask = (torch.arange(T_).expand(B, -1) <= torch.tensor([2, 8, 5, 0, 7, 3]).unsqueeze(1)).float()
    v, lh, gh, m = O.sparc_forward(V, L, mask, 1.0 / P)
    lo = O.sparc_loss(v, lh, gh, m, 0.1)
    lo.backward()
    for r in range(2):
        assert abs(res[r][1] - lo.item()) < 1e-5
        assert torch.allclose(torch.tensor(res[r][2]), V.grad[r * 3:(r + 1) * 3], atol=1e-5)
        assert torch.allclose(torch.tensor(res[r][3]), L.grad[r * 3:(r + 1) * 3], atol=1e-5)


def test_allpairs_ddp_mean_reduce_gives_global_gradient():
    """under torch DDP (parameter gradients are MEAN-reduced) the sharded loss must hand every rank
    world_size x its share, so that the head receives d(global loss)/d(weights) -- `grad_reduction="mean"`."""
    res = _spawn("allpairs_ddp")
    Bi, P, D = 6, 20, 16
    V = O.rn(91, Bi, P, D)
    T = O.rn(92, Bi, D)
    torch.manual_seed(5)
    head = torch.nn.Linear(D, D, bias=False)
    lo = O.pacl_allpairs_loss(head(V), T, 0.1)
    lo.backward()
    for r in range(2):
        assert abs(res[r][1] - lo.item()) < 1e-5
        assert torch.allclose(torch.tensor(res[r][2]), head.weight.grad, atol=1e-5)
