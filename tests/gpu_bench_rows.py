"""Per-row measurements of the hot path on one B200: our C-ABI path beside the eager-PyTorch
restatement of the reference (the oracle, run on the same GPU), with the §8d algorithmic bytes / flops.

    python tests/gpu_bench_rows.py [--json out.json]

Diagnostic companion of bench.py (which measures the BASELINE.json headline metric only); results are copied into
DESIGN.md / profiles/.  Timing: CUDA events around 7 back-to-back calls, 3 warm-ups, best of 3.
"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_oracle as O  # noqa: E402
import clip_embeds_b200.functional as Fk  # noqa: E402
from clip_embeds_b200 import losses  # noqa: E402

try:
    PEAKS = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    PEAKS = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def on_gpu(fn):
    """The oracle is a CPU restatement (its factory calls default to the CPU): run it with cuda as default device."""
    def run():
        with torch.device("cuda"):
            fn()
    return run


def timed(fn, warm=3, iters=7, reps=3):
    """ms per call: `iters` calls back to back between two CUDA events (as a training loop issues them, so that host
    launch overhead overlaps device work), best of `reps`."""
    for _ in range(warm):
        fn()
    best = float("inf")
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / iters)
    return best


def row(name, ours_ms, ref_ms, units, unit_name, gbytes=None, tflop=None):
    r = {"row": name, "ours_ms": ours_ms, "eager_torch_ms": ref_ms, "speedup_vs_eager": ref_ms / ours_ms if ref_ms else None,
         unit_name + "_per_s": units / (ours_ms * 1e-3)}
    if gbytes is not None:
        r["algorithmic_GB"] = gbytes
        r["achieved_GBps"] = gbytes / (ours_ms * 1e-3)
        r["frac_of_measured_hbm"] = r["achieved_GBps"] / PEAKS["hbm_gbs"]
    if tflop is not None:
        r["algorithmic_TFLOP"] = tflop
        r["achieved_TFLOPps"] = tflop / (ours_ms * 1e-3)
        r["frac_of_measured_bf16_burst"] = r["achieved_TFLOPps"] / PEAKS["bf16_tflops"]
    print(json.dumps(r), flush=True)
    return r


def main():
    dev = "cuda"
    out = []
    # ---- a1/a2 + a4: PACL paired forward() + ClipLoss, fwd+bwd, C2 shape, bf16 (reference training semantics)
    B, P, D = 1024, 576, 768
    V = O.rn(1, B, P, D).to(torch.bfloat16).to(dev).requires_grad_()
    T = O.rn(2, B, D).to(torch.bfloat16).to(dev).requires_grad_()
    crit = losses.ClipLoss(0.1)

    def ours_paired():
        V.grad = None
        T.grad = None
        img, txt = Fk.pacl_pool(V, T, "sigmoid")
        crit(img, txt).backward()

    Vr = V.detach().float().requires_grad_()
    Tr = T.detach().float().requires_grad_()

    def ref_paired():
        Vr.grad = None
        Tr.grad = None
        img, txt = O.pacl_forward(Vr, Tr, "sigmoid")
        O.pacl_clip_loss(img, txt, 0.1).backward()

    out.append(row("a1+a2+a4 PACL paired forward()+ClipLoss fwd+bwd, B=1024 P=576 D=768 bf16 (eager ref: fp32)",
                   timed(ours_paired), timed(on_gpu(ref_paired)), B, "pairs", gbytes=3 * B * P * D * 2 / 1e9))
    del Vr, Tr

    # ---- a3: eval scorer, What'sUp-shaped set (2247 items x 2 captions), fp32 inputs as the eval oracle
    items, K = 2247, 2
    Ve = O.rn(7, items, P, D).to(dev)
    Te = O.rn(8, items, K, D).to(dev)

    def ours_eval():
        Fk.pacl_eval_scores(Ve, Te, 100.0)

    def ref_eval():       # the reference protocol: one forward per item (eval_pacl.py:50-57)
        for i in range(0, items, 1):
            img, txt = O.pacl_forward(Ve[i:i + 1].expand(K, P, D), Te[i], "sigmoid")
            (100.0 * img @ txt.T).diagonal()

    out.append(row("a3 eval scorer, 2247 items x 2 captions, fp32 (eager ref: per-item loop)", timed(ours_eval),
                   timed(on_gpu(ref_eval), 1, 2, 2), items, "items", gbytes=items * P * D * 4 / 1e9))
    # ---- f2: the same set through the whole protocol (scores + accuracy accounting on the device)
    from clip_embeds_b200 import evalproto
    set_id = torch.arange(items) // 4
    rel_id = torch.arange(items) % 4

    def ours_proto():
        evalproto.whatsup_accuracies(Ve, Te, set_id, rel_id)

    def ref_proto():      # eval_pacl.py:38-104: per-item forward, host comparison (.item() sync per item), dict bookkeeping
        sc = []
        for i in range(items):
            img, txt = O.pacl_forward(Ve[i:i + 1].expand(K, P, D), Te[i], "sigmoid")
            pr = 100.0 * img @ txt.T
            sc.append([float(pr[0][0]), float(pr[1][1])])
        O.whatsup_accounting(torch.tensor(sc), set_id, rel_id)

    out.append(row("f2 What'sUp protocol end to end (scores + individual/pair/set accuracies), 2247 items x 2 captions, fp32 "
                   "(eager ref: per-item loop + host bookkeeping)", timed(ours_proto), timed(on_gpu(ref_proto), 1, 1, 2),
                   items, "items", gbytes=items * P * D * 4 / 1e9))
    del Ve, Te

    # ---- a4: PACL ClipLoss at the reference training batch (B=4096), bf16 features
    B4 = 4096
    x = torch.nn.functional.normalize(O.rn(5, B4, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    y = torch.nn.functional.normalize(O.rn(6, B4, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()

    def ours_ce():
        x.grad = None
        y.grad = None
        crit(x, y).backward()

    xr, yr = x.detach().float().requires_grad_(), y.detach().float().requires_grad_()

    def ref_ce():
        xr.grad = None
        yr.grad = None
        O.pacl_clip_loss(xr, yr, 0.1).backward()

    out.append(row("a4 PACL ClipLoss fwd+bwd, B=4096 D=768 bf16 (logits never written)", timed(ours_ce), timed(on_gpu(ref_ce)), B4,
                   "pairs", tflop=6.0 * B4 * B4 * D / 1e12))

    # ---- a5+a7: SPARC alignment + SparcLoss, C3 shape on one GPU (B=512, T=77)
    Bs, Tt = 512, 77
    Vs = O.rn(3, Bs, P, D).to(torch.bfloat16).to(dev).requires_grad_()
    Ls = O.rn(4, Bs, Tt, D).to(torch.bfloat16).to(dev).requires_grad_()
    eot = torch.randint(5, Tt, (Bs,), generator=torch.Generator().manual_seed(9))
    mask = (torch.arange(Tt)[None, :] <= eot[:, None]).float().to(dev)
    sl = losses.SparcLoss(0.1)

    def ours_sparc():
        Vs.grad = None
        Ls.grad = None
        l_hat, g_hat = Fk.sparc_align(Vs, Ls, 1.0 / P)
        sl(Vs, l_hat, g_hat, mask).backward()

    Vsr, Lsr = Vs.detach().float().requires_grad_(), Ls.detach().float().requires_grad_()

    def ref_sparc():
        Vsr.grad = None
        Lsr.grad = None
        v, l_hat, g_hat, m = O.sparc_forward(Vsr, Lsr, mask, 1.0 / P)
        O.sparc_loss(v, l_hat, g_hat, m, 0.1).backward()

    out.append(row("a5+a7 SPARC align + SparcLoss fwd+bwd, B=512 T=77 P=576 D=768 bf16 (eager ref: fp32)", timed(ours_sparc),
                   timed(on_gpu(ref_sparc)), Bs, "pairs", gbytes=(3 * Bs * P * D * 2 + 5 * Bs * Tt * D * 2) / 1e9))
    del Vsr, Lsr

    # ---- a8: NegCLIP ClipLoss with hard negatives, one rank's share of C4 at W=8: b=4096 local rows, N=32768 + 8192 hard
    b, N, H = 4096, 32768, 8192
    img_loc = torch.nn.functional.normalize(O.rn(5, b, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    txt_all = torch.nn.functional.normalize(O.rn(6, N + H, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    txt_loc = torch.nn.functional.normalize(O.rn(7, b + H // 8, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    img_all = torch.nn.functional.normalize(O.rn(8, N, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    lab_t = torch.full((b + H // 8,), -100, dtype=torch.int64, device=dev)
    lab_t[:b] = torch.arange(b, device=dev)

    def ours_neg():
        for t in (img_loc, txt_all, txt_loc, img_all):
            t.grad = None
        li = Fk.feat_row_ce(img_loc, txt_all, 100.0, 0.0, None, 0)
        lt = Fk.feat_row_ce(txt_loc, img_all, 100.0, 0.0, lab_t, 0)
        ((li + lt) / 2).backward()

    fl = [t.detach().float().requires_grad_() for t in (img_loc, txt_all, txt_loc, img_all)]

    def ref_neg():       # loss.py:156-164 local-loss branch: two GEMMs + two cross-entropies, logits materialised
        for t in fl:
            t.grad = None
        li = 100.0 * fl[0] @ fl[1].T
        lt = 100.0 * fl[2] @ fl[3].T
        loss = (torch.nn.functional.cross_entropy(li, torch.arange(b, device=dev)) +
                torch.nn.functional.cross_entropy(lt, lab_t, ignore_index=-100)) / 2
        loss.backward()

    flop = 6.0 * (b * (N + H) + (b + H // 8) * N) * D / 1e12
    out.append(row("a8 NegCLIP local-loss rank share (W=8 of C4): [4096 x 40960] + [5120 x 32768] logits, D=768, bf16 "
                   "(eager ref: fp32 TF32-off)", timed(ours_neg), timed(on_gpu(ref_neg)), b, "pairs", tflop=flop))
    del img_loc, txt_all, txt_loc, img_all, fl

    # ---- f1: visual projection head LayerNorm -> Patch_Projection(1024 -> 768), fwd+bwd, 1024 x 576 tokens
    import torch.nn as nn
    from clip_embeds_b200.heads import VisualProjection
    Bh, Din, Dout = 1024, 1024, 768
    torch.manual_seed(0)
    vis = VisualProjection(Din, Dout).to(dev).eval()
    xh = torch.randn(Bh, P, Din, device=dev).to(torch.bfloat16)
    gyh = torch.randn(Bh, P, Dout, device=dev).to(torch.bfloat16)

    def ours_heads():
        for p_ in vis.parameters():
            p_.grad = None
        vis(xh).backward(gyh)

    class _PP(nn.Module):          # the reference's Patch_Projection structure (pacl.py:35-48) in eager torch, bf16 weights
        def __init__(self):
            super().__init__()
            self.l = nn.Linear(Din, Dout)
            self.n = nn.Sequential(nn.Linear(Din, Dout), nn.GELU(), nn.Linear(Dout, Dout))

        def forward(self, z):
            return self.l(z) + self.n(z)

    refh = nn.Sequential(nn.LayerNorm(Din), nn.Dropout(0.1), _PP()).to(dev).to(torch.bfloat16).eval()

    def ref_heads():
        for p_ in refh.parameters():
            p_.grad = None
        refh(xh).backward(gyh)

    Rt = Bh * P
    flop_h = 2.0 * Rt * (2 * Din * Dout + Dout * Dout) * 3 / 1e12       # fwd + (dgrad + wgrad)
    out.append(row("f1 visual projection head (LayerNorm + Patch_Projection 1024->768) fwd+bwd, 1024x576 tokens, bf16 "
                   "(eager ref: bf16 cuBLAS)", timed(ours_heads), timed(ref_heads), Bh, "images", tflop=flop_h))
    if "--json" in sys.argv:
        with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
            json.dump({"peaks": PEAKS, "rows": out}, f, indent=1)


if __name__ == "__main__":
    main()
