"""Feature-CE probe (diagnostic): NegCLIP rank share of C4 at W=8 and the PACL ClipLoss at B=4096, fwd+bwd.
CLIPK_CE_ENGINE=1|2 selects the single-CTA / CTA-pair engine."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
import clip_embeds_b200.functional as Fk  # noqa: E402
from clip_embeds_b200 import losses  # noqa: E402

dev = "cuda"
D = 768
nrm = torch.nn.functional.normalize


def timed(fn, warm=3, iters=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


b, N, H = 4096, 32768, 8192
img_loc = nrm(O.rn(5, b, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
txt_all = nrm(O.rn(6, N + H, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
txt_loc = nrm(O.rn(7, b + H // 8, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
img_all = nrm(O.rn(8, N, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
lab_t = torch.full((b + H // 8,), -100, dtype=torch.int64, device=dev)
lab_t[:b] = torch.arange(b, device=dev)


def neg():
    for t in (img_loc, txt_all, txt_loc, img_all):
        t.grad = None
    li = Fk.feat_row_ce(img_loc, txt_all, 100.0, 0.0, None, 0)
    lt = Fk.feat_row_ce(txt_loc, img_all, 100.0, 0.0, lab_t, 0)
    ((li + lt) / 2).backward()


def neg_fwd():
    with torch.no_grad():
        Fk.feat_row_ce(img_loc, txt_all, 100.0, 0.0, None, 0)
        Fk.feat_row_ce(txt_loc, img_all, 100.0, 0.0, lab_t, 0)


flop = 6.0 * (b * (N + H) + (b + H // 8) * N) * D
if len(sys.argv) > 1 and sys.argv[1] == "once":          # for ncu: one warm step, one profiled step
    neg()
    torch.cuda.synchronize()
    neg()
    torch.cuda.synchronize()
    sys.exit(0)
t = timed(neg)
tf = timed(neg_fwd)
print(f"engine={os.environ.get('CLIPK_CE_ENGINE', 'auto')} negclip share fwd+bwd {t:.3f} ms ({flop / t / 1e9:.0f} TF/s algorithmic), "
      f"fwd only {tf:.3f} ms ({flop / 3 / tf / 1e9:.0f} TF/s)", flush=True)

B4 = 4096
x = nrm(O.rn(5, B4, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
y = nrm(O.rn(6, B4, D), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
crit = losses.ClipLoss(0.1)


def ce():
    x.grad = None
    y.grad = None
    crit(x, y).backward()


t = timed(ce)
print(f"engine={os.environ.get('CLIPK_CE_ENGINE', 'auto')} ClipLoss B=4096 fwd+bwd {t:.3f} ms ({6.0 * B4 * B4 * D / t / 1e9:.0f} TF/s)", flush=True)
