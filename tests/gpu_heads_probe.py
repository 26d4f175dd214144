"""Projection-head probe (diagnostic): VisualProjection (LayerNorm -> Patch_Projection 1024 -> 768) forward and
forward+backward at the C2 token count (1024 images x 576 patches), beside eager torch (the reference's module in bf16).
`once` runs one warm + one measured step for ncu."""
import os
import sys

import torch
import torch.nn as nn

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200.heads import VisualProjection  # noqa: E402

B = int(os.environ.get("HEADS_B", "1024"))
P, Din, Dout = 576, 1024, 768
dev = "cuda"
torch.manual_seed(0)
vis = VisualProjection(Din, Dout).to(dev).eval()
x = torch.randn(B, P, Din, device=dev).to(torch.bfloat16)
gy = torch.randn(B, P, Dout, device=dev).to(torch.bfloat16)


def fwd():
    with torch.no_grad():
        return vis(x)


def fb():
    for p in vis.parameters():
        p.grad = None
    vis(x).backward(gy)


def timed(fn, warm=2, iters=5):
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


if len(sys.argv) > 1 and sys.argv[1] == "once":
    fb()
    torch.cuda.synchronize()
    fb()
    torch.cuda.synchronize()
    sys.exit(0)

R = B * P
f_fwd = 2.0 * R * (2 * Din * Dout + Dout * Dout)
# bwd: dH, dxn (2 pairs), dW1, dW2, dW3
f_bwd = 2.0 * R * (Dout * Dout + 2 * Dout * Din + 2 * Dout * Din + Dout * Dout)
tf, tb = timed(fwd), timed(fb)
print(f"heads B={B}: fwd {tf:.3f} ms ({f_fwd / tf / 1e9:.0f} TF/s), fwd+bwd {tb:.3f} ms ({(f_fwd + f_bwd) / tb / 1e9:.0f} TF/s), "
      f"{B / tb * 1e3:.0f} images/s", flush=True)

# eager torch: the reference's module structure under bf16 parameters (generous baseline: the reference trains in fp32)
ref = nn.Sequential(nn.LayerNorm(Din), nn.Dropout(0.1), nn.Identity()).to(dev)


class PP(nn.Module):
    def __init__(self):
        super().__init__()
        self.l = nn.Linear(Din, Dout)
        self.n = nn.Sequential(nn.Linear(Din, Dout), nn.GELU(), nn.Linear(Dout, Dout))

    def forward(self, z):
        return self.l(z) + self.n(z)


ref[2] = PP().to(dev)
ref = ref.to(torch.bfloat16).eval()


def ref_fb():
    for p in ref.parameters():
        p.grad = None
    ref(x).backward(gy)


def ref_fwd():
    with torch.no_grad():
        ref(x)


print(f"eager torch bf16 (cuBLAS): fwd {timed(ref_fwd):.3f} ms, fwd+bwd {timed(ref_fb):.3f} ms", flush=True)
