"""Diagnostic: N SPARC steps (align + SparcLoss, fwd + bwd) at BASELINE configs[2] shape, for ncu launch lists."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200 import losses  # noqa: E402
from clip_embeds_b200.models import SparcHead  # noqa: E402

B, T_, P, D = int(sys.argv[1]) if len(sys.argv) > 1 else 512, 77, 576, 768
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1
g = torch.Generator().manual_seed(3)
V = torch.randn(B, P, D, generator=g).to(torch.bfloat16).cuda().requires_grad_()
L = torch.randn(B, T_, D, generator=g).to(torch.bfloat16).cuda().requires_grad_()
eot = torch.randint(5, T_, (B,), generator=g)
mask = (torch.arange(T_)[None, :] <= eot[:, None]).float().cuda()
head = SparcHead(1.0 / P)
sl = losses.SparcLoss(0.1)
for _ in range(n):
    V.grad = None
    L.grad = None
    v2, lh, gh, m2 = head(V, L, mask)
    sl(v2, lh, gh, m2).backward()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    V.grad = None
    L.grad = None
    v2, lh, gh, m2 = head(V, L, mask)
    sl(v2, lh, gh, m2).backward()
e1.record()
torch.cuda.synchronize()
print(f"sparc step B={B}: {e0.elapsed_time(e1) / n:.3f} ms")
