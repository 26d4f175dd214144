"""Diagnostic: where the host time of one SPARC step goes (cProfile of 20 steps at B samples)."""
import cProfile
import os
import pstats
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200 import losses  # noqa: E402
from clip_embeds_b200.models import SparcHead  # noqa: E402

B, T_, P, D = int(sys.argv[1]) if len(sys.argv) > 1 else 64, 77, 576, 768
g = torch.Generator().manual_seed(3)
V = torch.randn(B, P, D, generator=g).to(torch.bfloat16).cuda().requires_grad_()
L = torch.randn(B, T_, D, generator=g).to(torch.bfloat16).cuda().requires_grad_()
eot = torch.randint(5, T_, (B,), generator=g)
mask = (torch.arange(T_)[None, :] <= eot[:, None]).float().cuda()
head = SparcHead(1.0 / P)
sl = losses.SparcLoss(0.1)


def step():
    V.grad = None
    L.grad = None
    v2, lh, gh, m2 = head(V, L, mask)
    sl(v2, lh, gh, m2).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(20):
    step()
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"B={B}: host enqueue {1e3 * (t1 - t0) / 20:.3f} ms/step, wall {1e3 * (t2 - t0) / 20:.3f} ms/step")
pr = cProfile.Profile()
pr.enable()
for _ in range(20):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(18)
