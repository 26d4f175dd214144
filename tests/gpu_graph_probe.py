"""CUDA-graph replay of the SPARC step on one GPU (diagnostic): eager vs graphed step time at small and full batch."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
from clip_embeds_b200 import losses  # noqa: E402
from clip_embeds_b200.graphs import GraphedStep  # noqa: E402
from clip_embeds_b200.models import SparcHead  # noqa: E402

T_, P, D = 77, 576, 768
dev = "cuda"


def timed(fn, warm=3, iters=20):
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


for B in (64, 512):
    V = O.rn(3, B, P, D).to(torch.bfloat16).to(dev).requires_grad_()
    L = O.rn(4, B, T_, D).to(torch.bfloat16).to(dev).requires_grad_()
    eot = torch.randint(5, T_, (B,), generator=torch.Generator().manual_seed(9))
    mask = (torch.arange(T_)[None, :] <= eot[:, None]).float().to(dev)
    head, sl = SparcHead(1.0 / P), losses.SparcLoss(0.1)

    def step():
        if V.grad is not None:
            V.grad.zero_()
            L.grad.zero_()
        v2, lh, gh, m2 = head(V, L, mask)
        loss = sl(v2, lh, gh, m2)
        loss.backward()
        return loss

    t_eager = timed(step)
    ref = float(step().item())
    gV = V.grad.clone()
    gs = GraphedStep(step)
    t_graph = timed(gs.replay)
    gs.replay()
    torch.cuda.synchronize()
    err = float((V.grad.float() - gV.float()).abs().max() / gV.float().abs().max())
    print(f"SPARC B={B}: eager {t_eager:.3f} ms/step, CUDA graph {t_graph:.3f} ms/step; loss {ref:.6f} vs {float(gs.loss.item()):.6f}, "
          f"max rel dV diff {err:.2e}", flush=True)
