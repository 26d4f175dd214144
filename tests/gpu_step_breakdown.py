"""Where the non-engine time of the bench step goes (diagnostic): full step vs scoring fwd+bwd vs InfoNCE on scores."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import clip_embeds_b200.functional as Fk  # noqa: E402
from clip_embeds_b200.losses import PaclAllPairsLoss  # noqa: E402

B, P, D = 1024, 576, 768
V = torch.randn(B, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
T = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
loss_fn = PaclAllPairsLoss(0.1, "sigmoid")
g = torch.randn(B, B, device="cuda") / B


def timed(fn, iters=10, warm=3):
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


def full():
    V.grad = None
    T.grad = None
    loss_fn(V, T).backward()


def scoring():
    V.grad = None
    T.grad = None
    Fk.pacl_scores(V, T, 10.0).backward(g)


S = torch.randn(B, B, device="cuda").requires_grad_()


def infonce():
    S.grad = None
    Fk.score_infonce(S).backward()


for name, fn in (("full step", full), ("scoring fwd+bwd", scoring), ("InfoNCE on scores fwd+bwd", infonce), ("full step", full),
                 ("scoring fwd+bwd", scoring)):
    print(f"{name}: {timed(fn):.3f} ms", flush=True)
