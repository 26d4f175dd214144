"""SPARC row probe (diagnostic): alignment + SparcLoss fwd+bwd at the C3 shape on one GPU; run under
`ncu --metrics gpu__time_duration.sum` for the launch list, or plain for the CUDA-event time."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
import clip_embeds_b200.functional as Fk  # noqa: E402
from clip_embeds_b200 import losses  # noqa: E402

Bs, Tt, P, D = 512, 77, 576, 768
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = "cuda"
Vs = O.rn(3, Bs, P, D).to(torch.bfloat16).to(dev).requires_grad_()
Ls = O.rn(4, Bs, Tt, D).to(torch.bfloat16).to(dev).requires_grad_()
eot = torch.randint(5, Tt, (Bs,), generator=torch.Generator().manual_seed(9))
mask = (torch.arange(Tt)[None, :] <= eot[:, None]).float().to(dev)
sl = losses.SparcLoss(0.1)


def step():
    Vs.grad = None
    Ls.grad = None
    l_hat, g_hat = Fk.sparc_align(Vs, Ls, 1.0 / P)
    sl(Vs, l_hat, g_hat, mask).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(iters):
    step()
e1.record()
torch.cuda.synchronize()
print(f"sparc align+loss fwd+bwd: {e0.elapsed_time(e1) / iters:.3f} ms/step", flush=True)
