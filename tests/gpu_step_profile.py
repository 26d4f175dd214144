"""Where does a bench step spend its time?  torch.profiler over a few PaclAllPairsLoss steps (diagnostic)."""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200.losses import PaclAllPairsLoss  # noqa: E402

B, P, D = 1024, 576, 768
V = torch.randn(B, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
T = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
loss_fn = PaclAllPairsLoss(0.1)


def step():
    V.grad = None
    T.grad = None
    loss = loss_fn(V, T)
    loss.backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(4):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
busy = sum(e.time_range.elapsed_us() for e in ev)
span = ev[-1].time_range.end - ev[0].time_range.start
print(f"GPU span {span/4/1e3:.3f} ms/step, sum of kernel time {busy/4/1e3:.3f} ms/step, idle {100*(1-busy/span):.1f}%")
gaps = []
for a, b in zip(ev[:-1], ev[1:]):
    g = b.time_range.start - a.time_range.end
    if g > 15:
        gaps.append((g, a.name[:60], b.name[:60]))
gaps.sort(reverse=True)
for g, a, b in gaps[:14]:
    print(f"  gap {g:8.1f} us  after {a}  before {b}")
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=60))
