"""Randomised shape sweep of the all-pairs path (staged, persistent; sigmoid / ones / softmax) against the oracle.
Diagnostic (run on the GPU box): python tests/gpu_shape_sweep.py [n_cases] [seed]"""
import os
import random
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402
import clip_embeds_b200.functional as Fk  # noqa: E402


def rel(a, b):
    return ((a - b).norm() / b.norm().clamp_min(1e-12)).item()


def case(Bi, Bt, P, D, act, group, seed):
    Vb = O.rn(seed, Bi, P, D).to(torch.bfloat16)
    Tb = O.rn(seed + 1, Bt, D).to(torch.bfloat16)
    gs = O.rn(seed + 2, Bi, Bt) / (Bi * Bt) ** 0.5
    Vo, To = Vb.float().requires_grad_(), Tb.float().requires_grad_()
    so = O.pacl_allpairs_scores(Vo, To, 10.0, act)
    (so * gs).sum().backward()
    V, T = Vb.cuda().requires_grad_(), Tb.cuda().requires_grad_()
    s = Fk.pacl_scores(V, T, 10.0, act, group)
    (s * gs.cuda()).sum().backward()
    torch.cuda.synchronize()
    es = (s.detach().cpu() - so.detach()).abs().max().item()
    rv, rt = rel(V.grad.float().cpu(), Vo.grad), rel(T.grad.float().cpu(), To.grad)
    ok = es < 3e-2 and rv < 4e-2 and rt < 4e-2
    print(f"{'OK  ' if ok else 'FAIL'} Bi={Bi} Bt={Bt} P={P} D={D} act={act} group={group}: |ds|={es:.2e} relV={rv:.2e} relT={rt:.2e}",
          flush=True)
    return ok


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    rng = random.Random(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    bad = 0
    for i in range(n):
        Bi = rng.choice([1, 2, 3, 5, 9, 17])
        Bt = rng.choice([1, 2, 7, 64, 127, 128, 129, 200, 257, 300])
        P = rng.choice([1, 7, 49, 64, 65, 192, 193, 197, 256, 257, 400])
        D = rng.choice([8, 16, 64, 72, 128, 136, 256, 264, 520])
        act = rng.choice(["sigmoid", "sigmoid", "ones", "softmax"])
        group = rng.choice([None, 1, 2, (2, 2), (4, 3), 0, (-1, 1), (-2, 2), (-3, 3), (-5, 2)])
        try:
            bad += 0 if case(Bi, Bt, P, D, act, group, 1000 + 7 * i) else 1
        except Exception as e:      # noqa: BLE001
            bad += 1
            print(f"EXC  Bi={Bi} Bt={Bt} P={P} D={D} act={act} group={group}: {type(e).__name__}: {str(e)[:200]}", flush=True)
            if "CUDA error" in str(e) or "launch failure" in str(e):
                break
    print(f"{n} cases, {bad} failures")
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())
