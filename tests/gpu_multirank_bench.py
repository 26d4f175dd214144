"""Multi-GPU measurements of BASELINE.json configs[2] (SPARC, batch 512, sample-sharded) and configs[3] (NegCLIP-style
ClipLoss with left/right hard negatives, global batch 32768, text all-gather over NVLink), forward + backward, through the
drop-in loss modules.  Diagnostic companion of bench.py (which measures the headline metric); run with torchrun:

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 \
        tests/gpu_multirank_bench.py [--json out.json]

Timing: CUDA events around `iters` back-to-back steps after warm-up, barrier + synchronize on both sides, max over ranks.
Synthetic inputs: unit-norm randn features, hard-negative indicator ~ Bernoulli(0.25) per sample (ragged H_r
per rank), logit_scale 100; SPARC: eot positions randint(5, 77), sigma = 1/576.
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from clip_embeds_b200 import losses
    from clip_embeds_b200.models import SparcHead
    out = []

    def timed(fn, warm=3, iters=10):
        for _ in range(warm):
            fn()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / iters], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---------------------------------------------------------------- C4: NegCLIP ClipLoss, N = 32768, D = 768
    N, D = 32768, 768
    b = N // world
    g = torch.Generator().manual_seed(5 + 1000 * rank)
    hard = int((torch.rand(b, generator=g) < 0.25).sum())
    img = torch.nn.functional.normalize(torch.randn(b, D, generator=g), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    txt = torch.nn.functional.normalize(torch.randn(b + hard, D, generator=g), dim=-1).to(torch.bfloat16).to(dev).requires_grad_()
    scale = torch.tensor(100.0, device=dev)
    fn = losses.OpenClipLoss(local_loss=True, gather_with_grad=True, rank=rank, world_size=world, usehardtext=True)
    hs = torch.tensor([hard], device=dev)
    allh = [torch.zeros_like(hs) for _ in range(world)]
    dist.all_gather(allh, hs)
    H = int(sum(int(h.item()) for h in allh))

    def negclip():
        img.grad = None
        txt.grad = None
        fn(img, txt, 100.0).backward()

    ms = timed(negclip)
    flop = 6.0 * N * (N + H) * D            # the minimal single-logits-matrix figure, all GPUs
    out.append({"config": "C4 NegCLIP ClipLoss local-loss + gather_with_grad + usehardtext, N=32768, D=768, bf16",
                "n_gpus": world, "hard_negatives_total": H, "ms_per_step": ms, "pairs_per_s": N / ms * 1e3,
                "algorithmic_TFLOPps_all_gpus": flop / ms / 1e9,
                "executed_TFLOPps_per_gpu": 8.0 * (b * (N + H) + (b + hard) * N) * D / ms / 1e9})
    del img, txt

    # ---------------------------------------------------------------- C3: SPARC, B = 512, T = 77, P = 576, D = 768
    B, T_, P = 512, 77, 576
    bl = B // world
    g = torch.Generator().manual_seed(3 + 1000 * rank)
    V = torch.randn(bl, P, D, generator=g).to(torch.bfloat16).to(dev).requires_grad_()
    L = torch.randn(bl, T_, D, generator=g).to(torch.bfloat16).to(dev).requires_grad_()
    eot = torch.randint(5, T_, (bl,), generator=g)
    mask = (torch.arange(T_)[None, :] <= eot[:, None]).float().to(dev)
    head = SparcHead(1.0 / P)
    sl = losses.SparcLoss(0.1, group=dist.group.WORLD)

    def sparc():
        V.grad = None
        L.grad = None
        v2, lh, gh, m2 = head(V, L, mask)
        sl(v2, lh, gh, m2).backward()

    ms = timed(sparc)
    out.append({"config": "C3 SPARC alignment + SparcLoss, B=512 (sample-sharded), T=77, P=576, D=768, bf16", "n_gpus": world,
                "ms_per_step": ms, "pairs_per_s": B / ms * 1e3,
                "algorithmic_GBps_per_gpu": (3 * bl * P * D * 2 + 5 * bl * T_ * D * 2) / ms / 1e6})
    if rank == 0:
        for r in out:
            print(json.dumps(r), flush=True)
        if "--json" in sys.argv:
            with open(sys.argv[sys.argv.index("--json") + 1], "w") as f:
                json.dump(out, f, indent=1)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
