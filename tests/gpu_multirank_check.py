"""Multi-GPU parity check (run with torchrun on N GPUs; prints PASS/FAIL lines on rank 0).

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
        tests/gpu_multirank_check.py

Compares the sharded CUDA losses with the single-process fp32 oracle on the same (bf16-rounded) inputs:
  * OpenClipLoss (local-loss and global, gather_with_grad, usehardtext with ragged hard-negative counts)
  * PaclAllPairsLoss (image-sharded, text all-gather, column-LSE merge, reduce-scatter of dT)
  * SparcLoss (sample-sharded, global mask count)
"""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_oracle as O  # noqa: E402


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-30))


def main():
    rank = int(os.environ["RANK"])
    world = int(os.environ["WORLD_SIZE"])
    lr = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    from clip_embeds_b200 import losses
    from clip_embeds_b200.models import SparcHead
    ok = True

    def report(name, cond, msg):
        nonlocal ok
        flags = torch.tensor([1.0 if cond else 0.0], device=dev)
        dist.all_reduce(flags, op=dist.ReduceOp.MIN)
        good = bool(flags.item() > 0.5)
        ok = ok and good
        if rank == 0:
            print(("PASS " if good else "FAIL ") + name + " " + msg, flush=True)

    # ---------------------------------------------------------------- open_clip ClipLoss with hard negatives
    b, D = 160, 256      # > 128 rows: the local-loss case runs the one-GEMM symmetric CE with device-masked slabs
    hs = [(17 * (r + 1)) % (b + 1) for r in range(world)]
    imgs = [O.l2n(O.rn(100 + r, b, D)).to(torch.bfloat16) for r in range(world)]
    txts = [O.l2n(O.rn(200 + r, b + hs[r], D)).to(torch.bfloat16) for r in range(world)]
    for local in (True, False):
        io = [x.float().requires_grad_() for x in imgs]
        to = [x.float().requires_grad_() for x in txts]
        lo = O.openclip_loss_ranks(io, to, 20.0, local_loss=local, usehardtext=True)
        sum(lo).backward()
        img = imgs[rank].to(dev).requires_grad_()
        txt = txts[rank].to(dev).requires_grad_()
        fn = losses.OpenClipLoss(local_loss=local, gather_with_grad=True, rank=rank, world_size=world, usehardtext=True)
        loss = fn(img, txt, torch.tensor(20.0, device=dev))
        loss.backward()
        e = abs(loss.item() - lo[rank].item())
        ri, rt = rel(img.grad.float().cpu(), io[rank].grad), rel(txt.grad.float().cpu(), to[rank].grad)
        report(f"openclip local_loss={local}", e < 3e-3 * max(1, abs(lo[rank].item())) and ri < 2e-2 and rt < 2e-2,
               f"loss {loss.item():.5f} vs {lo[rank].item():.5f} rel_dimg {ri:.2e} rel_dtxt {rt:.2e}")

    # ---------------------------------------------------------------- PACL all-pairs, image-sharded
    bl, P, D = 8, 196, 512
    B = bl * world
    Vb = O.rn(300, B, P, D).to(torch.bfloat16)
    Tb = O.rn(301, B, D).to(torch.bfloat16)
    Vo, To = Vb.float().requires_grad_(), Tb.float().requires_grad_()
    lo = O.pacl_allpairs_loss(Vo, To, 0.1)
    lo.backward()
    V = Vb[rank * bl:(rank + 1) * bl].to(dev).requires_grad_()
    T = Tb[rank * bl:(rank + 1) * bl].to(dev).requires_grad_()
    loss = losses.PaclAllPairsLoss(0.1, group=dist.group.WORLD)(V, T)
    loss.backward()
    rv = rel(V.grad.float().cpu(), Vo.grad[rank * bl:(rank + 1) * bl])
    rt = rel(T.grad.float().cpu(), To.grad[rank * bl:(rank + 1) * bl])
    report("pacl allpairs sharded", abs(loss.item() - lo.item()) < 5e-3 and rv < 3e-2 and rt < 3e-2,
           f"loss {loss.item():.5f} vs {lo.item():.5f} relV {rv:.2e} relT {rt:.2e}")

    # ---------------------------------------------------------------- SPARC, sample-sharded
    bl, T_, P, D = 3, 77, 196, 512
    B = bl * world
    Vb = O.rn(400, B, P, D).to(torch.bfloat16)
    Lb = O.rn(401, B, T_, D).to(torch.bfloat16)
    eot = torch.tensor([(7 * i + 5) % T_ for i in range(B)])
    mask = (torch.arange(T_).expand(B, -1) <= eot.unsqueeze(1)).float()
    Vo, Lo = Vb.float().requires_grad_(), Lb.float().requires_grad_()
    v, lh, gh, _ = O.sparc_forward(Vo, Lo, mask, 1.0 / P)
    lo = O.sparc_loss(v, lh, gh, mask, 0.1)
    lo.backward()
    sl = slice(rank * bl, (rank + 1) * bl)
    V = Vb[sl].to(dev).requires_grad_()
    L = Lb[sl].to(dev).requires_grad_()
    v2, lh2, gh2, m2 = SparcHead(1.0 / P)(V, L, mask[sl].to(dev))
    loss = losses.SparcLoss(0.1, group=dist.group.WORLD)(v2, lh2, gh2, m2)
    loss.backward()
    rv, rl = rel(V.grad.float().cpu(), Vo.grad[sl]), rel(L.grad.float().cpu(), Lo.grad[sl])
    report("sparc sharded", abs(loss.item() - lo.item()) < 5e-3 and rv < 4e-2 and rl < 4e-2,
           f"loss {loss.item():.5f} vs {lo.item():.5f} relV {rv:.2e} relL {rl:.2e}")

    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
