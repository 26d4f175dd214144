#!/usr/bin/env python
"""Benchmark of the hot path: PACL all-pairs patch-aligned scoring + InfoNCE, forward + backward.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference ...                     (CPU arm: the oracle port of the reference on host cores)

Workload (BASELINE.json configs[1]): ViT-L/14-336 shape — P = 576 patches, D = 768, global batch 1024 image-text
pairs, bf16, synthetic features (random-init, seeds 1/2).  A "step" is one forward + backward of
`PaclAllPairsLoss(temperature=0.1)` over the whole global batch: all-pairs text-conditioned scores [B, B], symmetric
InfoNCE, gradients w.r.t. every patch token and text embedding.  With N > 1 the images are sharded (B/N per rank,
STRONG scaling: the global batch is fixed), texts are all-gathered over NCCL, dT is reduce-scattered.

Prints ONE JSON line (see the key list at the bottom).  `value` = device-resident throughput; `e2e` = the same call
fed from pinned HOST buffers every step (H2D of V and T inside the timed region, D2H of the loss).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

B_GLOBAL, P, D = 1024, 576, 768
TEMPERATURE = 0.1
METRIC = "image-text pairs/sec, PACL fwd+bwd, ViT-L/14-336 shape"


def _peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            pk = json.load(f)
        return float(pk["bf16_tflops_sustained"]), "measured (MEASURED_PEAKS.json bf16_tflops_sustained: kernel timed inside a long step)"
    except Exception:
        return 1400.0, "fallback (B200_PROFILING.md sustained ~1.4 PFLOP/s)"


class ClockSampler:
    """Samples SM clocks and throttle reasons with nvidia-smi while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [x.strip() for x in ln.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[2:6]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_reference_sample(n_images, threads, iters=2):
    """The reference's own algorithm for this path (oracle port, torch CPU fp32): per-image eval-style loop
    (one image x all B texts), InfoNCE-style upstream gradient, forward + backward.  Returns seconds per sample."""
    import torch
    from oracle import ref_oracle as O
    torch.set_num_threads(threads)
    V = O.rn(1, n_images, P, D).requires_grad_()
    T = O.rn(2, B_GLOBAL, D).requires_grad_()
    g = O.rn(3, n_images, B_GLOBAL) / B_GLOBAL
    best = None
    for _ in range(iters):
        V.grad = None
        T.grad = None
        t0 = time.perf_counter()
        s = O.pacl_allpairs_scores(V, T, 1.0 / TEMPERATURE)
        (s * g).sum().backward()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    return best


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    cores = os.cpu_count() or 1
    n_img = 4
    times = []
    for _ in range(max(1, args.warmup)):
        cpu_reference_sample(n_img, cores, iters=1)
    t_all0 = time.perf_counter()
    for _ in range(args.steps):
        times.append(cpu_reference_sample(n_img, cores, iters=1))
        if time.perf_counter() - t_all0 > 150:
            break
    t = statistics.median(times)
    val = n_img / t
    sample = (f"{n_img} images x all {B_GLOBAL} texts (P={P}, D={D}), fp32 torch CPU, forward+backward of the per-image "
              f"reference loop; pairs/s = images per second against the full text batch")
    out = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": len(times), "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"PACL all-pairs fwd+bwd, B={B_GLOBAL} texts, P={P}, D={D} (bounded sample of {n_img} images per step)"},
        "cpu_baseline": {"value": val, "unit": "pairs/s", "cores": torch.get_num_threads(), "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(out), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def run_gpu(args):
    import torch
    import torch.distributed as dist
    from clip_embeds_b200 import _lib
    from clip_embeds_b200.losses import PaclAllPairsLoss

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.gpus != world and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    group = None
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
        group = dist.group.WORLD
    _lib.check(_lib.lib().clipk_check_device())
    b = B_GLOBAL // world
    assert b * world == B_GLOBAL

    # synthetic inputs (SURVEY §8d seeds: V=1, T=2; per-rank seed = base + 1000*rank), bf16
    gv = torch.Generator().manual_seed(1 + 1000 * rank)
    gt = torch.Generator().manual_seed(2 + 1000 * rank)
    V_host = torch.randn(b, P, D, generator=gv).to(torch.bfloat16).pin_memory()
    T_host = torch.randn(b, D, generator=gt).to(torch.bfloat16).pin_memory()
    V = V_host.to(dev).requires_grad_()
    T = T_host.to(dev).requires_grad_()
    loss_fn = PaclAllPairsLoss(TEMPERATURE, "sigmoid", group=group)

    def step(Vt, Tt):
        Vt.grad = None
        Tt.grad = None
        loss = loss_fn(Vt, Tt)
        loss.backward()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        step(V, T)
    barrier()

    # ---- timed region 1: device-resident inputs (value) ------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    lc0 = _lib.lib().clipk_launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for _ in range(args.steps):
        step(V, T)
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    launches = (_lib.lib().clipk_launch_count() - lc0) // args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---- engine-only time (roofline): events around the scoring forward and backward of each step -------------
    import clip_embeds_b200.functional as Fk
    eng_samples = []
    for _ in range(min(args.steps, 5) + 1):
        Vt = V.detach().requires_grad_()
        Tt = T.detach()
        if world > 1:
            Tall = torch.empty(B_GLOBAL, D, dtype=T.dtype, device=dev)
            dist.all_gather_into_tensor(Tall, Tt)
        else:
            Tall = Tt
        Tall = Tall.requires_grad_()
        g = torch.randn(b, B_GLOBAL, device=dev) / B_GLOBAL
        e = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
        torch.cuda.synchronize()
        e[0].record()
        s = Fk.pacl_scores(Vt, Tall, 1.0 / TEMPERATURE)
        e[1].record()
        e[2].record()
        s.backward(g)
        e[3].record()
        torch.cuda.synchronize()
        eng_samples.append(e[0].elapsed_time(e[1]) + e[2].elapsed_time(e[3]))
    eng_ms = statistics.median(eng_samples[1:])        # first pass is an untimed warm-up of this call pattern

    # ---- timed region 2: end to end from pinned host buffers (e2e) ---------------------------------------------
    copy_stream = torch.cuda.Stream()
    bufs = [(torch.empty_like(V_host, device=dev), torch.empty_like(T_host, device=dev)) for _ in range(2)]
    ready = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]

    def upload(i):
        with torch.cuda.stream(copy_stream):
            copy_stream.wait_event(consumed[i % 2])
            bufs[i % 2][0].copy_(V_host, non_blocking=True)
            bufs[i % 2][1].copy_(T_host, non_blocking=True)
            ready[i % 2].record(copy_stream)

    def e2e_loop(n):
        main = torch.cuda.current_stream()
        for c in consumed:
            c.record(main)
        upload(0)
        last = None
        for i in range(n):
            if i + 1 < n:
                upload(i + 1)                      # overlaps the H2D of step i+1 with the compute of step i
            main.wait_event(ready[i % 2])
            Vt = bufs[i % 2][0].detach().requires_grad_()
            Tt = bufs[i % 2][1].detach().requires_grad_()
            loss = step(Vt, Tt)
            consumed[i % 2].record(main)
            last = loss.item()                     # D2H read of the step's result
        return last

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    ev2, ev3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev2.record()
    last_loss = e2e_loop(args.steps)
    ev3.record()
    barrier()
    e2e_ms = ev2.elapsed_time(ev3)
    _ = time.perf_counter() - t0

    times = torch.tensor([dev_ms, e2e_ms, eng_ms], device=dev)
    if world > 1:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, eng_ms = [float(x) for x in times.tolist()]

    if rank == 0:
        ms_per_step = dev_ms / args.steps
        value = B_GLOBAL / (ms_per_step * 1e-3)
        e2e_value = B_GLOBAL / (e2e_ms / args.steps * 1e-3)
        peak, peak_src = _peaks()
        flops_per_rank = 12.0 * b * B_GLOBAL * P * D          # algorithmic: 4 fwd + 8 bwd GEMM-flops per (i,k,p,d)
        achieved = flops_per_rank / (eng_ms * 1e-3) / 1e12
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            n_img = 4
            t = cpu_reference_sample(n_img, cores, iters=2)
            cpu = {"value": n_img / t, "unit": "pairs/s", "cores": cores, "kind": "port",
                   "sample": f"{n_img} images x all {B_GLOBAL} texts, fp32 oracle port (per-image reference loop) fwd+bwd, best of 2"}
        out = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"PACL all-pairs scoring + InfoNCE fwd+bwd, global batch {B_GLOBAL}, P={P}, D={D}, "
                                   f"sigmoid(10 cos) activation, T=0.1 (BASELINE configs[1])",
                       "global_batch": B_GLOBAL, "per_gpu_images": b, "parallelism": f"image-sharded dp{world}",
                       "l2": "inputs larger than L2 (V = %.0f MB per rank)" % (b * P * D * 2 / 1e6)},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": world * (V_host.numel() + T_host.numel()) * 2,
                    "d2h_bytes_per_step": world * 4, "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(launches),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one step's engine launches, from the ncu --set full
                         # capture of profiles/r01_ncu_staged_kernels.summary.txt (2.94 GB per 128-image group), scaled to
                         # this rank's share of the images
                         "traffic": 2.94e9 * (b / 128.0), "kernel": "eng2::gemm2_kernel (tcgen05 cta_group::2 engine, all fused-epilogue instantiations of one step)",
                         "algorithmic_flops_per_step_per_gpu": flops_per_rank, "engine_ms_per_step": eng_ms,
                         "peak_source": peak_src},
            "clocks": clocks,
            "loss": last_loss,
        }
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="clipk", choices=["clipk", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_gpu(args)


if __name__ == "__main__":
    main()
