"""Diagnostic: SM clock and board power while a workload loops for a few seconds (nvidia-smi sampling, 100 ms)."""
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import clip_embeds_b200.functional as Fk  # noqa: E402


def sample(fn, seconds, label, work_flops=None):
    proc = subprocess.Popen(["nvidia-smi", "--id=0", "--query-gpu=clocks.sm,power.draw,clocks_event_reasons.sw_power_cap",
                             "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE, text=True)
    lines = []
    threading.Thread(target=lambda: [lines.append(l) for l in proc.stdout], daemon=True).start()
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    n = 0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    while time.perf_counter() - t0 < seconds:
        for _ in range(5):
            fn()
        n += 5
        torch.cuda.synchronize()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    time.sleep(0.2)
    proc.terminate()
    rows = [[x.strip() for x in l.split(",")] for l in lines if l.count(",") >= 2]
    rows = rows[len(rows) // 3:]           # steady state: drop the first third
    clk = [float(r[0]) for r in rows]
    pw = [float(r[1]) for r in rows]
    cap = sum(1 for r in rows if r[2].lower().startswith("active"))
    extra = f" {work_flops / ms / 1e9:.0f} TFLOP/s" if work_flops else ""
    print(f"{label}: {ms:.3f} ms/iter{extra} | sm clock median {statistics.median(clk):.0f} MHz (min {min(clk):.0f}) "
          f"power median {statistics.median(pw):.0f} W, power-cap active in {cap}/{len(rows)} samples", flush=True)


if __name__ == "__main__":
    B, P, D = 1024, 576, 768
    V = torch.randn(B, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
    T = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
    g = torch.randn(B, B, device="cuda") / B

    def fb():
        V.grad = None
        T.grad = None
        Fk.pacl_scores(V, T, 10.0, "sigmoid").backward(g)

    def fwd():
        with torch.no_grad():
            Fk.pacl_scores(V, T, 10.0, "sigmoid")
    secs = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
    sample(fb, secs, "all-pairs fwd+bwd (7 units)", 14.0 * B * B * P * D)
    sample(fwd, secs, "all-pairs fwd (2 units)", 4.0 * B * B * P * D)
    A = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    Bm = torch.randn(8192, 8192, device="cuda").to(torch.bfloat16)
    sample(lambda: torch.matmul(A, Bm), secs, "cuBLAS 8192^3", 2.0 * 8192 ** 3)
    A2 = torch.randn(256, 1024, 768, device="cuda").to(torch.bfloat16)
    B2 = torch.randn(256, 576, 768, device="cuda").to(torch.bfloat16)
    sample(lambda: torch.matmul(A2, B2.transpose(1, 2)), secs, "cuBLAS batched 256 x [1024x576x768]", 2.0 * 256 * 1024 * 576 * 768)
