"""Quick perf probe (diagnostic, not the bench): engine GEMM TFLOP/s and all-pairs fwd/bwd times."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from clip_embeds_b200 import _lib  # noqa: E402
import clip_embeds_b200.functional as Fk  # noqa: E402


def timeit(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(iters + 1)]
    ev[0].record()
    for i in range(iters):
        fn()
        ev[i + 1].record()
    torch.cuda.synchronize()
    ts = [ev[i].elapsed_time(ev[i + 1]) for i in range(iters)]
    return min(ts), sum(ts) / len(ts)


def gemm(M, N, K, nb=1, a_mn=0, b_mn=0):
    A = torch.randn(nb, M, K, device="cuda").to(torch.bfloat16)
    B = torch.randn(nb, N, K, device="cuda").to(torch.bfloat16)
    Ad = A.transpose(1, 2).contiguous() if a_mn else A
    Bd = B.transpose(1, 2).contiguous() if b_mn else B
    C = torch.empty(nb, M, N, dtype=torch.bfloat16, device="cuda")
    st = torch.cuda.current_stream().cuda_stream

    def run():
        _lib.call("clipk_gemm_bf16", Ad.data_ptr(), a_mn, M if a_mn else K, M * K, Bd.data_ptr(), b_mn, N if b_mn else K,
                  N * K, C.data_ptr(), N, M * N, 0, M, N, K, nb, 1.0, 0, st)
    mn, av = timeit(run)
    fl = 2.0 * M * N * K * nb
    print(f"gemm M={M} N={N} K={K} nb={nb} a_mn={a_mn} b_mn={b_mn}: {mn:.3f} ms  {fl / mn / 1e9:.1f} TFLOP/s", flush=True)
    t0 = timeit(lambda: torch.matmul(A, B.transpose(1, 2)))[0]
    print(f"   cuBLAS (torch.matmul) same shape: {t0:.3f} ms {fl / t0 / 1e9:.1f} TFLOP/s", flush=True)


def allpairs(B, P, D, group=None):
    V = torch.randn(B, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
    T = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
    g = torch.randn(B, B, device="cuda") / B

    def fwd():
        with torch.no_grad():
            Fk.pacl_scores(V, T, 10.0, "sigmoid", group)

    def fb():
        V.grad = None
        T.grad = None
        s = Fk.pacl_scores(V, T, 10.0, "sigmoid", group)
        s.backward(g)
    tf = timeit(fwd, 3, 1)[0]
    tb = timeit(fb, 3, 1)[0]
    fl = 12.0 * B * B * P * D
    print(f"allpairs B={B} P={P} D={D} group={group}: fwd {tf:.2f} ms ({4.0*B*B*P*D/tf/1e9:.0f} TF/s) fwd+bwd {tb:.2f} ms "
          f"-> {B / tb * 1e3:.0f} pairs/s, {fl / tb / 1e9:.0f} TFLOP/s algorithmic", flush=True)


def allpairs_rect(Bi, Bt, P, D, group):
    """One rank's share of the sharded step: Bi local images against Bt gathered texts."""
    V = torch.randn(Bi, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
    T = torch.randn(Bt, D, device="cuda").to(torch.bfloat16).requires_grad_()
    g = torch.randn(Bi, Bt, device="cuda") / Bt

    def fb():
        V.grad = None
        T.grad = None
        Fk.pacl_scores(V, T, 10.0, "sigmoid", group).backward(g)
    tb = timeit(fb, 10, 3)[0]
    print(f"allpairs Bi={Bi} Bt={Bt} group={group}: fwd+bwd {tb:.3f} ms -> {12.0*Bi*Bt*P*D/tb/1e9:.0f} TFLOP/s algorithmic", flush=True)


def allpairs_once(B, P, D, group):
    V = torch.randn(B, P, D, device="cuda").to(torch.bfloat16).requires_grad_()
    T = torch.randn(B, D, device="cuda").to(torch.bfloat16).requires_grad_()
    g = torch.randn(B, B, device="cuda") / B
    s = Fk.pacl_scores(V, T, 10.0, "sigmoid", group)
    s.backward(g)
    torch.cuda.synchronize()


if __name__ == "__main__":
    cmd = sys.argv[1] if len(sys.argv) > 1 else "all"
    if cmd in ("gemm", "all"):
        gemm(4096, 4096, 4096)
        for nb in (16, 64, 256):
            gemm(1024, 576, 768, nb=nb)
        gemm(1024, 768, 576, nb=64, b_mn=1)
        gemm(576, 768, 1024, nb=64, a_mn=1, b_mn=1)
    if cmd in ("ap", "all"):
        B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
        groups = [tuple(int(y) for y in x.split(":")) if ":" in x else int(x) for x in sys.argv[3:]] or [16, 32, 64]
        for grp in groups:
            allpairs(B, 576, 768, grp)
    if cmd == "apx":
        Bi, Bt = int(sys.argv[2]), int(sys.argv[3])
        for x in sys.argv[4:]:
            allpairs_rect(Bi, Bt, 576, 768, tuple(int(y) for y in x.split(":")) if ":" in x else int(x))
    if cmd == "once":
        a = sys.argv[3]
        allpairs_once(int(sys.argv[2]), 576, 768, tuple(int(y) for y in a.split(":")) if ":" in a else int(a))
    if cmd == "tile":
        gemm(1024, 576, 768, nb=64)
        gemm(1024, 576, 6144, nb=8)
        gemm(1024, 512, 768, nb=64)
        gemm(1024, 1024, 768, nb=32)
        gemm(2048, 2048, 768, nb=4)
    if cmd == "epi":
        gemm(1024, 576, 64, nb=256)
        gemm(1024, 576, 128, nb=256)
        gemm(1024, 576, 384, nb=256)
        gemm(1024, 576, 768, nb=256)
        gemm(1024, 768, 64, nb=256)
        gemm(1024, 768, 576, nb=256)
    if cmd == "k5":
        gemm(1024, 768, 576, nb=128, b_mn=1)
        gemm(1024, 768, 12672, nb=6, b_mn=1)
        gemm(1024, 768, 12672, nb=6, b_mn=0)
        gemm(1024, 768, 6336, nb=12, b_mn=1)
        gemm(1024, 768, 3168, nb=24, b_mn=1)
    if cmd == "bn":
        gemm(4096, 256, 4096, nb=16)
        gemm(4096, 192, 4096, nb=16)
        gemm(4096, 128, 4096, nb=32)
        gemm(4096, 64, 4096, nb=64)
    if cmd == "bn2":        # MMA cost versus N: 592 pair-tiles (8 full waves of 74 clusters), K = 4096
        for n in (256, 224, 192, 160, 128, 96, 64):
            gemm(4096, n, 4096, nb=37)
    if cmd == "trace":      # run with CLIPK_TRACE=1
        B, grp = int(sys.argv[2]), int(sys.argv[3])
        allpairs_once(B, 576, 768, grp)
        _lib.lib().clipk_trace_dump()
        allpairs_once(B, 576, 768, grp)
        _lib.lib().clipk_trace_dump()
    if cmd == "fwdonce":
        V = torch.randn(int(sys.argv[2]), int(os.environ.get("PROBE_P", "576")), 768, device="cuda").to(torch.bfloat16)
        T = torch.randn(1024, 768, device="cuda").to(torch.bfloat16)
        with torch.no_grad():
            a = sys.argv[3]
            for _ in range(int(os.environ.get("PROBE_REPEAT", "1"))):
                Fk.pacl_scores(V, T, 10.0, "sigmoid", tuple(int(y) for y in a.split(":")) if ":" in a else int(a))
                torch.cuda.synchronize()
    if cmd == "cpu":
        import time
        B = 1024
        V = torch.randn(B, 576, 768, device="cuda").to(torch.bfloat16)
        T = torch.randn(B, 768, device="cuda").to(torch.bfloat16)
        for grp in (8, 16, 128):
            with torch.no_grad():
                Fk.pacl_scores(V, T, 10.0, "sigmoid", grp)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                Fk.pacl_scores(V, T, 10.0, "sigmoid", grp)
                t1 = time.perf_counter()
                torch.cuda.synchronize()
                t2 = time.perf_counter()
            n = 2 * ((B + grp - 1) // grp)
            print(f"group={grp}: host enqueue {1e3*(t1-t0):.2f} ms for {n} launches ({1e6*(t1-t0)/n:.1f} us/launch), total {1e3*(t2-t0):.2f} ms", flush=True)
