python tests/gpu_perf_probe.py ap 1024 128:2 18:1 37:1 37:2 74:1 74:2 128:1 2>&1 | tail -8
CLIPK_PDL=0 python tests/gpu_perf_probe.py ap 1024 18:1 37:1 2>&1 | tail -2
