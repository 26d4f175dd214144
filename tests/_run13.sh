python tests/gpu_perf_probe.py ap 1024 128:2 64:4 64:3 96:3 128:3 128:4 128:2 2>&1 | tail -8
