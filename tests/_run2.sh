python tests/gpu_perf_probe.py ap 1024 256:2 256:3 256:4 512:2 342:3 171:3 205:5 2>&1 | tail -8
python tests/gpu_perf_probe.py apx 128 1024 128:1 64:2 43:3 32:4 2>&1 | tail -5
python tests/gpu_perf_probe.py apx 256 1024 128:2 86:3 64:4 2>&1 | tail -5
python tests/gpu_perf_probe.py apx 512 1024 128:2 256:2 171:3 128:4 2>&1 | tail -5
