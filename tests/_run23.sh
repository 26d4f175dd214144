python -m pytest tests/test_gpu_sparc.py -m gpu -x -q 2>&1 | tail -3
python tests/gpu_sparc_probe.py 20 2>&1 | tail -1
