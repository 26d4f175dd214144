python -m pytest tests/test_gpu_heads.py -m gpu -x -q 2>&1 | tail -4
