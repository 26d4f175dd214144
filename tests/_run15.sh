python -m pytest tests/test_gpu_evalproto.py -m gpu -x -q 2>&1 | tail -12
