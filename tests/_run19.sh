python -m pytest tests/test_gpu_ce.py tests/test_gpu_heads.py -m gpu -x -q 2>&1 | tail -3
python tests/gpu_ce_probe.py 2>&1 | tail -2
python tests/gpu_heads_probe.py 2>&1 | grep "heads B"
