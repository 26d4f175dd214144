python -m pytest tests/test_gpu_ce.py -m gpu -x -q 2>&1 | tail -3
CLIPK_CE_DLTMA=0 python tests/gpu_ce_probe.py 2>&1 | tail -2
CLIPK_CE_DLTMA=1 python tests/gpu_ce_probe.py 2>&1 | tail -2
