python -m pytest tests/test_gpu_pacl.py -m gpu -x -q 2>&1 | tail -3
python tests/gpu_bench_rows.py 2>&1 | grep -o '"row": "a[13][^"]*", "ours_ms": [0-9.]*'
