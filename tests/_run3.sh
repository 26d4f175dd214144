python -m pytest tests -m gpu -x -q 2>&1 | tail -8
python tests/gpu_sparc_probe.py 10 2>&1 | tail -2
python tests/gpu_bench_rows.py 2>&1 | grep -o '"row": "[^"]*", "ours_ms": [0-9.]*' 
