mkdir -p gpurun_out/s4
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python tests/gpu_bench_rows.py --json gpurun_out/s4/rows.json 2>&1 | grep -o '"row": "[^"]*", "ours_ms": [0-9.]*, "eager_torch_ms": [0-9.]*'
