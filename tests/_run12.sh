mkdir -p gpurun_out/s4
python tests/gpu_bench_rows.py --json gpurun_out/s4/rows.json 2>&1 | grep -o '"row": "[^"]*", "ours_ms": [0-9.]*, "eager_torch_ms": [0-9.]*'
python bench.py --steps 30 --warmup 5 > gpurun_out/s4/bench_30.json 2> gpurun_out/s4/bench_30.err; tail -c 1500 gpurun_out/s4/bench_30.json
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
