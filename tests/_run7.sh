mkdir -p gpurun_out/s4
python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/s4/bench_a.json 2> gpurun_out/s4/bench_a.err; tail -c 1800 gpurun_out/s4/bench_a.json
