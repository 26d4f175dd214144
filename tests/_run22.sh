mkdir -p gpurun_out/s4
python bench.py --steps 3 --warmup 3 > gpurun_out/s4/bench_short.json 2> gpurun_out/s4/bench_short.err && ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/bench_launches.csv python bench.py --steps 3 --warmup 3 > gpurun_out/s4/ncu_bench.log 2>&1
tail -c 300 gpurun_out/s4/ncu_bench.log
