mkdir -p gpurun_out/s4
python -m pytest tests/test_gpu_heads.py -m gpu -x -q 2>&1 | tail -3
python tests/gpu_heads_probe.py 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/heads_launches3.csv python tests/gpu_heads_probe.py once > gpurun_out/s4/ncu_heads3.log 2>&1
