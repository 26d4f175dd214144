mkdir -p gpurun_out/s4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/heads_launches2.csv python tests/gpu_heads_probe.py once > gpurun_out/s4/ncu_heads2.log 2>&1
