mkdir -p gpurun_out/s4
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/sparc_launches2.csv python tests/gpu_sparc_probe.py 1 > gpurun_out/s4/ncu_sparc2.log 2>&1
