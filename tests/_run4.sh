mkdir -p gpurun_out/s4
python tests/gpu_sparc_probe.py 20 2>&1 | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/sparc_launches3.csv python tests/gpu_sparc_probe.py 1 > gpurun_out/s4/ncu_sparc3.log 2>&1
