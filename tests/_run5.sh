mkdir -p gpurun_out/s4
python tests/gpu_ce_probe.py once > gpurun_out/s4/ce_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gemm2_kernel -s 8 -c 8 -o gpurun_out/s4/prof_ce python tests/gpu_ce_probe.py once > gpurun_out/s4/ncu_ce.log 2>&1
tail -3 gpurun_out/s4/ncu_ce.log
