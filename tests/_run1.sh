set -x
mkdir -p gpurun_out/s4
python -m pytest tests/test_gpu_ce.py -m gpu -x -q 2>&1 | tail -3
CLIPK_CE_ENGINE=1 python tests/gpu_ce_probe.py 2>&1 | tail -3
CLIPK_CE_ENGINE=2 python tests/gpu_ce_probe.py 2>&1 | tail -3
python tests/gpu_sparc_probe.py 10 2>&1 | tail -2
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/s4/sparc_launches.csv python tests/gpu_sparc_probe.py 1 > gpurun_out/s4/ncu_sparc.log 2>&1
python tests/gpu_perf_probe.py ap 1024 128:2 256:2 256:1 512:1 2>&1 | tail -5
