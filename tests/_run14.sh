python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for pdl in 0 1 0 1; do
echo "== CLIPK_PDL=$pdl"
CLIPK_PDL=$pdl python tests/gpu_perf_probe.py ap 1024 128:2 128:1 64:2 2>&1 | tail -3
done
for pdl in 0 1; do
echo "== CLIPK_PDL=$pdl"
CLIPK_PDL=$pdl python tests/gpu_sparc_probe.py 10 2>&1 | tail -1
CLIPK_PDL=$pdl python tests/gpu_ce_probe.py 2>&1 | tail -2
CLIPK_PDL=$pdl python tests/gpu_heads_probe.py 2>&1 | grep "heads B"
CLIPK_PDL=$pdl python tests/gpu_perf_probe.py apx 128 1024 128:1 64:2 2>&1 | tail -2
done
