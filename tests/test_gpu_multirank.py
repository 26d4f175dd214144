"""GPU, >= 2 ranks over NCCL: the sharded losses against the single-process fp32 oracle (OpenClipLoss local / global with
ragged hard negatives through the fixed-capacity device-masked gather, image-sharded PaclAllPairsLoss, sample-sharded
SparcLoss).  The checks live in tests/gpu_multirank_check.py (one process per GPU, launched here with torchrun); the test
is skipped on a box with a single GPU."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_losses_vs_oracle(world):
    if torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs, found {torch.cuda.device_count()}")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(ROOT, "tests", "gpu_multirank_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    print(r.stdout[-4000:])
    assert r.returncode == 0, r.stdout[-4000:] + "\n" + r.stderr[-4000:]
    assert "FAIL" not in r.stdout and r.stdout.count("PASS") >= 4
