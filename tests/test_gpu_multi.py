"""Multi-GPU scenarios as `-m gpu` tests (reference: `mpirun -np 4 multiverso.test ...`, Test/main.cpp:12-25):
torchrun one rank per visible GPU over tests/mp_device_check.py (BSP + async table scenarios, row ops,
KV, aggregate, fused Get+GEMM over peer shards, the device-side WordEmbedding block protocol).  Skipped
with fewer than two GPUs."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("which", ["sync", "async"])
def test_multi_gpu_scenarios(which):
    n = _ngpu()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    n = min(n, 8)
    port = 29700 + (os.getpid() % 200) + (0 if which == "sync" else 1)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
           "--master-addr", "127.0.0.1", "--master-port", str(port),
           os.path.join(ROOT, "tests", "mp_device_check.py"), which]
    r = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    tail = (r.stdout + r.stderr)[-3000:]
    assert r.returncode == 0, tail
    assert "FAIL" not in r.stdout, tail
