"""Multi-GPU parity (NCCL over NVLink): runs tests/dist_gpu_worker.py on min(device_count, 8) GPUs, one process each."""
import os
import socket
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("transport", ["p2p", "nccl"])
def test_sharded_multiply_over_nvlink(transport):
    """p2p: tiles pulled over NVLink peer memory + partials stored by the GEMM epilogue into the owner's HBM;
    nccl: the same plan with grouped NCCL send/recv (the fallback transport)."""
    import torch
    n = min(torch.cuda.device_count(), 8)
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run under `gpurun --gpus 2`)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), str(ROOT / "tests" / "dist_gpu_worker.py")]
    out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=900,
                         env=dict(os.environ, MARLIN_B200_TRANSPORT=transport))
    assert out.returncode == 0, out.stdout[-4000:]
    for r in range(n):
        assert f"rank {r} ok transport={transport}" in out.stdout
    if transport == "p2p":
        assert "mesh=yes" in out.stdout, "peer-memory transport was not active:\n" + out.stdout[-2000:]
