"""Data parallelism on real GPUs over NCCL: needs >= 2 devices, skipped otherwise (the CPU `gloo` tests
in tests/test_dp_gloo.py cover the protocol on every box)."""
import os
import subprocess
import sys
from pathlib import Path

import pytest

pytestmark = pytest.mark.gpu


def test_two_rank_nccl_step_matches_single_process(cuda):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import socket

    worker = Path(__file__).resolve().parent / "dp_nccl_worker.py"
    env = dict(os.environ, MASTER_ADDR="127.0.0.1")
    with socket.socket() as sock:
        sock.bind(("127.0.0.1", 0))
        port = sock.getsockname()[1]
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
                        "127.0.0.1", "--master-port", str(port), str(worker)], capture_output=True, text=True, env=env, timeout=300)
    assert r.returncode == 0 and "DP_OK" in r.stdout, r.stdout[-3000:] + r.stderr[-3000:]
