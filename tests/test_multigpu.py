"""Multi-rank correctness on hardware: two ranks, one per GPU, NCCL (needs >= 2 GPUs: `gpurun --gpus 2`; skipped on a
single-GPU box).  The sharded run must reproduce the single-GPU result on the union bundle: tests/multigpu_check.py."""
import os
import socket
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.gpu
@pytest.mark.timeout(600)
def test_two_ranks_reproduce_the_single_gpu_result_on_the_union_bundle():
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr",
           "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py")]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=540, cwd=ROOT)
    assert p.returncode == 0, (p.stdout[-3000:], p.stderr[-3000:])
    assert "MULTIGPU OK" in p.stdout
