"""GPU, >= 2 devices: the data-parallel training step -- gradient exchange fused into the
optimizer over NVLink peer memory (one-shot push, two-shot push, one-shot reads) vs the NCCL
all-reduce (tools/dp_check.py under torchrun): every rank ends with identical weights, and they
match the NCCL path's.
Skipped on single-GPU boxes; the host-side DP logic is covered over gloo in test_host_logic.py."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_peer_memory_step_matches_nccl_step_on_two_gpus():
    out = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
         "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "dp_check.py")],
        capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    assert out.stdout.count("ranks identical: True") >= 3 and "ranks identical: False" not in out.stdout, \
        out.stdout[-2000:]
    rel = float(out.stdout.split("rel diff")[1].split(";")[0])
    assert rel < 1e-5, out.stdout[-500:]
    # ranks that drew different initial weights end up identical (rank 0's init is broadcast)
    assert "unseeded init, 8 steps: replicas identical: True" in out.stdout, out.stdout[-500:]
