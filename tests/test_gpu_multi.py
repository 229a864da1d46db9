"""Multi-GPU acceptance test of SURVEY.md §8e: the detection list gathered from N frame-sharded ranks over NCCL is
byte-identical to the single-GPU list.  Needs >= 2 GPUs (gpurun --gpus 2); skipped on a one-GPU box, where the same
logic is covered by the emulated-rank tests in test_gpu_properties.py and the gloo tests in test_sharding_gloo.py."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_nccl_gather_equals_single_gpu():
    import torch

    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs two GPUs")
    world = 4 if n >= 4 else 2
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", "29533", os.path.join(ROOT, "tests", "dist_gather_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "match=True" in r.stdout
