"""N>1 on real GPUs (skipped on a single-GPU box; the host logic is covered on CPU by tests/test_dist_gloo.py)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_two_gpus_reproduce_single_gpu_events():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "multi_gpu_check OK" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
