"""N>1 on real GPUs (skipped on a single-GPU box; the block rule and the id hand-over are covered on CPU by
tests/test_abi_and_host.py, the host-side merge logic by tests/test_dist_gloo.py).  tools/multi_gpu_check.py drives the C ABI's own
multi-GPU entry points -- marxb200_comm_init(_file), marxb200_trace_sharded, marxb200_merge_events_begin/_end,
marxb200_tally_allreduce -- and compares the merged result with one GPU tracing the same rays."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(port, env_extra, nproc=2):
    env = dict(os.environ, **env_extra)
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc),
                          "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(ROOT, "tools", "multi_gpu_check.py")],
                         capture_output=True, text=True, timeout=900, env=env)
    assert out.returncode == 0 and "multi_gpu_check OK" in out.stdout, out.stdout[-3000:] + out.stderr[-3000:]
    return out.stdout


def _gpus():
    import torch
    return torch.cuda.device_count()


def test_two_gpus_reproduce_single_gpu_events_peer_writes():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    _run(29533, {"MGC_INIT": "bcast"})


def test_two_gpus_reproduce_single_gpu_events_nccl_send_recv_and_file_rendezvous():
    if _gpus() < 2:
        pytest.skip("needs 2 GPUs")
    out = _run(29534, {"MGC_INIT": "file", "MARXB200_MERGE_TRANSPORT": "nccl"})
    assert "ncclSend/ncclRecv" in out


def test_all_gpus_of_the_box():
    n = _gpus()
    if n < 3:
        pytest.skip("needs more than 2 GPUs")
    _run(29535, {"MGC_INIT": "bcast"}, nproc=n)
