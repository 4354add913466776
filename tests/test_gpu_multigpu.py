"""GPU (>= 2 devices): baseline-group sharding reproduces the single-GPU fit, with the per-iteration exchange done
either through NVLink peer memory (fused into the update kernels) or with the in-loop NCCL all-reduce.
Skipped on single-GPU boxes; the same data flow is covered on CPU by tests/test_cpu_sharding_gloo.py."""
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("comm", ["peer", "nccl"])
@pytest.mark.parametrize("reg", ["none", "sum"])
def test_sharded_fit_matches_single_gpu(native_built, reg, comm):
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py"), "hera37", reg, comm]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0 and "PASS" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


def test_diverged_rank_times_out_instead_of_hanging(native_built):
    """A rank that stops publishing (here: fewer steps on rank 1) makes calb2_fit return CALB2_ERR_TIMEOUT on the peer
    within the configured bound; no GPU is left spinning."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multigpu_check.py"), "hera37", "none", "peer", "timeout"]
    env = dict(os.environ, CALB2_PEER_TIMEOUT_MS="3000")
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert out.returncode == 0 and "PASS rank 0" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]
