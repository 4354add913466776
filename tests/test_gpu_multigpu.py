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


def test_driver_spreads_integrations_over_devices(native_built, monkeypatch):
    """Config 3 through the public API (calibration.py:1160-1167: the (polarization, time) units are independent): with two
    devices calibrate_and_model_dpss fits the two integrations of a 2-time fixture concurrently, one plan per device, and
    returns exactly what the single-device driver returns."""
    import numpy as np
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    from calamity_b200 import calibration
    from tests import fixtures_uv as fx

    outs = []
    for env in ({"CALAMITY_B200_DEVICE": "0"}, {"CALAMITY_B200_DEVICES": "0,1"}):
        monkeypatch.delenv("CALAMITY_B200_DEVICE", raising=False)
        monkeypatch.delenv("CALAMITY_B200_DEVICES", raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        uvd = fx.line_array(ntimes=2)
        uvd.data_array[uvd.Nbls :] *= 1.3  # the two integrations differ
        data = fx.add_noise_like_eor(fx.project_on_dpss(uvd, fx.dpss_vectors(uvd)))
        gains = fx.randomized_gains(data)
        assert len(calibration._device_indices(2, False)) == len(env.get("CALAMITY_B200_DEVICES", "0").split(","))
        model, resid, gains_out, hist = calibration.calibrate_and_model_dpss(
            min_dly=2.0 / 0.3, offset=2.0 / 0.3, uvdata=data, gains=gains, sky_model=None, maxsteps=200, tol=0.0,
            learning_rate=1e-2, correct_resid=True, correct_model=True)
        outs.append((model.data_array.copy(), resid.data_array.copy(), gains_out.gain_array.copy(),
                     np.asarray(hist[0][0]["loss"]), np.asarray(hist[0][1]["loss"])))
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
