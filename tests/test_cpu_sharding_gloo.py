"""CPU, world_size 2 over gloo: the N > 1 data flow of SURVEY.md section 8e-ii.  Each rank evaluates the loss
and gradient of ITS shard of the baseline groups with the oracle, the ranks all-reduce the gain gradient and the
scalar sums, and the result must equal the unsharded evaluation; coefficient gradients stay rank-private."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from calamity_b200 import synth
from calamity_b200.sharding import make_shard
from oracle import restatement as R


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _tensors(lay, prob, shard=None):
    take_b = (lambda x: x) if shard is None else shard.take_baselines
    take_c = (lambda x: x) if shard is None else shard.take_coeffs
    f64 = np.float64
    # a shard has no chunk structure of its own: one chunk with nbls == 1 per group (DPSS layout)
    nvecs = int(lay.group_ncomp.max())
    lay.chunks = [dict(nvecs=nvecs, ngrps=lay.ngroups, nbls=1, group0=0)]
    return dict(
        g_r=prob.g0_r.astype(f64), g_i=prob.g0_i.astype(f64),
        fg_r=lay.unflatten_coeffs(take_c(prob.c0_r), dtype=f64), fg_i=lay.unflatten_coeffs(take_c(prob.c0_i), dtype=f64),
        data_r=lay.unflatten_data(take_b(prob.data_r), f64), data_i=lay.unflatten_data(take_b(prob.data_i), f64),
        wgts=lay.unflatten_data(take_b(prob.wgts), f64), fg_comps=lay.dense_chunks(f64), corr_inds=lay.corr_inds())


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    prob = synth.make("test6", init_gain_scatter=0.05, coeff_error=0.1, flag_fraction=0.1)
    full = prob.layout()
    shard = make_shard(full, rank, world, mode=os.environ.get("CALB2_TEST_SHARD_MODE", "cyclic"))
    t = _tensors(shard.layout, prob, shard)
    # chi^2 part and the regulariser sums of this rank's groups
    loss, dgr, dgi, dfr, dfi = R.loss_and_grads(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"],
                                                t["wgts"], t["fg_comps"], t["corr_inds"])
    buf = torch.from_numpy(np.concatenate([dgr.ravel(), dgi.ravel(), [float(loss)]]))
    dist.all_reduce(buf)
    n = dgr.size
    if rank == 0:
        np.savez(out, dgr=buf[:n].numpy().reshape(dgr.shape), dgi=buf[n : 2 * n].numpy().reshape(dgi.shape),
                 loss=buf[-1].item(), dfr0=shard.layout.flatten_coeffs(dfr), cidx=shard.coef_index)
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("mode", ["cyclic", "contiguous"])
def test_sharded_gradient_equals_unsharded(tmp_path, mode, monkeypatch):
    monkeypatch.setenv("CALB2_TEST_SHARD_MODE", mode)
    out = str(tmp_path / "rank0.npz")
    port = _free_port()
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = np.load(out)
    prob = synth.make("test6", init_gain_scatter=0.05, coeff_error=0.1, flag_fraction=0.1)
    full = prob.layout()
    t = _tensors(full, prob)
    loss, dgr, dgi, dfr, dfi = R.loss_and_grads(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"],
                                                t["wgts"], t["fg_comps"], t["corr_inds"])
    assert abs(got["loss"] - float(loss)) < 1e-12 * abs(float(loss))
    assert np.allclose(got["dgr"], dgr, rtol=1e-10, atol=1e-14) and np.allclose(got["dgi"], dgi, rtol=1e-10, atol=1e-14)
    flat = full.flatten_coeffs(dfr).astype(np.float64)
    assert np.allclose(got["dfr0"], flat[got["cidx"]], rtol=1e-6)


def test_shards_tile_the_problem():
    prob = synth.make("hera37")
    full = prob.layout()
    for world in (2, 3, 8):
        shards = [make_shard(full, r, world, mode="contiguous") for r in range(world)]
        assert shards[0].g0 == 0 and shards[-1].g1 == full.ngroups
        assert sum(s.layout.nbls for s in shards) == full.nbls
        assert sum(s.layout.ncoef for s in shards) == full.ncoef
        got = np.concatenate([s.take_baselines(prob.data_r) for s in shards])
        assert np.array_equal(got, prob.data_r)
        loads = np.array([int(s.layout.group_ncomp.sum()) + 10 * s.layout.nbls for s in shards])
        assert loads.max() - loads.min() <= int(full.group_ncomp.max()) + 10
        # cyclic: every group exactly once, loads within a few groups of each other, antenna degrees even
        shards = [make_shard(full, r, world) for r in range(world)]
        assert np.array_equal(np.sort(np.concatenate([s.groups for s in shards])), np.arange(full.ngroups))
        assert np.array_equal(np.sort(np.concatenate([s.bl_index for s in shards])), np.arange(full.nbls))
        assert np.array_equal(np.sort(np.concatenate([s.coef_index for s in shards])), np.arange(full.ncoef))
        loads = np.array([int(s.layout.group_ncomp.sum()) + 10 * s.layout.nbls for s in shards])
        assert loads.max() - loads.min() <= 0.05 * loads.mean() + int(full.group_ncomp.max()) + 10
        deg = np.array([np.bincount(np.concatenate([s.layout.bl_ant0, s.layout.bl_ant1]), minlength=full.nants) for s in shards])
        assert deg.max() <= 2 * np.ceil((full.nants - 1) / world) + 2  # contiguous ranges: up to nants - 1 on one rank
