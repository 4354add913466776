"""CPU: the ragged, basis-batched float64 oracle (oracle/ragged.py) -- the yardstick of the scale-parity GPU tests --
is the dense restatement (oracle/restatement.py, calibration.py:1587-1656 / 447-738) on every small configuration,
including multi-baseline slots and multi-slot (joint) groups."""
import numpy as np
import pytest

from oracle import restatement as R
from oracle.ragged import RaggedProblem
from tests.helpers import flat_from_synth, mixed_problem, rel_err, small_problem


def _dense(p, dtype=np.float64):
    lay = p.lay
    return dict(g_r=p.g0_r.astype(dtype), g_i=p.g0_i.astype(dtype), fg_r=lay.unflatten_coeffs(p.c0_r, dtype=dtype),
                fg_i=lay.unflatten_coeffs(p.c0_i, dtype=dtype), data_r=lay.unflatten_data(p.data_r, dtype=dtype),
                data_i=lay.unflatten_data(p.data_i, dtype=dtype), wgts=lay.unflatten_data(p.wgts, dtype=dtype),
                fg_comps=lay.dense_chunks(dtype=dtype), corr_inds=lay.corr_inds())


def _flat64(lay, chunk_coeffs):
    """layout.flatten_coeffs without the cast to the layout's float32."""
    flat = np.zeros(lay.ncoef, dtype=np.float64)
    for ch, t in zip(lay.chunks, chunk_coeffs):
        t = np.asarray(t).reshape(ch["nvecs"], ch["ngrps"])
        for g in range(ch["ngrps"]):
            gg = ch["group0"] + g
            flat[lay.group_coef0[gg] : lay.group_coef0[gg] + lay.group_ncomp[gg]] = t[: lay.group_ncomp[gg], g]
    return flat


def _problems():
    yield "test6", flat_from_synth(small_problem("test6", init_gain_scatter=0.05, coeff_error=0.1, flag_fraction=0.1))
    yield "hera37", flat_from_synth(small_problem("hera37", init_gain_scatter=0.05, coeff_error=0.1))
    yield "mixed", mixed_problem(nants=12, nfreqs=48, seed=3, n_dpss_bls=20, joint=((3, 4, 7), (6, 3, 11)))


@pytest.mark.parametrize("reg", [None, "sum"])
def test_ragged_loss_and_gradient_equal_dense(reg):
    f = np.float64
    for name, p in _problems():
        t = _dense(p)
        pr, pi = R.sum_priors(t["data_r"], t["data_i"], t["wgts"], f)
        pr, pi = f(pr * 0.9), f(pi * 1.1)
        ol, ogr, ogi, ofr, ofi = R.loss_and_grads(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"],
                                                  t["wgts"], t["fg_comps"], t["corr_inds"], regularization=reg,
                                                  prior_r_sum=pr, prior_i_sum=pi)
        rp = RaggedProblem(p.lay)
        loss, dgr, dgi, dcr, dci = rp.loss_and_grads(*(x.astype(f) for x in (p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r,
                                                                             p.data_i, p.wgts)),
                                                     regularization=reg, prior_r_sum=pr, prior_i_sum=pi)
        assert abs(loss - ol) <= 1e-12 * abs(ol), name
        assert rel_err(dgr, ogr) < 1e-11 and rel_err(dgi, ogi) < 1e-11, name
        assert rel_err(dcr, _flat64(p.lay, ofr)) < 1e-11 and rel_err(dci, _flat64(p.lay, ofi)) < 1e-11, name


def test_ragged_fit_equals_dense_fit():
    f = np.float64
    name, p = list(_problems())[2]
    t = _dense(p)
    kw = dict(maxsteps=25, tol=0.0, optimizer="Adamax", learning_rate=1e-2)
    o = R.fit(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"],
              t["corr_inds"], sky_model_r=t["data_r"], sky_model_i=t["data_i"], model_regularization="sum", **kw)
    pr, pi = R.sum_priors(t["data_r"], t["data_i"], t["wgts"], f)
    rp = RaggedProblem(p.lay)
    r = rp.fit(p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r, p.data_i, p.wgts, model_regularization="sum",
               prior_r_sum=pr, prior_i_sum=pi, **kw)
    assert np.allclose(r[4]["loss"], o[4]["loss"], rtol=1e-11)
    assert rel_err(r[0], o[0]) < 1e-10 and rel_err(r[2], _flat64(p.lay, o[2])) < 1e-10
