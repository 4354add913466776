"""CPU: the analytic gradient of the NumPy restatement agrees with torch autograd through the op-for-op
port of the TensorFlow graph (calibration.py:1587-1656, 664-666) and with finite differences."""
import numpy as np
import pytest
import torch

from oracle import restatement as R
from oracle import torch_port as T
from tests.helpers import reference_tensors, rel_err, small_problem


@pytest.mark.parametrize("reg", [None, "sum"])
def test_analytic_gradient_matches_autograd(reg):
    prob = small_problem("test6", init_gain_scatter=0.05, coeff_error=0.1, flag_fraction=0.1)
    t = reference_tensors(prob, np.float64)
    sky_r, sky_i = t["data_r"], t["data_i"]
    pr, pi = R.sum_priors(sky_r, sky_i, t["wgts"], np.float64)
    loss, dgr, dgi, dfr, dfi = R.loss_and_grads(
        t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"], t["corr_inds"],
        regularization=reg, prior_r_sum=pr * 0.9, prior_i_sum=pi * 1.1)
    tp = T.TorchProblem(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"],
                        t["corr_inds"], model_regularization=reg, sky_model_r=sky_r, sky_model_i=sky_i,
                        dtype=torch.float64)
    if reg == "sum":
        tp.prior_r = tp.prior_r * 0.9
        tp.prior_i = tp.prior_i * 1.1
    tl, tg = tp.grads()
    assert abs(float(tl) - float(loss)) <= 1e-12 * abs(float(loss))
    assert rel_err(dgr, tg[0].numpy()) < 1e-10
    assert rel_err(dgi, tg[1].numpy()) < 1e-10
    nchunks = len(dfr)
    for c in range(nchunks):
        assert rel_err(dfr[c], tg[2 + c].numpy()) < 1e-10
        assert rel_err(dfi[c], tg[2 + nchunks + c].numpy()) < 1e-10
    # loss_value is the same number
    lv = R.loss_value(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"],
                      t["corr_inds"], regularization=reg, prior_r_sum=pr * 0.9, prior_i_sum=pi * 1.1)
    assert abs(float(lv) - float(loss)) <= 1e-14 * abs(float(loss))


def test_finite_difference_spot_check():
    prob = small_problem("test6", init_gain_scatter=0.05, coeff_error=0.1)
    t = reference_tensors(prob, np.float64)
    args = (t["data_r"], t["data_i"], t["wgts"], t["fg_comps"], t["corr_inds"])
    loss, dgr, dgi, dfr, dfi = R.loss_and_grads(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], *args)
    eps = 1e-6
    for (a, f) in [(0, 3), (4, 150)]:
        gp = t["g_r"].copy(); gp[a, f] += eps
        gm = t["g_r"].copy(); gm[a, f] -= eps
        fd = (R.loss_value(gp, t["g_i"], t["fg_r"], t["fg_i"], *args) - R.loss_value(gm, t["g_i"], t["fg_r"], t["fg_i"], *args)) / (2 * eps)
        assert abs(fd - dgr[a, f]) < 1e-5 * max(abs(dgr[a, f]), np.abs(dgr).max() * 1e-3)
    fr = [x.copy() for x in t["fg_r"]]
    fr[0][2, 7, 0, 0] += eps
    fd = (R.loss_value(t["g_r"], t["g_i"], fr, t["fg_i"], *args) - loss) / eps
    assert abs(fd - dfr[0][2, 7, 0, 0]) < 1e-4 * np.abs(dfr[0]).max()


@pytest.mark.parametrize("optimizer", ["Adamax", "Adam"])
def test_numpy_and_torch_fit_loops_agree(optimizer):
    """Two independent restatements of the loop (analytic gradient + NumPy rules vs autograd + torch rules)."""
    prob = small_problem("test6", init_gain_scatter=0.02, coeff_error=0.05)
    t = reference_tensors(prob, np.float64)
    kw = dict(maxsteps=25, tol=0.0, optimizer=optimizer, learning_rate=1e-2)
    a = R.fit(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"],
              t["corr_inds"], **kw)
    b = T.fit(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"],
              t["corr_inds"], dtype=torch.float64, **kw)
    assert np.allclose(a[4]["loss"], b[4]["loss"], rtol=1e-9, atol=0)
    assert rel_err(a[0], b[0]) < 1e-9 and rel_err(a[2][0], b[2][0]) < 1e-9
