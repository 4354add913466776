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


def test_keras_rules_against_torch_optim_where_the_algebra_coincides():
    """Independent check of the restated Keras update rules: torch.optim implements the same published algorithms for
    SGD(+momentum/nesterov), RMSprop(momentum=0), Adagrad and Adadelta (up to the documented differences in where
    epsilon sits, which are matched here by construction), so a few steps on random float64 data must agree."""
    import torch

    from oracle.restatement import KerasOptimizer

    rng = np.random.default_rng(3)
    x0 = rng.standard_normal(50)
    grads = [rng.standard_normal(50) for _ in range(6)]

    def run_keras(name, **kw):
        p = x0.copy()
        opt = KerasOptimizer(name, **kw)
        for g in grads:
            opt.apply([p], [g.copy()], [False])
        return p

    def run_torch(make):
        p = torch.tensor(x0.copy(), requires_grad=True)
        opt = make([p])
        for g in grads:
            p.grad = torch.tensor(g.copy())
            opt.step()
        return p.detach().numpy()

    # Keras momentum accumulates -lr*g (velocity in parameter units); torch accumulates g and scales by lr: identical
    # for a constant learning rate
    np.testing.assert_allclose(run_keras("SGD", learning_rate=0.1, momentum=0.9),
                               run_torch(lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9)), rtol=1e-12)
    np.testing.assert_allclose(run_keras("SGD", learning_rate=0.1, momentum=0.9, nesterov=True),
                               run_torch(lambda ps: torch.optim.SGD(ps, lr=0.1, momentum=0.9, nesterov=True)), rtol=1e-12)
    np.testing.assert_allclose(run_keras("RMSprop", learning_rate=0.01, rho=0.9, epsilon=1e-7),
                               run_torch(lambda ps: torch.optim.RMSprop(ps, lr=0.01, alpha=0.9, eps=1e-7)), rtol=1e-12)
    np.testing.assert_allclose(run_keras("Adagrad", learning_rate=0.1, initial_accumulator_value=0.1, epsilon=1e-7),
                               run_torch(lambda ps: torch.optim.Adagrad(ps, lr=0.1, initial_accumulator_value=0.1, eps=1e-7)),
                               rtol=1e-12)
    np.testing.assert_allclose(run_keras("Adadelta", learning_rate=1.0, rho=0.95, epsilon=1e-7),
                               run_torch(lambda ps: torch.optim.Adadelta(ps, lr=1.0, rho=0.95, eps=1e-7)), rtol=1e-12)


def test_keras_nadam_and_ftrl_closed_forms():
    """Nadam's first step and Ftrl with l1 = l2 = 0 have closed forms that pin the restated rules."""
    from oracle.restatement import KerasOptimizer

    g = np.array([0.3, -2.0, 1e-3])
    p = np.array([1.0, -1.0, 0.5])
    opt = KerasOptimizer("Nadam", learning_rate=0.01)
    q = p.copy()
    opt.apply([q], [g.copy()], [False])
    b1, b2, eps = 0.9, 0.999, 1e-7
    u1 = b1 * (1 - 0.5 * 0.96 ** 0.004)
    u2 = b1 * (1 - 0.5 * 0.96 ** 0.008)
    m = (1 - b1) * g
    v = (1 - b2) * g * g
    mbar = (1 - u1) * g / (1 - u1) + u2 * (m / (1 - u1 * u2))
    np.testing.assert_allclose(q, p - 0.01 * mbar / (np.sqrt(v / (1 - b2)) + eps), rtol=1e-12)
    # Ftrl, no regularisation: theta_t = -lr * sum(z) / sqrt(acc) with the per-coordinate adaptive rate
    opt = KerasOptimizer("Ftrl", learning_rate=0.1)
    q = p.copy()
    acc = np.full(3, 0.1)
    lin = np.zeros(3)
    for _ in range(3):
        acc_new = acc + g * g
        lin += g - (np.sqrt(acc_new) - np.sqrt(acc)) / 0.1 * q
        expect = -lin / (np.sqrt(acc_new) / 0.1)
        acc = acc_new
        opt.apply([q], [g.copy()], [False])
        np.testing.assert_allclose(q, expect, rtol=1e-12)


def test_lamb_closed_form_and_per_variable_ratio():
    """tensorflow_addons LAMB (calibration.py:15, 26): after the bias correction the first update is g / (|g| + eps) (+ weight
    decay), scaled per VARIABLE by ||w|| / ||update||; a zero variable or a zero update takes ratio 1."""
    from oracle.restatement import KerasOptimizer

    lr, eps, wd = 0.01, 1e-6, 0.05
    g = [np.array([0.3, -2.0, 1e-3]), np.array([5.0, 5.0])]
    p = [np.array([1.0, -1.0, 0.5]), np.array([0.0, 0.0])]
    opt = KerasOptimizer("LAMB", learning_rate=lr, weight_decay=wd)
    q = [x.copy() for x in p]
    opt.apply(q, [x.copy() for x in g], [False, False])
    for v in range(2):
        upd = g[v] / (np.abs(g[v]) + eps) + wd * p[v]
        wn, un = np.linalg.norm(p[v]), np.linalg.norm(upd)
        ratio = wn / un if wn > 0 and un > 0 else 1.0
        np.testing.assert_allclose(q[v], p[v] - lr * ratio * upd, rtol=1e-12)
    assert np.linalg.norm(p[1]) == 0 and not np.allclose(q[1], 0)  # ratio 1 for the zero variable, it still moves
    # the ratio couples the elements of ONE variable only: the same numbers as a single variable give another step
    opt1 = KerasOptimizer("LAMB", learning_rate=lr, weight_decay=wd)
    q1 = np.concatenate(p)
    opt1.apply([q1], [np.concatenate(g)], [False])
    assert not np.allclose(q1, np.concatenate(q))
