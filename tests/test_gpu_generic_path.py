"""GPU: the generic (unfused, precision-templated) device path -- float64 fits (`dtype=np.float64`, calibration.py:974;
`--precision 64`, 1795) and float32 problems whose groups are too large for the fused kernel's staged tile.

float64 results are compared with the float64 restatement at float64 tolerances (1e-9), far inside BASELINE.json's
float32 tolerances; the float32 generic path must agree with the fused float32 kernels and with the oracle."""
import numpy as np
import pytest

from oracle import restatement as R
from tests.helpers import reference_tensors, rel_err, small_problem

pytestmark = pytest.mark.gpu


def _plan64(prob):
    from calamity_b200.fitter import FitPlan
    from calamity_b200.layout import RaggedLayout

    t = reference_tensors(prob, np.float64)
    lay = RaggedLayout.from_dense(t["fg_comps"], t["corr_inds"], prob.nants, dtype=np.float64)
    plan = FitPlan(lay, device=0)
    assert plan.info["generic"] == 1 and plan.info["dtype"] == 1
    plan.set_integration(lay.flatten_data(t["data_r"]), lay.flatten_data(t["data_i"]), lay.flatten_data(t["wgts"]))
    plan.set_gains(t["g_r"], t["g_i"])
    plan.set_coeffs(lay.flatten_coeffs(t["fg_r"]), lay.flatten_coeffs(t["fg_i"]))
    return plan, lay, t


@pytest.mark.parametrize("name,kw", [("test6", {}), ("test6", {"flag_fraction": 0.2}), ("hera37", {})])
@pytest.mark.parametrize("reg", [None, "sum"])
def test_float64_loss_and_gradient(native_built, name, kw, reg):
    prob = small_problem(name, init_gain_scatter=0.05, coeff_error=0.1, **kw)
    plan, lay, t = _plan64(prob)
    pr, pi = R.sum_priors(t["data_r"], t["data_i"], t["wgts"], np.float64)
    pr, pi = float(pr) * 0.9, float(pi) * 1.1
    ol, ogr, ogi, ofr, ofi = R.loss_and_grads(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"],
                                              t["wgts"], t["fg_comps"], t["corr_inds"], regularization=reg,
                                              prior_r_sum=pr, prior_i_sum=pi)
    loss, dgr, dgi, dcr, dci = plan.loss_and_grads(model_regularization=reg, prior_r_sum=pr, prior_i_sum=pi)
    plan.close()
    assert loss.dtype == np.float64 and dgr.dtype == np.float64
    assert abs(float(loss) - float(ol)) <= 1e-11 * abs(float(ol)), (loss, ol)
    assert rel_err(dgr, ogr) < 1e-10 and rel_err(dgi, ogi) < 1e-10
    assert rel_err(dcr, lay.flatten_coeffs(ofr)) < 1e-10 and rel_err(dci, lay.flatten_coeffs(ofi)) < 1e-10


@pytest.mark.parametrize("optimizer,opt_kw", [("Adamax", dict(learning_rate=1e-2)), ("Adam", dict(learning_rate=1e-2)),
                                              ("Nadam", dict(learning_rate=1e-2)), ("Adagrad", dict(learning_rate=1e-2))])
@pytest.mark.parametrize("reg", [None, "sum"])
def test_float64_fit_trajectory(native_built, optimizer, opt_kw, reg):
    nsteps = 60
    prob = small_problem("test6", init_gain_scatter=0.02, coeff_error=0.05)
    plan, lay, t = _plan64(prob)
    kw = dict(maxsteps=nsteps, tol=0.0, optimizer=optimizer, model_regularization=reg, **opt_kw)
    o = R.fit(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"], t["corr_inds"],
              sky_model_r=t["data_r"], sky_model_i=t["data_i"], **kw)
    pr, pi = R.sum_priors(t["data_r"], t["data_i"], t["wgts"], np.float64)
    hist, res = plan.fit(prior_r_sum=float(pr), prior_i_sum=float(pi), **kw)
    g_r, g_i = plan.get_gains()
    c_r, c_i = plan.get_coeffs()
    plan.close()
    assert hist.dtype == np.float64 and res["nsteps_recorded"] == nsteps and res["nsteps_total"] == nsteps + 1
    ref = np.asarray(o[4]["loss"], dtype=np.float64)
    err = np.abs(hist - ref) / ref
    print(f"\nfloat64 {optimizer}/{reg}: max rel loss err {err.max():.2e}")
    assert err.max() < 1e-9
    assert rel_err(g_r, o[0]) < 1e-8 and rel_err(g_i, o[1]) < 1e-8
    assert rel_err(c_r, lay.flatten_coeffs(o[2])) < 1e-8 and rel_err(c_i, lay.flatten_coeffs(o[3])) < 1e-8


def test_float64_loop_semantics_and_setup_stage(native_built):
    """tol stop, use_min, profile steps, freeze_model, lstsq initialisation and the model cube in float64."""
    prob = small_problem("test6", init_gain_scatter=0.02, coeff_error=0.05)
    plan, lay, t = _plan64(prob)
    kw = dict(optimizer="Adamax", learning_rate=1e-2)
    o = R.fit(t["g_r"], t["g_i"], t["fg_r"], t["fg_i"], t["data_r"], t["data_i"], t["wgts"], t["fg_comps"], t["corr_inds"],
              maxsteps=200, tol=1e-6, use_min=True, n_profile_steps=3, **kw)
    hist, res = plan.fit(maxsteps=200, tol=1e-6, use_min=True, n_profile_steps=3, **kw)
    assert len(hist) == len(o[4]["loss"]) < 200 and res["nsteps_total"] == len(hist) + 4
    np.testing.assert_allclose(hist, np.asarray(o[4]["loss"]), rtol=1e-9)
    g_r, _ = plan.get_gains()
    assert rel_err(g_r, o[0]) < 1e-8
    # freeze_model: coefficients untouched
    plan.set_gains(t["g_r"], t["g_i"])
    c0 = lay.flatten_coeffs(t["fg_r"])
    plan.set_coeffs(c0, lay.flatten_coeffs(t["fg_i"]))
    plan.fit(maxsteps=10, tol=0.0, freeze_model=True, **kw)
    np.testing.assert_array_equal(plan.get_coeffs()[0], c0)
    # least-squares initialisation (calibration.py:828-913) and model (402-444)
    ref_c = R.init_coeffs(t["data_r"], t["wgts"], t["fg_comps"])
    plan.init_coeffs(lay.flatten_data(t["data_r"]), lay.flatten_data(t["data_i"]))
    c_r, _ = plan.get_coeffs()
    assert rel_err(c_r, lay.flatten_coeffs(ref_c)) < 1e-9
    m_r, _ = plan.get_model()
    cube = R.model_cube(prob.nants, prob.nfreqs, t["fg_comps"], ref_c, t["corr_inds"])
    assert rel_err(m_r, cube[lay.bl_ant0, lay.bl_ant1]) < 1e-9
    plan.close()


@pytest.mark.parametrize("tc", [False, True], ids=["cuda-cores", "tensor-cores"])
def test_generic_float32_matches_fused_kernels(native_built, monkeypatch, tc):
    """CALB2_GENERIC=1 routes an ordinary float32 problem through the generic path: same trajectory as the fused kernels.
    Two float32 CUDA-core evaluations agree to 1e-5 in the parameters; the tensor-core shape (3-term TF32 split, truncating
    accumulator: gradients within ~5e-6 of float64) is held to the north-star bound of 1e-4."""
    from calamity_b200.fitter import FitPlan

    prob = small_problem("hera37", init_gain_scatter=0.02, coeff_error=0.05)
    kw = dict(optimizer="Adamax", maxsteps=40, tol=0.0, learning_rate=1e-2)
    monkeypatch.setenv("CALB2_TC", "1" if tc else "0")
    monkeypatch.setenv("CALB2_TC_MIN", "1")
    outs = []
    for generic in (0, 1):
        monkeypatch.setenv("CALB2_GENERIC", str(generic))
        plan = FitPlan(prob.layout(), device=0)
        assert plan.info["generic"] == generic
        if not generic:
            assert (plan.info["n_tc_ctas"] > 0) == tc
        plan.set_integration(prob.data_r, prob.data_i, prob.wgts)
        plan.set_gains(prob.g0_r, prob.g0_i)
        plan.set_coeffs(prob.c0_r, prob.c0_i)
        hist, _ = plan.fit(**kw)
        outs.append((hist, plan.get_gains()[0], plan.get_coeffs()[0]))
        plan.close()
    np.testing.assert_allclose(outs[1][0], outs[0][0], rtol=2e-6)
    ptol = 1e-4 if tc else 1e-5
    assert rel_err(outs[1][1], outs[0][1]) < ptol and rel_err(outs[1][2], outs[0][2]) < ptol
