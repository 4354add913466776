"""GPU: setup-stage entry points against the oracle -- lstsq coefficient initialisation
(calibration.py:828-913), regulariser priors (619-625), SNR weights (1235-1242)."""
import numpy as np
import pytest

from calamity_b200.layout import RaggedLayout
from oracle import restatement as R
from tests.helpers import reference_tensors, rel_err, small_problem

pytestmark = pytest.mark.gpu


def _redundant_layout(nfreqs=96, seed=3):
    """Two multi-baseline fitting groups (shared coefficients across 'redundant' baselines, non-orthonormal
    dense design matrix) plus single-baseline groups, like a use_redundancy=True / mixed-mode dict."""
    rng = np.random.default_rng(seed)
    comps = {}
    # fitting group with 2 redundant sub-groups of 3 and 2 baselines: basis spans 2*nfreqs
    q, _ = np.linalg.qr(rng.standard_normal((2 * nfreqs, 9)))
    comps[(((0, 1), (1, 2), (2, 3)), ((0, 2), (1, 3)))] = q
    q, _ = np.linalg.qr(rng.standard_normal((nfreqs, 5)))
    comps[(((0, 3), (1, 4)),)] = q
    for ap in [(0, 4), (2, 4), (3, 4)]:
        comps[((ap,),)] = rng.standard_normal((nfreqs, 4))  # deliberately NOT orthonormal
    return comps


def test_init_coeffs_matches_lstsq(native_built):
    from calamity_b200.calibration import chunk_fg_comp_dict_by_nbls
    from calamity_b200.fitter import FitPlan

    nfreqs, nants = 96, 5
    comps = _redundant_layout(nfreqs)
    ants_map = {a: a for a in range(nants)}
    for use_red in (True, False):
        chunked = chunk_fg_comp_dict_by_nbls(comps, use_redundancy=use_red)
        lay = RaggedLayout.from_chunked_dict(chunked, ants_map, nfreqs, nants=nants)
        rng = np.random.default_rng(11)
        sky_r = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
        sky_i = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
        w = (rng.random((lay.nbls, nfreqs)) > 0.15).astype(np.float32)
        w /= w.sum()
        with FitPlan(lay, device=0) as plan:
            plan.set_integration(sky_r, sky_i, w)
            plan.set_gains(np.ones((nants, nfreqs), np.float32), np.zeros((nants, nfreqs), np.float32))
            plan.init_coeffs(sky_r, sky_i)
            c_r, c_i = plan.get_coeffs()
            pr, pi = plan.prior_sums(sky_r, sky_i)
        dense = lay.dense_chunks(np.float64)
        want_r = R.init_coeffs(lay.unflatten_data(sky_r, np.float64), lay.unflatten_data(w, np.float64), dense)
        want_i = R.init_coeffs(lay.unflatten_data(sky_i, np.float64), lay.unflatten_data(w, np.float64), dense)
        assert rel_err(c_r, lay.flatten_coeffs(want_r)) < 1e-4
        assert rel_err(c_i, lay.flatten_coeffs(want_i)) < 1e-4
        assert abs(float(pr) - float(np.sum(sky_r.astype(np.float64) * w))) < 1e-6
        assert abs(float(pi) - float(np.sum(sky_i.astype(np.float64) * w))) < 1e-6


def test_redundant_groups_loss_and_gradient(native_built):
    """Multi-baseline slots and multi-slot groups (mixed-mode style) through the fused kernel."""
    from calamity_b200.calibration import chunk_fg_comp_dict_by_nbls
    from calamity_b200.fitter import FitPlan

    nfreqs, nants = 96, 5
    comps = _redundant_layout(nfreqs)
    ants_map = {a: a for a in range(nants)}
    chunked = chunk_fg_comp_dict_by_nbls(comps, use_redundancy=True)
    lay = RaggedLayout.from_chunked_dict(chunked, ants_map, nfreqs, nants=nants)
    rng = np.random.default_rng(5)
    d_r = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
    d_i = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
    w = rng.random((lay.nbls, nfreqs)).astype(np.float32)
    w /= w.sum()
    g_r = (1 + 0.1 * rng.standard_normal((nants, nfreqs))).astype(np.float32)
    g_i = (0.1 * rng.standard_normal((nants, nfreqs))).astype(np.float32)
    c_r = rng.standard_normal(lay.ncoef).astype(np.float32)
    c_i = rng.standard_normal(lay.ncoef).astype(np.float32)
    for reg in (None, "sum"):
        with FitPlan(lay, device=0) as plan:
            plan.set_integration(d_r, d_i, w)
            plan.set_gains(g_r, g_i)
            plan.set_coeffs(c_r, c_i)
            loss, dgr, dgi, dcr, dci = plan.loss_and_grads(model_regularization=reg, prior_r_sum=0.3, prior_i_sum=-0.2)
            m_r, m_i = plan.get_model()
        f64 = np.float64
        args = (g_r.astype(f64), g_i.astype(f64), lay.unflatten_coeffs(c_r, dtype=f64), lay.unflatten_coeffs(c_i, dtype=f64),
                lay.unflatten_data(d_r, f64), lay.unflatten_data(d_i, f64), lay.unflatten_data(w, f64),
                lay.dense_chunks(f64), lay.corr_inds())
        ol, ogr, ogi, ofr, ofi = R.loss_and_grads(*args, regularization=reg, prior_r_sum=0.3, prior_i_sum=-0.2)
        assert abs(float(loss) - float(ol)) <= 1e-5 * abs(float(ol))
        assert rel_err(dgr, ogr) < 1e-4 and rel_err(dgi, ogi) < 1e-4
        assert rel_err(dcr, lay.flatten_coeffs(ofr)) < 1e-4 and rel_err(dci, lay.flatten_coeffs(ofi)) < 1e-4
        cube = R.model_cube(nants, nfreqs, args[7], args[2], args[8])
        want = np.stack([cube[i, j] for i, j in zip(lay.bl_ant0, lay.bl_ant1)])
        assert rel_err(m_r, want) < 1e-5


def test_model_snr_weights(native_built):
    from calamity_b200.fitter import FitPlan

    prob = small_problem("test6", flag_fraction=0.1)
    t = reference_tensors(prob, np.float64)
    with FitPlan(prob.layout(), device=0) as plan:
        plan.set_integration(prob.data_r, prob.data_i, prob.wgts)
        plan.set_gains(prob.g0_r, prob.g0_i)
        plan.set_coeffs(prob.c0_r, prob.c0_i)
        plan.apply_model_snr_weights()
        w = plan.get_weights()
    wm = [R.fg_vis(fr, fi, c) for fr, fi, c in zip(t["fg_r"], t["fg_i"], t["fg_comps"])]
    new = [(np.square(v[0]) + np.square(v[1])) * w0 for v, w0 in zip(wm, t["wgts"])]
    tot = np.sum([np.sum(x) for x in new])
    want = t["lay"].flatten_data([x / tot for x in new])
    assert rel_err(w, want) < 1e-5
    assert abs(float(w.sum()) - 1.0) < 1e-5
