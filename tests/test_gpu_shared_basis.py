"""GPU: the shared-basis kernel (calfit_shared.cuh -- each distinct basis stored once, register-tiled contraction over
the groups that share it) against the float64 oracle and against the streaming kernel, on every entry point that runs
a basis pass: loss / gradient, the fit loop (both regularisations), coefficient initialisation, model visibilities,
SNR weights, redundant (multi-baseline) slots, determinism.  Scale cases live in tests/test_gpu_scale_parity.py."""
import numpy as np
import pytest

from calamity_b200.layout import RaggedLayout
from oracle.ragged import RaggedProblem
from tests.helpers import FlatProblem, flat_from_synth, rel_err, small_problem

pytestmark = pytest.mark.gpu
F = np.float64


def _plan(p, tc=True, **kw):
    import os

    from calamity_b200.fitter import FitPlan

    os.environ["CALB2_TC"] = "1" if tc else "0"  # read by calb2_plan_create: tensor-core shape on (default) / off
    os.environ["CALB2_TC_MIN"] = "1"  # small test classes too (the default keeps classes of < 16 groups on the CUDA cores)
    try:
        plan = FitPlan(p.lay, device=0, **kw)
    finally:
        os.environ.pop("CALB2_TC", None)
        os.environ.pop("CALB2_TC_MIN", None)
    plan.set_integration(p.data_r, p.data_i, p.wgts)
    plan.set_gains(p.g0_r, p.g0_i)
    plan.set_coeffs(p.c0_r, p.c0_i)
    return plan


def _args64(p):
    return [np.asarray(x, dtype=F) for x in (p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r, p.data_i, p.wgts)]


def redundant_shared_problem(nants=20, nfreqs=96, seed=2):
    """Single-slot groups of 1-3 redundant baselines each (use_redundancy=True style) that share a handful of bases of
    different sizes (including one that is not a multiple of 8, one larger than 16), plus a multi-slot group that must
    stay on the streaming path."""
    rng = np.random.default_rng(seed)
    pairs = [(i, j) for i in range(nants) for j in range(i + 1, nants)]
    order = list(rng.permutation(len(pairs)))
    bases = [np.linalg.qr(rng.standard_normal((nfreqs, n)))[0] for n in (5, 8, 19, 33)]
    comps = {}
    for n in range(60):
        nb = 1 + n % 3
        red = tuple(pairs[order.pop()] for _ in range(nb))
        comps[(red,)] = bases[n % len(bases)]
    q, _ = np.linalg.qr(rng.standard_normal((2 * nfreqs, 7)))
    comps[((pairs[order.pop()], pairs[order.pop()]), (pairs[order.pop()],))] = q
    from calamity_b200.calibration import chunk_fg_comp_dict_by_nbls

    chunked = chunk_fg_comp_dict_by_nbls(comps, use_redundancy=True)
    lay = RaggedLayout.from_chunked_dict(chunked, {a: a for a in range(nants)}, nfreqs, nants=nants)
    p = FlatProblem()
    p.lay, p.nants, p.nfreqs, p.nbls = lay, nants, nfreqs, lay.nbls
    p.data_r = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
    p.data_i = rng.standard_normal((lay.nbls, nfreqs)).astype(np.float32)
    w = (rng.random((lay.nbls, nfreqs)) > 0.1).astype(F)
    p.wgts = (w / w.sum()).astype(np.float32)
    p.g0_r = (1.0 + 0.05 * rng.standard_normal((nants, nfreqs))).astype(np.float32)
    p.g0_i = (0.05 * rng.standard_normal((nants, nfreqs))).astype(np.float32)
    p.c0_r = rng.standard_normal(lay.ncoef).astype(np.float32)
    p.c0_i = rng.standard_normal(lay.ncoef).astype(np.float32)
    return p


def _cases():
    yield "test6-flags", flat_from_synth(small_problem("test6", init_gain_scatter=0.05, coeff_error=0.1, flag_fraction=0.2)), 1
    yield "hera37", flat_from_synth(small_problem("hera37", init_gain_scatter=0.05, coeff_error=0.1)), 0
    yield "hera37-all", flat_from_synth(small_problem("hera37", init_gain_scatter=0.05, coeff_error=0.1)), 1
    yield "redundant", redundant_shared_problem(), 1


@pytest.mark.parametrize("tc", [True, False], ids=["tensor-cores", "cuda-cores"])
@pytest.mark.parametrize("reg", [None, "sum"])
def test_loss_and_gradient_against_oracle(native_built, reg, tc):
    for name, p, mode in _cases():
        rp = RaggedProblem(p.lay)
        pr = float(np.sum(p.data_r.astype(F) * p.wgts)) * 0.9
        pi = float(np.sum(p.data_i.astype(F) * p.wgts)) * 1.1
        ol, ogr, ogi, ocr, oci = rp.loss_and_grads(*_args64(p), regularization=reg, prior_r_sum=F(pr), prior_i_sum=F(pi))
        plan = _plan(p, tc=tc, shared_basis=mode)
        info = dict(plan.info)
        loss, dgr, dgi, dcr, dci = plan.loss_and_grads(model_regularization=reg, prior_r_sum=pr, prior_i_sum=pi)
        plan.close()
        assert info["n_class_slots"] > 0, name
        # the tensor-core shape takes single-baseline classes of <= 128 vectors ('redundant' has multi-baseline slots)
        assert (info["n_tc_slots"] > 0) == (tc and name != "redundant"), (name, info["n_tc_slots"])
        if mode == 1 and name != "redundant":
            assert info["n_class_slots"] == info["nslots_total"] and info["nitems"] == 0, name
        errs = (abs(float(loss) - float(ol)) / abs(float(ol)), rel_err(dgr, ogr), rel_err(dgi, ogi), rel_err(dcr, ocr),
                rel_err(dci, oci))
        print(f"\n  {name} reg={reg}: classes {info['n_classes']}, class slots {info['n_class_slots']}/{info['nslots_total']}, "
              f"errs {errs[0]:.1e} {errs[1]:.1e} {errs[2]:.1e} {errs[3]:.1e} {errs[4]:.1e}")
        assert errs[0] <= 1e-5, name
        assert max(errs[1:]) < 1e-4, name


@pytest.mark.parametrize("tc", [True, False], ids=["tensor-cores", "cuda-cores"])
@pytest.mark.parametrize("optimizer,reg", [("Adamax", None), ("Adamax", "sum"), ("Adam", "sum")])
def test_fit_trajectory_against_oracle(native_built, optimizer, reg, tc):
    nsteps = 60
    for name, p, mode in list(_cases())[1:]:
        if name == "redundant":  # random data: scale the step so the loss falls smoothly
            lr = 1e-3
        else:
            lr = 1e-2
        rp = RaggedProblem(p.lay)
        pr = float(np.sum(p.data_r.astype(F) * p.wgts))
        pi = float(np.sum(p.data_i.astype(F) * p.wgts))
        kw = dict(optimizer=optimizer, maxsteps=nsteps, tol=0.0, learning_rate=lr, model_regularization=reg)
        o = rp.fit(p.g0_r, p.g0_i, p.c0_r, p.c0_i, p.data_r, p.data_i, p.wgts, prior_r_sum=F(np.float32(pr)),
                   prior_i_sum=F(np.float32(pi)), **kw)
        plan = _plan(p, tc=tc, shared_basis=mode)
        hist, res = plan.fit(prior_r_sum=pr, prior_i_sum=pi, **kw)
        g_r, g_i = plan.get_gains()
        c_r, c_i = plan.get_coeffs()
        plan.close()
        ref = np.asarray(o[4]["loss"], dtype=F)
        err = np.abs(hist.astype(F) - ref) / ref
        perr = (rel_err(g_r, o[0]), rel_err(g_i, o[1]), rel_err(c_r, o[2]), rel_err(c_i, o[3]))
        print(f"\n  {name} {optimizer}/{reg}: loss {ref[0]:.3e} -> {ref[-1]:.3e}, max rel loss err {err.max():.1e}, params {max(perr):.1e}")
        assert res["nsteps_recorded"] == nsteps and err.max() < 1e-5 and max(perr) < 1e-4, name


def test_shared_and_streaming_paths_agree_everywhere(native_built):
    """Same inputs through both kernels: loop semantics (use_min, tol), model visibilities, lstsq initialisation, SNR
    weights, freeze_model; and the shared path is bit-stable run to run."""
    p = flat_from_synth(small_problem("hera37", init_gain_scatter=0.02, coeff_error=0.05, flag_fraction=0.1))
    outs = {}
    for mode in (-1, 1, 1):
        plan = _plan(p, shared_basis=mode)
        plan.init_coeffs(p.data_r, p.data_i)
        c_init = plan.get_coeffs()
        m_init = plan.get_model()
        plan.apply_model_snr_weights()
        w_snr = plan.get_weights()
        hist, res = plan.fit(optimizer="Adamax", maxsteps=40, tol=0.0, learning_rate=1e-2, use_min=True, n_profile_steps=2)
        g = plan.get_gains()
        c = plan.get_coeffs()
        m = plan.get_model()
        hist_f, _ = plan.fit(optimizer="Adamax", maxsteps=20, tol=0.0, learning_rate=1e-2, freeze_model=True)
        c_frozen = plan.get_coeffs()
        hist_tol, res_tol = plan.fit(optimizer="Adamax", maxsteps=300, tol=1e-7, learning_rate=1e-2)
        plan.close()
        cur = dict(c_init=c_init, m_init=m_init, w_snr=(w_snr,), hist=(hist,), g=g, c=c, m=m, hist_f=(hist_f,), c_frozen=c_frozen)
        assert np.array_equal(c_frozen[0], c[0])
        assert len(hist_tol) < 300 and res_tol["nsteps_total"] == len(hist_tol) + 1
        if mode in outs:  # second run on the shared path: bit-identical
            for key, arrs in cur.items():
                for a, b in zip(arrs, outs[mode][key]):
                    assert np.array_equal(a, b), key
        outs[mode] = cur
    for key in outs[-1]:
        for a, b in zip(outs[-1][key], outs[1][key]):
            # two float32 implementations with different summation orders: the parity tolerances of BASELINE.json
            assert rel_err(a, b) < (1e-5 if key.startswith("hist") else 1e-4), (key, rel_err(a, b))


def test_plan_info_counts_the_deduplication(native_built):
    from calamity_b200.fitter import FitPlan

    prob = small_problem("hera37")
    lay = prob.layout()
    distinct = {id(b) for b in lay.blocks}
    with FitPlan(lay, device=0, shared_basis=1) as plan:
        info = dict(plan.info)
    assert info["n_classes"] == len(distinct)
    assert info["n_a_class_nz"] == info["n_a_nz"] == lay.sizes()["n_a_nz"]
    assert info["n_a_class"] < info["n_a_nz"] / 5
    with FitPlan(lay, device=0, shared_basis=-1) as plan:
        assert plan.info["n_classes"] == 0 and plan.info["n_a_stored"] >= info["n_a_nz"]
    # a deep copy of the dict (distinct objects, equal contents) still dedups: classes are found by content
    import copy

    lay2 = copy.deepcopy(lay)
    lay2.blocks = [b.copy() for b in lay2.blocks]
    lay2._finalize()
    assert len(set(lay2.group_class.tolist())) == len(distinct)


@pytest.mark.parametrize("reg", [None, "sum"])
def test_graph_replay_of_the_tensor_core_pass_is_bit_identical(native_built, reg):
    """Small problems replay a captured CUDA graph of iterations (calb2_fit, use_graph): with the tensor-core kernel(s), the
    streaming items forked onto the second stream and the forked coefficient update inside the capture, the replay must give
    the bits of the kernel-by-kernel launch."""
    from calamity_b200 import synth
    from calamity_b200.fitter import FitPlan

    prob = synth.make("hera37", init_gain_scatter=0.02, coeff_error=0.05)
    outs = []
    for graph in (False, True):
        plan = FitPlan(prob.layout(), device=0)
        assert plan.info["n_tc_ctas"] > 0  # classes of >= 16 groups: the default sends them to the tensor cores
        plan.set_integration(prob.data_r, prob.data_i, prob.wgts)
        plan.set_gains(prob.g0_r, prob.g0_i)
        plan.set_coeffs(prob.c0_r, prob.c0_i)
        pr, pi = plan.prior_sums(prob.data_r, prob.data_i)
        hist, _ = plan.fit(optimizer="Adamax", maxsteps=100, tol=0.0, learning_rate=1e-2, model_regularization=reg,
                           prior_r_sum=0.9 * pr, prior_i_sum=1.1 * pi, use_graph=graph)
        outs.append((hist, plan.get_gains()[0], plan.get_coeffs()[0]))
        plan.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)
    assert outs[0][0][-1] < 0.05 * outs[0][0][0]
