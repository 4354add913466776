"""Shared test helpers: reference-layout views of a synthetic problem for the oracle."""
import numpy as np

from calamity_b200 import synth


def reference_tensors(prob, dtype=np.float64):
    """Reference (dense, chunked) tensors of a SyntheticProblem, in `dtype`."""
    lay = prob.layout()
    return dict(
        lay=lay,
        g_r=prob.g0_r.astype(dtype), g_i=prob.g0_i.astype(dtype),
        fg_r=lay.unflatten_coeffs(prob.c0_r, dtype=dtype), fg_i=lay.unflatten_coeffs(prob.c0_i, dtype=dtype),
        data_r=lay.unflatten_data(prob.data_r, dtype=dtype), data_i=lay.unflatten_data(prob.data_i, dtype=dtype),
        wgts=lay.unflatten_data(prob.wgts, dtype=dtype),
        fg_comps=lay.dense_chunks(dtype=dtype), corr_inds=lay.corr_inds(),
    )


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-300))


def small_problem(name="test6", **kw):
    return synth.make(name, **kw)


class FlatProblem:
    """One integration in the library's flat layout: lay, data_r/data_i/wgts [nbls, nf], g0_r/g0_i, c0_r/c0_i."""


def mixed_problem(nants=128, nfreqs=1024, seed=5, n_dpss_bls=300, joint=((11, 43, 300),), nf_dpss_df=None):
    """Config-5-shaped input (BASELINE.json configs[4]; calibration.py:1353-1500, modeling.py:377-474): joint fitting
    groups with `nrg` redundant sub-groups of ~`nper` baselines each sharing `ncomp` coefficients over a basis that
    spans nrg * nfreqs samples (seeded orthonormal columns standing in for the covariance eigenvectors of
    simple_cov.py:100-182), next to per-baseline DPSS groups.  Built as the reference's dict and chunked by the
    product's own chunk_fg_comp_dict_by_nbls, so the layout is what calibrate_and_model_mixed would hand over."""
    from calamity_b200.calibration import chunk_fg_comp_dict_by_nbls
    from calamity_b200.layout import RaggedLayout
    from oracle.ragged import RaggedProblem

    rng = np.random.default_rng(seed)
    antpos = synth.hex_antenna_positions(nants)
    freqs = 100e6 + (100e6 / 1024.0) * np.arange(nfreqs)
    pairs = [(i, j) for i in range(nants) for j in range(i + 1, nants)]
    order = rng.permutation(len(pairs))
    comps, cursor = {}, 0
    for (nrg, nper, ncomp) in joint:
        red = []
        for r in range(nrg):
            n = max(1, nper + int(rng.integers(-3, 4)))  # ragged redundant groups
            red.append(tuple(pairs[k] for k in order[cursor : cursor + n]))
            cursor += n
        q, _ = np.linalg.qr(rng.standard_normal((nrg * nfreqs, ncomp)))
        comps[tuple(red)] = q
    dpss = synth.dpss_comps_dict(antpos, freqs)
    for k in order[cursor : cursor + n_dpss_bls]:
        comps[((pairs[k],),)] = dpss[((pairs[k],),)]
    chunked = chunk_fg_comp_dict_by_nbls(comps, use_redundancy=False)
    lay = RaggedLayout.from_chunked_dict(chunked, {a: a for a in range(nants)}, nfreqs, nants=nants)
    rp = RaggedProblem(lay)
    p = FlatProblem()
    p.lay, p.nants, p.nfreqs, p.nbls = lay, nants, nfreqs, lay.nbls
    c_true = rng.standard_normal(lay.ncoef) + 1j * rng.standard_normal(lay.ncoef)
    for g in range(lay.ngroups):  # decaying spectrum per group
        n = lay.group_ncomp[g]
        c_true[lay.group_coef0[g] : lay.group_coef0[g] + n] *= np.exp(-3.0 * np.arange(n) / max(n, 1))
    v_r, v_i = rp.slot_vis(c_true.real.copy(), c_true.imag.copy())
    vis = (v_r + 1j * v_i)[rp.bl_slot]
    g_true = 1.0 + 0.1 * (rng.standard_normal((nants, nfreqs)) + 1j * rng.standard_normal((nants, nfreqs)))
    data = g_true[lay.bl_ant0] * np.conj(g_true[lay.bl_ant1]) * vis
    data = data + 1e-4 * (rng.standard_normal(data.shape) + 1j * rng.standard_normal(data.shape))
    flags = rng.random(data.shape) < 0.05
    rms = np.sqrt(np.mean(np.abs(data[~flags]) ** 2.0))
    data = data / rms
    p.data_r = np.ascontiguousarray(data.real, dtype=np.float32)
    p.data_i = np.ascontiguousarray(data.imag, dtype=np.float32)
    w = (~flags).astype(np.float64)
    p.wgts = np.ascontiguousarray(w / w.sum(), dtype=np.float32)
    g0 = 1.0 + 0.02 * (rng.standard_normal((nants, nfreqs)) + 1j * rng.standard_normal((nants, nfreqs)))
    p.g0_r = np.ascontiguousarray(g0.real, dtype=np.float32)
    p.g0_i = np.ascontiguousarray(g0.imag, dtype=np.float32)
    c0 = c_true / rms * (1.0 + 0.05 * rng.standard_normal(lay.ncoef))
    p.c0_r = np.ascontiguousarray(c0.real, dtype=np.float32)
    p.c0_i = np.ascontiguousarray(c0.imag, dtype=np.float32)
    return p


def flat_from_synth(prob):
    """SyntheticProblem -> FlatProblem view (the attributes the scale-parity tests use)."""
    p = FlatProblem()
    p.lay, p.nants, p.nfreqs, p.nbls = prob.layout(), prob.nants, prob.nfreqs, prob.nbls
    for name in ("data_r", "data_i", "wgts", "g0_r", "g0_i", "c0_r", "c0_i"):
        setattr(p, name, getattr(prob, name))
    return p


def long_baseline_subset(prob, every=8, n_long=200):
    """Indices of the `n_long` baselines with the largest bases (the multi-warp slots / 8-slot items the fused kernel
    only meets at HERA-128/350 scale) plus every `every`-th baseline of the rest, in canonical order."""
    long_ones = np.argsort(-prob.ncomp, kind="stable")[:n_long]
    return np.unique(np.concatenate([long_ones, np.arange(0, prob.nbls, every)]))
