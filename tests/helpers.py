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
