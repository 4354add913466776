"""Seeded inputs shared by tests/golden/make_golden.py (which feeds them to the reference's own code) and by the
tests that compare the oracle, the host logic and the CUDA path against the stored reference outputs."""
import numpy as np

from tests import fixtures_uv as fx
from tests.helpers import reference_tensors, small_problem

PRIOR_R, PRIOR_I = 0.0123, -0.0045


def mixed_dict(nfreqs=24, seed=0):
    rng = np.random.default_rng(seed)
    d = {}
    d[(((0, 1), (1, 2)), ((0, 2),))] = rng.standard_normal((2 * nfreqs, 5))
    d[(((2, 3), (3, 4)), ((1, 3), (2, 4)))] = rng.standard_normal((2 * nfreqs, 7))
    d[(((0, 4),),)] = rng.standard_normal((nfreqs, 3))
    d[(((0, 3),), ((1, 4),), ((0, 5),), ((1, 5),), ((2, 5),), ((3, 5),))] = rng.standard_normal((6 * nfreqs, 9))
    d[(((4, 5),),)] = rng.standard_normal((nfreqs, 4))
    return d, nfreqs


def reference_problem(dtype):
    prob = small_problem("test6", init_gain_scatter=0.03, coeff_error=0.08, flag_fraction=0.05)
    t = reference_tensors(prob, dtype)
    t["prob"] = prob
    return t


FIT_CASES = {
    "adamax": dict(optimizer="Adamax", maxsteps=40, tol=0.0, learning_rate=1e-2),
    "adam_sum": dict(optimizer="Adam", maxsteps=40, tol=0.0, learning_rate=1e-2, model_regularization="sum"),
    "adamax_sum_usemin_profile": dict(optimizer="Adamax", maxsteps=30, tol=0.0, learning_rate=0.2, use_min=True,
                                      n_profile_steps=2, model_regularization="sum", profile_log_dir="/tmp/calb2_shim_prof"),
    "adamax_freeze": dict(optimizer="Adamax", maxsteps=40, tol=0.0, learning_rate=1e-2, freeze_model=True),
    "adamax_tol": dict(optimizer="Adamax", maxsteps=500, tol=2e-6, learning_rate=1e-2),
    # the other entries of the reference's OPTIMIZERS table (calibration.py:17-27), hyper-parameters through **opt_kwargs
    "sgd_momentum": dict(optimizer="SGD", maxsteps=30, tol=0.0, learning_rate=2e-3, momentum=0.9, nesterov=True),
    "rmsprop": dict(optimizer="RMSprop", maxsteps=30, tol=0.0, learning_rate=1e-3, momentum=0.5),
    "adagrad_sum": dict(optimizer="Adagrad", maxsteps=30, tol=0.0, learning_rate=1e-2, model_regularization="sum"),
    "adadelta": dict(optimizer="Adadelta", maxsteps=30, tol=0.0, learning_rate=1.0),
    "nadam": dict(optimizer="Nadam", maxsteps=30, tol=0.0, learning_rate=1e-2),
    # tensorflow_addons' LAMB (calibration.py:15, 26): the reference's loop, the shim's statement of the tfa rule
    "lamb": dict(optimizer="LAMB", maxsteps=30, tol=0.0, learning_rate=1e-2, weight_decay=1e-2),
}


def random_chunk_data(corr_inds, nfreqs, seed):
    rng = np.random.default_rng(seed)
    return [rng.standard_normal((len(chunk), len(chunk[0]), nfreqs)) for chunk in corr_inds]


def random_chunk_weights(corr_inds, nfreqs, seed):
    rng = np.random.default_rng(seed)
    w = [(rng.random((len(chunk), len(chunk[0]), nfreqs)) > 0.15).astype(np.float64) for chunk in corr_inds]
    tot = sum(x.sum() for x in w)
    return [x / tot for x in w]


DRIVER_CASES = {
    "sum": dict(min_dly=2.0 / 0.3, offset=2.0 / 0.3, sky_model=None, maxsteps=60, tol=0.0, learning_rate=1e-2,
                correct_resid=True, correct_model=True, model_regularization="sum"),
    "post_hoc": dict(min_dly=2.0 / 0.3, offset=2.0 / 0.3, sky_model=None, maxsteps=60, tol=0.0, learning_rate=1e-2,
                     correct_resid=False, correct_model=False, model_regularization="post_hoc", nsamples_in_weights=False),
}


def driver_inputs():
    uvd = fx.line_array()
    uvd = fx.project_on_dpss(uvd, fx.dpss_vectors(uvd))
    uvd = fx.add_noise_like_eor(uvd, level_db=-40.0)
    rng = np.random.default_rng(77)
    uvd.flag_array[:] = rng.random(uvd.flag_array.shape) < 0.05
    uvd.nsample_array[:] = rng.integers(1, 3, uvd.nsample_array.shape)
    return uvd, fx.randomized_gains(uvd)
